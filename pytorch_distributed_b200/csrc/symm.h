#pragma once
#include <cuda.h>

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "comm_types.h"

namespace ptd {

bool multicast_supported(int device);

class SymmArena {
 public:
  // multi-process: this rank's buffer on `device`; peers are mapped later with map_peer()
  SymmArena(int device, int rank, int world, int64_t bytes);
  // single process, one buffer per device, everything mapped (DataParallel engine)
  static std::shared_ptr<SymmArena> create_local(const std::vector<int>& devices, int64_t bytes, bool want_multicast);
  // adopt pointers produced elsewhere (torch.distributed._symmetric_memory fallback)
  static std::shared_ptr<SymmArena> from_pointers(int rank, int world, const std::vector<int64_t>& ptrs, int64_t mc_ptr, int64_t bytes,
                                                  int device);
  ~SymmArena();
  SymmArena(const SymmArena&) = delete;
  SymmArena& operator=(const SymmArena&) = delete;

  int export_fd();
  void open_socket(const std::string& name);
  void send_fd(const std::string& peer_name, int fd, int tag);
  std::vector<int> recv_fd();  // {tag, src_rank, fd}
  void map_peer(int peer, int fd);
  int mc_create();
  void mc_import(int fd);
  void mc_add_device();
  void mc_bind_and_map();
  void disable_multicast(const std::string& why);

  CommCtx ctx(int channel, int as_rank = 0) const;
  int rank() const { return rank_; }
  int world() const { return world_; }
  int64_t bytes() const { return (int64_t)bytes_; }
  int64_t ptr(int r) const { return (int64_t)ptrs_[r]; }
  int64_t mc_ptr() const { return (int64_t)mc_ptr_; }
  bool multicast_candidate() const { return mc_ok_; }
  bool has_multicast() const { return mc_ptr_ != 0; }
  const std::string& mc_error() const { return mc_error_; }
  int device(int r = 0) const { return devices_[single_process_ ? r : 0]; }
  bool single_process() const { return single_process_; }
  uint32_t status() const { return status_ ? *reinterpret_cast<volatile uint32_t*>(status_) : 0; }
  void set_timeout_ms(uint32_t ms) { timeout_ms_ = ms; }
  int64_t ll_seq_ptr(int r = 0) const { return (int64_t)(seq_[single_process_ ? r : 0] + kMaxChannels * kMaxBlocks); }

 private:
  SymmArena() = default;
  void init_local_state(int device);

  int rank_ = 0, world_ = 1;
  std::vector<int> devices_;
  size_t bytes_ = 0, gran_ = 0;
  std::vector<CUmemGenericAllocationHandle> handles_;
  std::vector<CUdeviceptr> ptrs_;
  CUmemGenericAllocationHandle mc_handle_ = 0;
  CUdeviceptr mc_ptr_ = 0;
  bool mc_ok_ = false, single_process_ = false, borrowed_ = false;
  std::string mc_error_;
  std::vector<uint32_t*> seq_;
  uint32_t* status_ = nullptr;
  uint32_t timeout_ms_ = 30000;
  int sock_ = -1;
};

}  // namespace ptd
