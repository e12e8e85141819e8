// Symmetric-memory runtime: the data plane's allocator and rendezvous (replaces ProcessGroupNCCL's
// communicator setup behind /root/reference/distributed.py:132 and the single-process NCCL/peer-copy
// setup behind /root/reference/dataparallel.py:138).
//
// Every rank allocates one physical buffer with the CUDA VMM API (cuMemCreate), exports it as a POSIX file
// descriptor, ships the descriptor to its peers over an abstract-namespace unix socket (SCM_RIGHTS), and maps
// every peer's buffer into its own address space => base[r] is a load/store-able pointer to rank r's arena
// over NVLink.  If the devices support NVLS, one multicast object spanning all ranks is bound to the same
// physical memory and mapped a second time (mc_base): a store to mc_base+x is replicated by NVSwitch into all
// arenas, a multimem.ld_reduce from mc_base+x returns the in-switch sum of all arenas at x.
//
// The same class also serves the single-process engine (DataParallel): `create_local` builds the per-device
// buffers inside one process and grants every device access to every mapping.
//
// libcuda is resolved at run time through cudaGetDriverEntryPoint so the extension also loads on GPU-less
// build hosts.
#include <cuda.h>
#include <cuda_runtime.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <cstring>
#include <memory>
#include <sstream>
#include <stdexcept>

#include "symm.h"

namespace ptd {

namespace {

template <typename F>
F drv(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || fn == nullptr || q != cudaDriverEntryPointSuccess) {
    throw std::runtime_error(std::string("cannot resolve driver entry point ") + name);
  }
  return reinterpret_cast<F>(fn);
}

#define DRV(name) static auto p_##name = drv<decltype(&name)>(#name)

void check(CUresult r, const char* what) {
  if (r != CUDA_SUCCESS) {
    const char* msg = nullptr;
    static auto get_str = drv<CUresult (*)(CUresult, const char**)>("cuGetErrorString");
    get_str(r, &msg);
    std::ostringstream os;
    os << what << " failed: " << (msg ? msg : "?") << " (" << (int)r << ")";
    throw std::runtime_error(os.str());
  }
}
void rcheck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + " failed: " + cudaGetErrorString(e));
}

size_t round_up(size_t x, size_t g) { return (x + g - 1) / g * g; }

CUmemAllocationProp alloc_prop(int device) {
  CUmemAllocationProp prop;
  std::memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

void set_access(CUdeviceptr va, size_t size, const std::vector<int>& devices) {
  DRV(cuMemSetAccess);
  std::vector<CUmemAccessDesc> desc(devices.size());
  for (size_t i = 0; i < devices.size(); ++i) {
    desc[i].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    desc[i].location.id = devices[i];
    desc[i].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  }
  check(p_cuMemSetAccess(va, size, desc.data(), desc.size()), "cuMemSetAccess");
}

CUdeviceptr map_handle(CUmemGenericAllocationHandle h, size_t size, size_t gran, const std::vector<int>& devices) {
  DRV(cuMemAddressReserve);
  DRV(cuMemMap);
  CUdeviceptr va = 0;
  check(p_cuMemAddressReserve(&va, size, gran, 0, 0), "cuMemAddressReserve");
  check(p_cuMemMap(va, size, 0, h, 0), "cuMemMap");
  set_access(va, size, devices);
  return va;
}

sockaddr_un abstract_addr(const std::string& name, socklen_t* len) {
  sockaddr_un a;
  std::memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  if (name.size() + 2 > sizeof(a.sun_path)) throw std::runtime_error("socket name too long");
  std::memcpy(a.sun_path + 1, name.data(), name.size());  // leading NUL => abstract namespace (no filesystem entry)
  *len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + name.size());
  return a;
}

}  // namespace

bool multicast_supported(int device) {
  DRV(cuDeviceGetAttribute);
  DRV(cuDeviceGet);
  CUdevice dev;
  if (p_cuDeviceGet(&dev, device) != CUDA_SUCCESS) return false;
  int v = 0;
  if (p_cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS) return false;
  return v != 0;
}

SymmArena::SymmArena(int device, int rank, int world, int64_t bytes) : rank_(rank), world_(world) {
  if (world < 1 || world > kMaxWorld) throw std::runtime_error("world size out of range");
  devices_.assign(1, device);
  handles_.assign(world, 0);
  ptrs_.assign(world, 0);
  rcheck(cudaSetDevice(device), "cudaSetDevice");
  rcheck(cudaFree(nullptr), "context init");
  DRV(cuMemGetAllocationGranularity);
  DRV(cuMemCreate);
  CUmemAllocationProp prop = alloc_prop(device);
  size_t gran = 0;
  check(p_cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
  gran_ = gran;
  mc_ok_ = multicast_supported(device) && world > 1;
  if (mc_ok_) {
    DRV(cuMulticastGetGranularity);
    CUmulticastObjectProp mp;
    std::memset(&mp, 0, sizeof(mp));
    mp.numDevices = world;
    mp.size = round_up((size_t)bytes, gran);
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    if (p_cuMulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > 0) {
      if (mg > (size_t)512 << 20) p_cuMulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_MINIMUM);
      gran_ = std::max(gran_, mg);
    } else {
      mc_ok_ = false;
    }
  }
  bytes_ = round_up((size_t)bytes, gran_);
  check(p_cuMemCreate(&handles_[rank], bytes_, &prop, 0), "cuMemCreate");
  ptrs_[rank] = map_handle(handles_[rank], bytes_, gran_, devices_);
  rcheck(cudaMemset(reinterpret_cast<void*>(ptrs_[rank]), 0, bytes_), "cudaMemset(arena)");
  rcheck(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  init_local_state(device);
}

void SymmArena::init_local_state(int device) {
  rcheck(cudaSetDevice(device), "cudaSetDevice");
  void* s = nullptr;
  rcheck(cudaMalloc(&s, sizeof(uint32_t) * (kMaxChannels * kMaxBlocks + 64)), "cudaMalloc(seq)");
  rcheck(cudaMemset(s, 0, sizeof(uint32_t) * (kMaxChannels * kMaxBlocks + 64)), "cudaMemset(seq)");
  seq_.push_back(reinterpret_cast<uint32_t*>(s));
  if (status_ == nullptr) {
    rcheck(cudaHostAlloc(reinterpret_cast<void**>(&status_), 64, cudaHostAllocMapped | cudaHostAllocPortable), "cudaHostAlloc(status)");
    std::memset(status_, 0, 64);
  }
}

std::shared_ptr<SymmArena> SymmArena::create_local(const std::vector<int>& devices, int64_t bytes, bool want_multicast) {
  auto a = std::shared_ptr<SymmArena>(new SymmArena());
  const int world = (int)devices.size();
  if (world < 1 || world > kMaxWorld) throw std::runtime_error("device count out of range");
  a->rank_ = 0;
  a->world_ = world;
  a->devices_ = devices;
  a->handles_.assign(world, 0);
  a->ptrs_.assign(world, 0);
  a->single_process_ = true;
  DRV(cuMemGetAllocationGranularity);
  DRV(cuMemCreate);
  size_t gran = 0;
  for (int d : devices) {
    rcheck(cudaSetDevice(d), "cudaSetDevice");
    rcheck(cudaFree(nullptr), "context init");
    CUmemAllocationProp prop = alloc_prop(d);
    size_t g = 0;
    check(p_cuMemGetAllocationGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
    gran = std::max(gran, g);
  }
  a->gran_ = gran;
  a->mc_ok_ = want_multicast && world > 1;
  for (int d : devices) a->mc_ok_ = a->mc_ok_ && multicast_supported(d);
  CUmulticastObjectProp mp;
  std::memset(&mp, 0, sizeof(mp));
  if (a->mc_ok_) {
    DRV(cuMulticastGetGranularity);
    mp.numDevices = world;
    mp.size = round_up((size_t)bytes, gran);
    size_t mg = 0;
    if (p_cuMulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > 0) {
      if (mg > (size_t)512 << 20) p_cuMulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_MINIMUM);
      a->gran_ = std::max(a->gran_, mg);
    } else {
      a->mc_ok_ = false;
    }
  }
  a->bytes_ = round_up((size_t)bytes, a->gran_);
  for (int i = 0; i < world; ++i) {
    rcheck(cudaSetDevice(devices[i]), "cudaSetDevice");
    CUmemAllocationProp prop = alloc_prop(devices[i]);
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_NONE;
    check(p_cuMemCreate(&a->handles_[i], a->bytes_, &prop, 0), "cuMemCreate");
    a->ptrs_[i] = map_handle(a->handles_[i], a->bytes_, a->gran_, devices);
    rcheck(cudaMemset(reinterpret_cast<void*>(a->ptrs_[i]), 0, a->bytes_), "cudaMemset(arena)");
    rcheck(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    a->init_local_state(devices[i]);
  }
  if (a->mc_ok_) {
    try {
      DRV(cuMulticastCreate);
      DRV(cuMulticastAddDevice);
      DRV(cuMulticastBindMem);
      DRV(cuDeviceGet);
      mp.size = a->bytes_;
      mp.handleTypes = 0;
      check(p_cuMulticastCreate(&a->mc_handle_, &mp), "cuMulticastCreate");
      for (int d : devices) {
        CUdevice dev;
        check(p_cuDeviceGet(&dev, d), "cuDeviceGet");
        check(p_cuMulticastAddDevice(a->mc_handle_, dev), "cuMulticastAddDevice");
      }
      for (int i = 0; i < world; ++i) {
        rcheck(cudaSetDevice(devices[i]), "cudaSetDevice");
        check(p_cuMulticastBindMem(a->mc_handle_, 0, a->handles_[i], 0, a->bytes_, 0), "cuMulticastBindMem");
      }
      a->mc_ptr_ = map_handle(a->mc_handle_, a->bytes_, a->gran_, devices);
    } catch (const std::exception& e) {
      a->mc_ok_ = false;
      a->mc_ptr_ = 0;
      a->mc_error_ = e.what();
    }
  }
  return a;
}

std::shared_ptr<SymmArena> SymmArena::from_pointers(int rank, int world, const std::vector<int64_t>& ptrs, int64_t mc_ptr, int64_t bytes,
                                                    int device) {
  auto a = std::shared_ptr<SymmArena>(new SymmArena());
  a->rank_ = rank;
  a->world_ = world;
  a->devices_.assign(1, device);
  a->ptrs_.resize(world);
  for (int i = 0; i < world; ++i) a->ptrs_[i] = (CUdeviceptr)ptrs[i];
  a->mc_ptr_ = (CUdeviceptr)mc_ptr;
  a->mc_ok_ = mc_ptr != 0;
  a->bytes_ = (size_t)bytes;
  a->borrowed_ = true;
  a->init_local_state(device);
  return a;
}

SymmArena::~SymmArena() {
  if (sock_ >= 0) ::close(sock_);
  if (borrowed_) return;
  // Best effort teardown; errors are ignored (the context may already be gone at interpreter exit).
  try {
    DRV(cuMemUnmap);
    DRV(cuMemAddressFree);
    DRV(cuMemRelease);
    if (mc_ptr_) { p_cuMemUnmap(mc_ptr_, bytes_); p_cuMemAddressFree(mc_ptr_, bytes_); }
    for (size_t i = 0; i < ptrs_.size(); ++i)
      if (ptrs_[i]) { p_cuMemUnmap(ptrs_[i], bytes_); p_cuMemAddressFree(ptrs_[i], bytes_); }
    if (mc_handle_) p_cuMemRelease(mc_handle_);
    for (auto h : handles_) if (h) p_cuMemRelease(h);
  } catch (...) {
  }
}

// ------------------------------------------------------------------ descriptor exchange
int SymmArena::export_fd() {
  DRV(cuMemExportToShareableHandle);
  int fd = -1;
  check(p_cuMemExportToShareableHandle(&fd, handles_[rank_], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle");
  return fd;
}

void SymmArena::open_socket(const std::string& name) {
  sock_ = ::socket(AF_UNIX, SOCK_DGRAM, 0);
  if (sock_ < 0) throw std::runtime_error("socket() failed");
  socklen_t len;
  sockaddr_un a = abstract_addr(name, &len);
  if (::bind(sock_, reinterpret_cast<sockaddr*>(&a), len) != 0) throw std::runtime_error("bind(" + name + ") failed: " + std::strerror(errno));
  timeval tv{120, 0};
  ::setsockopt(sock_, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
}

void SymmArena::send_fd(const std::string& peer_name, int fd, int tag) {
  socklen_t len;
  sockaddr_un a = abstract_addr(peer_name, &len);
  int payload[2] = {tag, rank_};
  iovec iov{payload, sizeof(payload)};
  char ctrl[CMSG_SPACE(sizeof(int))];
  std::memset(ctrl, 0, sizeof(ctrl));
  msghdr msg;
  std::memset(&msg, 0, sizeof(msg));
  msg.msg_name = &a;
  msg.msg_namelen = len;
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int));
  std::memcpy(CMSG_DATA(c), &fd, sizeof(int));
  if (::sendmsg(sock_, &msg, 0) < 0) throw std::runtime_error("sendmsg to " + peer_name + " failed: " + std::strerror(errno));
}

std::vector<int> SymmArena::recv_fd() {
  int payload[2] = {0, 0};
  iovec iov{payload, sizeof(payload)};
  char ctrl[CMSG_SPACE(sizeof(int))];
  msghdr msg;
  std::memset(&msg, 0, sizeof(msg));
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  if (::recvmsg(sock_, &msg, 0) < 0) throw std::runtime_error(std::string("recvmsg failed: ") + std::strerror(errno));
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  if (!c || c->cmsg_type != SCM_RIGHTS) throw std::runtime_error("no descriptor in message");
  int fd = -1;
  std::memcpy(&fd, CMSG_DATA(c), sizeof(int));
  return {payload[0], payload[1], fd};
}

void SymmArena::map_peer(int peer, int fd) {
  DRV(cuMemImportFromShareableHandle);
  check(p_cuMemImportFromShareableHandle(&handles_[peer], reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                         CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), "cuMemImportFromShareableHandle");
  ::close(fd);
  ptrs_[peer] = map_handle(handles_[peer], bytes_, gran_, devices_);
}

int SymmArena::mc_create() {
  DRV(cuMulticastCreate);
  DRV(cuMemExportToShareableHandle);
  CUmulticastObjectProp mp;
  std::memset(&mp, 0, sizeof(mp));
  mp.numDevices = world_;
  mp.size = bytes_;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  check(p_cuMulticastCreate(&mc_handle_, &mp), "cuMulticastCreate");
  int fd = -1;
  check(p_cuMemExportToShareableHandle(&fd, mc_handle_, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "export multicast handle");
  return fd;
}

void SymmArena::mc_import(int fd) {
  DRV(cuMemImportFromShareableHandle);
  check(p_cuMemImportFromShareableHandle(&mc_handle_, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                         CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), "import multicast handle");
  ::close(fd);
}

void SymmArena::mc_add_device() {
  DRV(cuMulticastAddDevice);
  DRV(cuDeviceGet);
  CUdevice dev;
  check(p_cuDeviceGet(&dev, devices_[0]), "cuDeviceGet");
  check(p_cuMulticastAddDevice(mc_handle_, dev), "cuMulticastAddDevice");
}

void SymmArena::mc_bind_and_map() {
  DRV(cuMulticastBindMem);
  check(p_cuMulticastBindMem(mc_handle_, 0, handles_[rank_], 0, bytes_, 0), "cuMulticastBindMem");
  mc_ptr_ = map_handle(mc_handle_, bytes_, gran_, devices_);
}

void SymmArena::disable_multicast(const std::string& why) {
  mc_ok_ = false;
  mc_ptr_ = 0;
  mc_error_ = why;
}

CommCtx SymmArena::ctx(int channel, int as_rank) const {
  CommCtx c;
  std::memset(&c, 0, sizeof(c));
  c.rank = single_process_ ? as_rank : rank_;
  c.world = world_;
  c.channel = channel;
  c.timeout_ms = timeout_ms_;
  for (int i = 0; i < world_; ++i) c.base[i] = reinterpret_cast<char*>(ptrs_[i]);
  c.mc_base = reinterpret_cast<char*>(mc_ptr_);
  c.seq = seq_[single_process_ ? as_rank : 0];
  c.status = status_;
  return c;
}

}  // namespace ptd
