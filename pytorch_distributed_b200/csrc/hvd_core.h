// Horovod-style tensor-fusion core (replaces the C++ background loop that
// /root/reference/horovod_distributed.py:125,159-164 pulls in through `import horovod.torch`).
//
// Gradient hooks enqueue (name, bytes) requests from the autograd thread; the core packs them into fusion groups
// (close a group when it reaches the fusion threshold, or when the optimizer asks for a flush) and hands closed
// groups to a dispatcher thread that launches ONE fused peer-memory all-reduce per group.  Group composition only
// depends on enqueue order and sizes, which are identical on every rank for the same model, so - unlike horovod's
// MPI/gloo coordinator - no cross-rank negotiation round is needed; `cycle_time_ms` only bounds how long a closed
// group may sit before the dispatcher wakes up.
#pragma once
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

namespace ptd {

class FusionQueue {
 public:
  // A group closes when it reaches `fusion_threshold_bytes` (horovod's fusion buffer size) OR `cycle_bytes` - the
  // deterministic stand-in for horovod's cycle time: closing on a wall-clock tick would need a cross-rank negotiation
  // round to agree on the group's content, closing on a byte budget gives every rank the same groups for free and still
  // lets the all-reduce of early gradients overlap the rest of backward.  `cycle_bytes` <= 0 disables it.
  FusionQueue(int64_t fusion_threshold_bytes, double cycle_time_ms, int64_t cycle_bytes = 0);
  int64_t enqueue(const std::string& name, int64_t nbytes, int64_t order_key);
  // Blocks (up to timeout_ms) for the next closed group; returns the handles in enqueue order (empty on timeout/shutdown).
  std::vector<int64_t> next_group(double timeout_ms);
  void flush();
  void mark_done(const std::vector<int64_t>& handles);
  bool wait(int64_t handle, double timeout_ms);
  // Blocks until every enqueued request has been marked done (true), or timeout / shutdown / wake() (false).
  bool wait_idle(double timeout_ms);
  void wake();                       // interrupt wait_idle (the dispatcher reports an error)
  void set_cycle_bytes(int64_t n);   // autotuner
  int64_t cycle_bytes();
  int64_t pending();
  void shutdown();
  std::map<std::string, double> stats();
  // timeline: one record per tensor {name, bytes, group, enqueue_us, dispatch_us} since the last call (bounded ring)
  void enable_timeline(bool on);
  std::vector<std::tuple<std::string, int64_t, int64_t, double, double>> timeline();

 private:
  struct Entry { int64_t handle; std::string name; int64_t nbytes; int64_t order_key; std::chrono::steady_clock::time_point t_enq; };
  void close_open_locked();

  int64_t limit_locked() const { return (cycle_bytes_ > 0 && cycle_bytes_ < threshold_) ? cycle_bytes_ : threshold_; }

  const int64_t threshold_;
  const double cycle_ms_;
  int64_t cycle_bytes_ = 0;
  bool woken_ = false, timeline_on_ = false;
  std::chrono::steady_clock::time_point t0_ = std::chrono::steady_clock::now();
  std::vector<std::tuple<std::string, int64_t, int64_t, double, double>> timeline_;
  std::mutex mu_;
  std::condition_variable cv_groups_, cv_done_;
  std::vector<Entry> open_;
  int64_t open_bytes_ = 0;
  std::deque<std::vector<Entry>> closed_;
  std::unordered_map<int64_t, bool> done_;
  int64_t next_handle_ = 1, outstanding_ = 0;
  bool shutdown_ = false;
  // timeline counters
  int64_t n_groups_ = 0, n_tensors_ = 0, bytes_total_ = 0;
  double queue_ms_total_ = 0, queue_ms_max_ = 0;
};

}  // namespace ptd
