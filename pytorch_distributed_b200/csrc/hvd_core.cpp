#include "hvd_core.h"

#include <algorithm>

namespace ptd {

using clk = std::chrono::steady_clock;

FusionQueue::FusionQueue(int64_t fusion_threshold_bytes, double cycle_time_ms, int64_t cycle_bytes)
    : threshold_(std::max<int64_t>(fusion_threshold_bytes, 1)), cycle_ms_(cycle_time_ms), cycle_bytes_(cycle_bytes) {}

void FusionQueue::close_open_locked() {
  if (open_.empty()) return;
  closed_.emplace_back(std::move(open_));
  open_.clear();
  open_bytes_ = 0;
  cv_groups_.notify_one();
}

int64_t FusionQueue::enqueue(const std::string& name, int64_t nbytes, int64_t order_key) {
  std::lock_guard<std::mutex> lk(mu_);
  const int64_t h = next_handle_++;
  // A tensor that would overflow the fusion buffer closes the current group first (horovod semantics:
  // a fused response never exceeds the threshold unless a single tensor does).
  const int64_t limit = limit_locked();
  if (!open_.empty() && open_bytes_ + nbytes > threshold_) close_open_locked();
  open_.push_back(Entry{h, name, nbytes, order_key, clk::now()});
  open_bytes_ += nbytes;
  done_[h] = false;
  ++outstanding_;
  if (open_bytes_ >= limit) close_open_locked();
  return h;
}

void FusionQueue::flush() {
  std::lock_guard<std::mutex> lk(mu_);
  close_open_locked();
}

std::vector<int64_t> FusionQueue::next_group(double timeout_ms) {
  std::unique_lock<std::mutex> lk(mu_);
  const auto deadline = clk::now() + std::chrono::microseconds((int64_t)(timeout_ms * 1000));
  while (closed_.empty() && !shutdown_) {
    // wake up at least once per cycle so a stalled producer is visible in the timeline
    const auto step = std::min(deadline, clk::now() + std::chrono::microseconds((int64_t)(std::max(cycle_ms_, 0.05) * 1000)));
    cv_groups_.wait_until(lk, step);
    if (clk::now() >= deadline) break;
  }
  std::vector<int64_t> out;
  if (closed_.empty()) return out;
  auto group = std::move(closed_.front());
  closed_.pop_front();
  const auto now = clk::now();
  ++n_groups_;
  for (const auto& e : group) {
    out.push_back(e.handle);
    ++n_tensors_;
    bytes_total_ += e.nbytes;
    const double ms = std::chrono::duration<double, std::milli>(now - e.t_enq).count();
    queue_ms_total_ += ms;
    queue_ms_max_ = std::max(queue_ms_max_, ms);
    if (timeline_on_ && timeline_.size() < (1u << 20)) {
      timeline_.emplace_back(e.name, e.nbytes, n_groups_, std::chrono::duration<double, std::micro>(e.t_enq - t0_).count(),
                             std::chrono::duration<double, std::micro>(now - t0_).count());
    }
  }
  return out;
}

void FusionQueue::mark_done(const std::vector<int64_t>& handles) {
  std::lock_guard<std::mutex> lk(mu_);
  for (int64_t h : handles) {
    auto it = done_.find(h);
    if (it != done_.end()) {
      done_.erase(it);       // nobody reaps per-handle state: keeping it would grow the map by one entry per tensor per step
      --outstanding_;
    }
  }
  cv_done_.notify_all();
}

bool FusionQueue::wait_idle(double timeout_ms) {
  std::unique_lock<std::mutex> lk(mu_);
  const auto deadline = clk::now() + std::chrono::microseconds((int64_t)(timeout_ms * 1000));
  woken_ = false;
  while (outstanding_ > 0 && !shutdown_ && !woken_) {
    if (cv_done_.wait_until(lk, deadline) == std::cv_status::timeout) break;
  }
  return outstanding_ == 0;
}

void FusionQueue::wake() {
  std::lock_guard<std::mutex> lk(mu_);
  woken_ = true;
  cv_done_.notify_all();
}

void FusionQueue::set_cycle_bytes(int64_t n) {
  std::lock_guard<std::mutex> lk(mu_);
  cycle_bytes_ = n;
}

int64_t FusionQueue::cycle_bytes() {
  std::lock_guard<std::mutex> lk(mu_);
  return cycle_bytes_;
}

void FusionQueue::enable_timeline(bool on) {
  std::lock_guard<std::mutex> lk(mu_);
  timeline_on_ = on;
}

std::vector<std::tuple<std::string, int64_t, int64_t, double, double>> FusionQueue::timeline() {
  std::lock_guard<std::mutex> lk(mu_);
  auto out = std::move(timeline_);
  timeline_.clear();
  return out;
}

bool FusionQueue::wait(int64_t handle, double timeout_ms) {
  std::unique_lock<std::mutex> lk(mu_);
  const auto deadline = clk::now() + std::chrono::microseconds((int64_t)(timeout_ms * 1000));
  while (true) {
    auto it = done_.find(handle);
    if (it == done_.end()) return true;  // unknown or already done (mark_done erases)
    if (shutdown_) return false;
    if (cv_done_.wait_until(lk, deadline) == std::cv_status::timeout) return false;
  }
}

int64_t FusionQueue::pending() {
  std::lock_guard<std::mutex> lk(mu_);
  return outstanding_;
}

void FusionQueue::shutdown() {
  std::lock_guard<std::mutex> lk(mu_);
  shutdown_ = true;
  cv_groups_.notify_all();
  cv_done_.notify_all();
}

std::map<std::string, double> FusionQueue::stats() {
  std::lock_guard<std::mutex> lk(mu_);
  return {{"groups", (double)n_groups_},
          {"tensors", (double)n_tensors_},
          {"bytes", (double)bytes_total_},
          {"tensors_per_group", n_groups_ ? (double)n_tensors_ / n_groups_ : 0.0},
          {"queue_ms_mean", n_tensors_ ? queue_ms_total_ / n_tensors_ : 0.0},
          {"queue_ms_max", queue_ms_max_}};
}

}  // namespace ptd
