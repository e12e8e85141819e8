// Host-side declarations shared by the .cu / .cpp translation units and the bindings.
#pragma once
#include <torch/extension.h>

#include <cstdint>
#include <string>
#include <vector>

#include "comm_types.h"

namespace ptd {

// One contiguous piece of one tensor inside one CTA's arena range (static per plan).
struct Seg {
  int32_t tensor;     // index into the launch's pointer pack
  int32_t len;        // elements
  int64_t src_off;    // element offset inside the tensor
  int64_t arena_off;  // element offset from the plan's element 0
};
static_assert(sizeof(Seg) == 24, "Seg layout is mirrored by numpy in parallel/plan.py");

// ---- collectives.cu
void launch_plan(const CommCtx& ctx, int kind, int wire_dtype, bool nvls, int grid, const std::vector<at::Tensor>& tensors,
                 int64_t seg_begin_ptr, int64_t segs_ptr, int64_t data_off_bytes, int64_t block_elems, int64_t plan_calls_ptr,
                 int64_t found_inf_ptr, double scale, bool writeback, int root, int flags = 0, int64_t result_off_bytes = -1);
at::Tensor pack_pointers(const std::vector<at::Tensor>& tensors);
void launch_barrier(const CommCtx& ctx);
void launch_metrics(const CommCtx& ctx, const at::Tensor& logits, const at::Tensor& target, const c10::optional<at::Tensor>& loss,
                    int64_t ll_seq_ptr, at::Tensor out);
void launch_ll_allreduce(const CommCtx& ctx, const at::Tensor& in, at::Tensor out, double scale, int64_t ll_seq_ptr);

// ---- optim.cu
void fused_sgd_flat(at::Tensor grad, at::Tensor master, at::Tensor momentum, c10::optional<at::Tensor> model_copy,
                    at::Tensor hyper, c10::optional<at::Tensor> found_inf, bool nesterov, bool first_step);
void fused_sgd_multi(std::vector<at::Tensor> grads, std::vector<at::Tensor> params, std::vector<at::Tensor> momenta,
                     std::vector<at::Tensor> model_copies, at::Tensor hyper, c10::optional<at::Tensor> found_inf, bool nesterov,
                     bool first_step);
void multi_tensor_scale(std::vector<at::Tensor> src, std::vector<at::Tensor> dst, double scale, at::Tensor found_inf);
void multi_tensor_axpby(std::vector<at::Tensor> x, std::vector<at::Tensor> y, std::vector<at::Tensor> out, double a, double b,
                        at::Tensor found_inf);
void amp_update_scale(at::Tensor scale, at::Tensor growth_tracker, at::Tensor found_inf, double growth, double backoff,
                      int64_t interval, at::Tensor hyper);

// ---- bn_act.cu
std::vector<at::Tensor> bn_act_forward(const at::Tensor& x, const c10::optional<at::Tensor>& residual, const at::Tensor& weight,
                                       const at::Tensor& bias, at::Tensor running_mean, at::Tensor running_var,
                                       c10::optional<at::Tensor> num_batches_tracked, bool training, double momentum, double eps, bool relu,
                                       bool need_mask, at::Tensor work, bool stats_ready);
std::vector<at::Tensor> bn_act_backward(const at::Tensor& dy, const at::Tensor& x, const c10::optional<at::Tensor>& mask,
                                        const at::Tensor& weight, const at::Tensor& saved, bool relu, bool has_residual, at::Tensor work);
std::vector<at::Tensor> bn_act_backward2(const at::Tensor& dy_a, const at::Tensor& dy_b, const at::Tensor& x,
                                         const c10::optional<at::Tensor>& mask, const at::Tensor& weight, const at::Tensor& saved, bool relu,
                                         at::Tensor work);

std::vector<at::Tensor> stem_forward(const at::Tensor& x, const at::Tensor& weight, const at::Tensor& bias, at::Tensor running_mean,
                                     at::Tensor running_var, c10::optional<at::Tensor> num_batches_tracked, bool training, double momentum,
                                     double eps, bool need_code, at::Tensor work);
std::vector<at::Tensor> stem_forward_pre(const at::Tensor& x, const at::Tensor& weight, const at::Tensor& bias, at::Tensor running_mean,
                                         at::Tensor running_var, c10::optional<at::Tensor> num_batches_tracked, bool training, double momentum,
                                         double eps, bool need_code, at::Tensor work);
std::vector<at::Tensor> stem_backward(const at::Tensor& dp, const at::Tensor& x, const at::Tensor& code, const at::Tensor& weight,
                                      const at::Tensor& saved, at::Tensor work);

// ---- gemm_bnstats.cu (tcgen05 / TMA / TMEM)
at::Tensor conv1x1_bnstats(const at::Tensor& x, const at::Tensor& weight, at::Tensor gsum);

// ---- stem_conv.cu
at::Tensor stem_im2col(const at::Tensor& x);

// ---- data_ops.cu
at::Tensor normalize_nhwc(const at::Tensor& src, const at::Tensor& mean, const at::Tensor& std, int64_t out_dtype, bool channels_last);

void p2p_copy_multi(std::vector<at::Tensor> src, std::vector<at::Tensor> dst, int64_t run_device);

}  // namespace ptd
