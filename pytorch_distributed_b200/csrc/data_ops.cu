// Input pipeline and single-process multi-GPU helpers.
//
//  normalize_nhwc : the on-GPU half of the reference's `data_prefetcher` (/root/reference/apex_distributed.py:115-169):
//                   uint8/float NCHW batch -> (x*a[c] + b[c]) -> bf16/fp16/fp32, optionally re-laid out as NHWC, in
//                   ONE pass instead of .float() + sub_ + div_ (+ a later layout/dtype conversion inside the model).
//  p2p_copy_multi : multi-tensor copy whose sources/destinations may live on peer GPUs (scatter / gather of
//                   nn.DataParallel, /root/reference/dataparallel.py:138,246) - one kernel, NVLink loads/stores.
#include <ATen/cuda/CUDAContext.h>
#include <ATen/cuda/PeerToPeerAccess.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "common.cuh"
#include "host.h"

namespace ptd {

template <typename In> __device__ __forceinline__ float in_to_f32(In v);
template <> __device__ __forceinline__ float in_to_f32<uint8_t>(uint8_t v) { return (float)v; }
template <> __device__ __forceinline__ float in_to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float in_to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float in_to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// One thread per pixel: reads the C planes (coalesced across the warp per plane), writes C contiguous outputs (NHWC)
// or C planes (NCHW).
template <typename In, typename Out, int C, bool NHWC_OUT>
__global__ void __launch_bounds__(256) normalize_kernel(const In* __restrict__ src, Out* __restrict__ dst, const float* __restrict__ a,
                                                        const float* __restrict__ b, int64_t hw, int64_t total_pixels) {
  float sa[C], sb[C];
#pragma unroll
  for (int c = 0; c < C; ++c) { sa[c] = a[c]; sb[c] = b[c]; }
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total_pixels; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = p / hw, i = p - n * hw;
    const In* s = src + n * C * hw + i;
    float v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = in_to_f32<In>(s[c * hw]) * sa[c] + sb[c];
    if constexpr (NHWC_OUT) {
      Out* d = dst + p * C;
#pragma unroll
      for (int c = 0; c < C; ++c) d[c] = from_f32<Out>(v[c]);
    } else {
      Out* d = dst + n * C * hw + i;
#pragma unroll
      for (int c = 0; c < C; ++c) d[c * hw] = from_f32<Out>(v[c]);
    }
  }
}

template <typename In, typename Out>
static void launch_norm(const at::Tensor& src, at::Tensor& dst, const at::Tensor& a, const at::Tensor& b, bool nhwc) {
  const int64_t hw = src.size(2) * src.size(3), total = src.size(0) * hw;
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sms * 16);
  cudaStream_t st = at::cuda::getCurrentCUDAStream();
  const In* s = reinterpret_cast<const In*>(src.data_ptr());
  Out* d = reinterpret_cast<Out*>(dst.data_ptr());
  if (nhwc) normalize_kernel<In, Out, 3, true><<<grid, 256, 0, st>>>(s, d, a.data_ptr<float>(), b.data_ptr<float>(), hw, total);
  else      normalize_kernel<In, Out, 3, false><<<grid, 256, 0, st>>>(s, d, a.data_ptr<float>(), b.data_ptr<float>(), hw, total);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// dst[n,c,h,w] = src[n,c,h,w] * a[c] + b[c]; src is NCHW-contiguous with C == 3.
at::Tensor normalize_nhwc(const at::Tensor& src, const at::Tensor& a, const at::Tensor& b, int64_t out_dtype, bool channels_last) {
  TORCH_CHECK(src.is_cuda() && src.dim() == 4 && src.size(1) == 3 && src.is_contiguous(), "normalize expects a contiguous NCHW batch with 3 channels");
  TORCH_CHECK(a.scalar_type() == at::kFloat && b.scalar_type() == at::kFloat && a.numel() == 3 && b.numel() == 3);
  c10::cuda::CUDAGuard guard(src.device());
  const at::ScalarType ot = out_dtype == kBF16 ? at::kBFloat16 : out_dtype == kF16 ? at::kHalf : at::kFloat;
  at::Tensor dst = at::empty(src.sizes(), src.options().dtype(ot).memory_format(channels_last ? at::MemoryFormat::ChannelsLast : at::MemoryFormat::Contiguous));
#define NORM_OUT(In) \
  switch (ot) { \
    case at::kBFloat16: launch_norm<In, __nv_bfloat16>(src, dst, a, b, channels_last); break; \
    case at::kHalf: launch_norm<In, __half>(src, dst, a, b, channels_last); break; \
    default: launch_norm<In, float>(src, dst, a, b, channels_last); break; \
  }
  switch (src.scalar_type()) {
    case at::kByte: NORM_OUT(uint8_t); break;
    case at::kFloat: NORM_OUT(float); break;
    case at::kHalf: NORM_OUT(__half); break;
    case at::kBFloat16: NORM_OUT(__nv_bfloat16); break;
    default: TORCH_CHECK(false, "unsupported input dtype for normalize");
  }
#undef NORM_OUT
  return dst;
}

// ------------------------------------------------------------------ peer copies
constexpr int kCopyItems = 96;
struct CopyArgs {
  const char* src[kCopyItems];
  char* dst[kCopyItems];
  int64_t bytes[kCopyItems];
  int n;
};

__global__ void __launch_bounds__(512) p2p_copy_kernel(const __grid_constant__ CopyArgs a) {
  for (int t = blockIdx.y; t < a.n; t += gridDim.y) {
    const char* s = a.src[t];
    char* d = a.dst[t];
    const int64_t nb = a.bytes[t];
    const bool vec = (((uintptr_t)s | (uintptr_t)d) & 15) == 0;
    const int64_t nv = vec ? nb >> 4 : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
      V4 v = ld_sys(s + (i << 4));
      st_sys(d + (i << 4), v);
    }
    for (int64_t i = (nv << 4) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += (int64_t)gridDim.x * blockDim.x) d[i] = s[i];
  }
}

// Copy src[i] -> dst[i] (same numel/dtype, dense) with a kernel running on `run_device`; tensors may live on any peer.
void p2p_copy_multi(std::vector<at::Tensor> src, std::vector<at::Tensor> dst, int64_t run_device) {
  TORCH_CHECK(src.size() == dst.size());
  if (src.empty()) return;
  c10::cuda::CUDAGuard guard((c10::DeviceIndex)run_device);
  cudaStream_t st = at::cuda::getCurrentCUDAStream();
  size_t i = 0;
  while (i < src.size()) {
    CopyArgs a;
    a.n = 0;
    int64_t maxb = 0;
    for (; i < src.size() && a.n < kCopyItems; ++i) {
      const auto& s = src[i];
      auto& d = dst[i];
      TORCH_CHECK(s.numel() == d.numel() && s.scalar_type() == d.scalar_type(), "p2p copy: shape/dtype mismatch");
      TORCH_CHECK(s.is_non_overlapping_and_dense() && d.is_non_overlapping_and_dense() && s.strides() == d.strides(), "p2p copy: layouts must match");
      for (auto dev : {s.get_device(), d.get_device()})
        if (dev != run_device) TORCH_CHECK(at::cuda::get_p2p_access((c10::DeviceIndex)run_device, (c10::DeviceIndex)dev), "no peer access ", run_device, "->", dev);
      a.src[a.n] = reinterpret_cast<const char*>(s.data_ptr());
      a.dst[a.n] = reinterpret_cast<char*>(d.data_ptr());
      a.bytes[a.n] = s.numel() * s.element_size();
      maxb = std::max(maxb, a.bytes[a.n]);
      ++a.n;
    }
    const int gx = (int)std::max<int64_t>(1, std::min<int64_t>((maxb / 16 + 511) / 512, 32));
    const int gy = std::min(a.n, 16);
    p2p_copy_kernel<<<dim3(gx, gy), 512, 0, st>>>(a);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
  }
}

}  // namespace ptd
