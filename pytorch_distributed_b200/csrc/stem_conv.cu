// ResNet stem convolution (7x7, stride 2, pad 3, C_in = 3) as im2col + the tcgen05 GEMM of gemm_bnstats.cu (sm_100a).
//
// cuDNN serves this layer with legacy sm80 kernels (C_in = 3 fits no tensor-core tile): 1.5 ms forward + 0.8 ms wgrad per
// 256-image step, 10 % of the whole ResNet-50 step (profiles/step_breakdown_r1.md; padding C_in to 4 or 8 does not help,
// tools/conv_stem_probe.py).  The layer is only 60 GFLOP; written as a GEMM it is bound by its 411 MB output:
//   A[M, 192]  = im2col(x)        M = N*OH*OW output pixels, one 384-byte row per pixel        (this file)
//   Y[M, 64]   = A x Wp^T         persistent tcgen05 GEMM, BatchNorm statistics in its epilogue  (gemm_bnstats.cu)
//   dWp[64,192]= dY^T x A         library GEMM over the saved A                                  (ops/stem_conv.py)
// K ordering of a row: k = r*24 + s*3 + c for filter row r < 7, filter column s < 7, channel c < 3; positions with
// s*3 + c >= 21 and k >= 168 are zero (the packed weights are zero there too).  A filter row is 21 CONTIGUOUS input
// elements in NHWC, so a 16-byte granule of A is 8 consecutive input elements: one thread builds one granule.
// Reference call site: torchvision resnet.conv1 reached through /root/reference/distributed.py:136-139.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "common.cuh"
#include "host.h"

namespace ptd {

constexpr int kStemK = 192;            // padded GEMM K (3 x 64)
constexpr int kStemRowK = 24;          // padded elements per filter row (21 real)
constexpr int kStemGranules = kStemK / 8;

__global__ void __launch_bounds__(256) stem_im2col_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ a, int H, int W,
                                                          int OH, int OW, int64_t total) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int row_elems = W * 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int gq = (int)(i % kStemGranules);
    const int64_t pix = i / kStemGranules;
    const int ow = (int)(pix % OW);
    const int64_t t = pix / OW;
    const int oh = (int)(t % OH);
    const int64_t n = t / OH;
    const int r = gq / 3, q = gq - 3 * r;               // filter row, granule inside the row (r == 7: zero padding of K)
    unsigned short v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0;
    const int ih = 2 * oh - 3 + r;
    if (r < 7 && ih >= 0 && ih < H) {
      const unsigned short* row = reinterpret_cast<const unsigned short*>(x) + (n * H + ih) * (int64_t)row_elems;
      const int e0 = (2 * ow - 3) * 3 + q * 8;           // element offset of this granule inside the input row
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int sc = q * 8 + j, e = e0 + j;
        if (sc < 21 && e >= 0 && e < row_elems) v[j] = __ldg(row + e);
      }
    }
    V4 o{(uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16), (uint32_t)v[4] | ((uint32_t)v[5] << 16),
         (uint32_t)v[6] | ((uint32_t)v[7] << 16)};
    st_v4(a + i * 8, o);
  }
}

// x: [N, 3, H, W] channels_last bf16 (physically N x H x W x 3).  returns A as a [N, 192, OH, OW] channels_last view
// (physically [N*OH*OW, 192] row-major), which is exactly the activation layout conv1x1_bnstats() takes.
at::Tensor stem_im2col(const at::Tensor& x) {
  TORCH_CHECK(x.is_cuda() && x.dim() == 4 && x.size(1) == 3 && x.scalar_type() == at::kBFloat16 &&
                  x.is_contiguous(at::MemoryFormat::ChannelsLast),
              "stem_im2col: x must be a [N, 3, H, W] channels_last bf16 CUDA tensor");
  const int64_t N = x.size(0);
  const int H = (int)x.size(2), W = (int)x.size(3);
  const int OH = (H + 6 - 7) / 2 + 1, OW = (W + 6 - 7) / 2 + 1;
  TORCH_CHECK(H >= 7 && W >= 7 && (int64_t)W * 3 < (1 << 30), "stem_im2col: unsupported image size");
  c10::cuda::CUDAGuard guard(x.device());
  at::Tensor a = at::empty({N, OH, OW, kStemK}, x.options());
  const int64_t total = N * OH * OW * kStemGranules;
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)sms * 32));
  stem_im2col_kernel<<<grid, 256, 0, at::cuda::getCurrentCUDAStream()>>>(reinterpret_cast<const __nv_bfloat16*>(x.data_ptr()),
                                                                         reinterpret_cast<__nv_bfloat16*>(a.data_ptr()), H, W, OH, OW, total);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return a.permute({0, 3, 1, 2});
}

}  // namespace ptd
