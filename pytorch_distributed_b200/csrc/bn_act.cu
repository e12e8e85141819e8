// Fused NHWC BatchNorm (+ residual add) (+ ReLU), forward and backward, for sm_100a.
//
// The reference reaches cuDNN BN + ATen add + ATen ReLU through torchvision's ResNet
// (/root/reference/distributed.py:136-139,250).  On B200 a bf16 ResNet-50 step is bound by exactly those
// memory passes, so they are fused here:
//   forward : stats pass (1 read)  + apply pass  (x [+res] -> y : 1-2 reads, 1 write)       eager: 3-5 R, 2-3 W
//   backward: reduce pass (dy,y,x) + apply pass  (dy,y,x -> dx [,dres])                      eager: 6 R, 2-3 W
// Layout: activations are channels_last, i.e. a row-major [M = N*H*W, C] matrix.  A thread owns 8 consecutive
// channels (one 16-byte vector for 16-bit dtypes) and walks down the rows, so every warp access is a fully
// coalesced 128..512-byte line and the per-channel reductions stay in registers until the end of the CTA.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "common.cuh"
#include "host.h"

namespace ptd {

constexpr int kBnThreads = 256;

struct RowMap {
  int cgs;     // channel groups (C / 8)
  int tpr;     // threads per row
  int rpp;     // rows per pass of one CTA
  int rlocal;  // this thread's row inside a pass
  int cg0;     // this thread's first channel group
  bool active;
};
__device__ __forceinline__ RowMap row_map(int C) {
  RowMap m;
  m.cgs = C >> 3;
  m.tpr = min(m.cgs, (int)blockDim.x);
  m.rpp = blockDim.x / m.tpr;
  m.rlocal = threadIdx.x / m.tpr;
  m.cg0 = threadIdx.x % m.tpr;
  m.active = m.rlocal < m.rpp;
  return m;
}

__device__ __forceinline__ float ld_w(const void* p, int dt, int i) {
  switch (dt) {
    case kF32: return reinterpret_cast<const float*>(p)[i];
    case kBF16: return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
    default: return __half2float(reinterpret_cast<const __half*>(p)[i]);
  }
}
__device__ __forceinline__ void st_w(void* p, int dt, int i, float v) {
  switch (dt) {
    case kF32: reinterpret_cast<float*>(p)[i] = v; break;
    case kBF16: reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v); break;
    default: reinterpret_cast<__half*>(p)[i] = __float2half_rn(v); break;
  }
}

// Flush per-thread channel partials: lanes of a warp that share a channel group are combined by shuffles,
// then one shared-memory atomic per (warp, channel), then one global atomic per (CTA, channel).
__device__ __forceinline__ void flush_partials(const RowMap& m, int cg, float (&a)[8], float (&b)[8], float* sm, int C) {
  if (m.tpr < 32 && (m.tpr & (m.tpr - 1)) == 0) {
    for (int o = m.tpr; o < 32; o <<= 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
        b[k] += __shfl_xor_sync(0xffffffffu, b[k], o);
      }
    }
    if ((threadIdx.x & 31) >= m.tpr) return;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    atomicAdd(&sm[cg * 8 + k], a[k]);
    atomicAdd(&sm[C + cg * 8 + k], b[k]);
  }
}

// ------------------------------------------------------------------ forward: statistics
template <typename T>
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const T* __restrict__ x, float* __restrict__ gsum, int64_t M, int C,
                                                              int rows_per_block) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const RowMap m = row_map(C);
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(M, r0 + (int64_t)rows_per_block);
  const bool warp_uniform = (m.tpr >= 32) || ((m.tpr & (m.tpr - 1)) == 0);
  if (m.active || warp_uniform) {
    for (int cg = m.cg0; cg < m.cgs; cg += m.tpr) {
      float s[8], q[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { s[k] = 0.f; q[k] = 0.f; }
      if (m.active) {
        const T* p = x + cg * 8;
        int64_t r = r0 + m.rlocal;
        for (; r + 3 * (int64_t)m.rpp < r1; r += 4 * (int64_t)m.rpp) {
          float f[4][8];
#pragma unroll
          for (int u = 0; u < 4; ++u) load8<T>(p + (r + (int64_t)u * m.rpp) * C, f[u]);
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) { s[k] += f[u][k]; q[k] += f[u][k] * f[u][k]; }
        }
        for (; r < r1; r += m.rpp) {
          float f[8];
          load8<T>(p + r * C, f);
#pragma unroll
          for (int k = 0; k < 8; ++k) { s[k] += f[k]; q[k] += f[k] * f[k]; }
        }
      }
      flush_partials(m, cg, s, q, sm, C);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&gsum[i], sm[i]);
}

// ------------------------------------------------------------------ forward: normalise (+res) (+relu)
template <typename T, bool RELU, bool RES>
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y,
                                                              const float* __restrict__ gsum, const void* __restrict__ w,
                                                              const void* __restrict__ b, int wdt, float* __restrict__ running_mean,
                                                              float* __restrict__ running_var, float* __restrict__ saved, int64_t M, int C,
                                                              float eps, float momentum, int training) {
  const RowMap m = row_map(C);
  const float inv_m = 1.f / (float)M;
  if (!m.active) return;
  for (int cg = m.cg0; cg < m.cgs; cg += m.tpr) {
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = cg * 8 + k;
      float mean, var;
      if (training) {
        mean = gsum[c] * inv_m;
        var = fmaxf(gsum[C + c] * inv_m - mean * mean, 0.f);
      } else {
        mean = running_mean[c];
        var = running_var[c];
      }
      const float invstd = rsqrtf(var + eps);
      sc[k] = ld_w(w, wdt, c) * invstd;
      sh[k] = ld_w(b, wdt, c) - mean * sc[k];
      if (training && blockIdx.x == 0 && m.rlocal == 0) {
        saved[c] = mean;
        saved[C + c] = invstd;
        if (running_mean) {
          const float unbiased = M > 1 ? var * ((float)M / (float)(M - 1)) : var;
          running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
          running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
        }
      }
    }
    const int64_t stride = (int64_t)gridDim.x * m.rpp;
    const int64_t coff = cg * 8;
    int64_t r = (int64_t)blockIdx.x * m.rpp + m.rlocal;
    for (; r + 3 * stride < M; r += 4 * stride) {
      float f[4][8], g[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) load8<T>(x + (r + u * stride) * C + coff, f[u]);
      if constexpr (RES) {
#pragma unroll
        for (int u = 0; u < 4; ++u) load8<T>(res + (r + u * stride) * C + coff, g[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float v = f[u][k] * sc[k] + sh[k];
          if constexpr (RES) v += g[u][k];
          if constexpr (RELU) v = fmaxf(v, 0.f);
          f[u][k] = v;
        }
        store8<T>(y + (r + u * stride) * C + coff, f[u]);
      }
    }
    for (; r < M; r += stride) {
      float f[8], g[8];
      load8<T>(x + r * C + coff, f);
      if constexpr (RES) load8<T>(res + r * C + coff, g);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v = f[k] * sc[k] + sh[k];
        if constexpr (RES) v += g[k];
        if constexpr (RELU) v = fmaxf(v, 0.f);
        f[k] = v;
      }
      store8<T>(y + r * C + coff, f);
    }
  }
}

// ------------------------------------------------------------------ backward: reductions
// gsum[0:C] = sum dz ; gsum[C:2C] = sum dz * xhat   with dz = dy * (y > 0) when RELU
template <typename T, bool RELU>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ y, const T* __restrict__ x,
                                                                   const float* __restrict__ saved, float* __restrict__ gsum, int64_t M, int C,
                                                                   int rows_per_block) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const RowMap m = row_map(C);
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(M, r0 + (int64_t)rows_per_block);
  const bool warp_uniform = (m.tpr >= 32) || ((m.tpr & (m.tpr - 1)) == 0);
  if (m.active || warp_uniform) {
    for (int cg = m.cg0; cg < m.cgs; cg += m.tpr) {
      float s[8], q[8], mean[8], invstd[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { s[k] = 0.f; q[k] = 0.f; mean[k] = saved[cg * 8 + k]; invstd[k] = saved[C + cg * 8 + k]; }
      if (m.active) {
        const int64_t coff = cg * 8;
        int64_t r = r0 + m.rlocal;
        for (; r + (int64_t)m.rpp < r1; r += 2 * (int64_t)m.rpp) {
          float d[2][8], o[2][8], v[2][8];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int64_t off = (r + (int64_t)u * m.rpp) * C + coff;
            load8<T>(dy + off, d[u]);
            load8<T>(x + off, v[u]);
            if constexpr (RELU) load8<T>(y + off, o[u]);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float dz = d[u][k];
              if constexpr (RELU) dz = o[u][k] > 0.f ? dz : 0.f;
              s[k] += dz;
              q[k] += dz * (v[u][k] - mean[k]) * invstd[k];
            }
        }
        for (; r < r1; r += m.rpp) {
          float d[8], o[8], v[8];
          const int64_t off = r * C + coff;
          load8<T>(dy + off, d);
          load8<T>(x + off, v);
          if constexpr (RELU) load8<T>(y + off, o);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float dz = d[k];
            if constexpr (RELU) dz = o[k] > 0.f ? dz : 0.f;
            s[k] += dz;
            q[k] += dz * (v[k] - mean[k]) * invstd[k];
          }
        }
      }
      flush_partials(m, cg, s, q, sm, C);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&gsum[i], sm[i]);
}

// ------------------------------------------------------------------ backward: apply
// dx = gamma*invstd * (dz - mean(dz) - xhat * mean(dz*xhat)) ; dres = dz ; dgamma = sum dz*xhat ; dbeta = sum dz
template <typename T, bool RELU, bool RES>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ y, const T* __restrict__ x,
                                                                  const float* __restrict__ saved, const float* __restrict__ gsum,
                                                                  const void* __restrict__ w, int wdt, T* __restrict__ dx, T* __restrict__ dres,
                                                                  void* __restrict__ dw, void* __restrict__ db, int64_t M, int C) {
  const RowMap m = row_map(C);
  if (!m.active) return;
  const float inv_m = 1.f / (float)M;
  for (int cg = m.cg0; cg < m.cgs; cg += m.tpr) {
    float mean[8], invstd[8], k1[8], k2[8], sc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = cg * 8 + k;
      mean[k] = saved[c];
      invstd[k] = saved[C + c];
      const float sdz = gsum[c], sdzx = gsum[C + c];
      sc[k] = ld_w(w, wdt, c) * invstd[k];
      k1[k] = sdz * inv_m;
      k2[k] = sdzx * inv_m;
      if (blockIdx.x == 0 && m.rlocal == 0) {
        st_w(dw, wdt, c, sdzx);
        st_w(db, wdt, c, sdz);
      }
    }
    const int64_t stride = (int64_t)gridDim.x * m.rpp;
    const int64_t coff = cg * 8;
    int64_t r = (int64_t)blockIdx.x * m.rpp + m.rlocal;
    for (; r + stride < M; r += 2 * stride) {
      float d[2][8], o[2][8], v[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t off = (r + u * stride) * C + coff;
        load8<T>(dy + off, d[u]);
        load8<T>(x + off, v[u]);
        if constexpr (RELU) load8<T>(y + off, o[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t off = (r + u * stride) * C + coff;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float dz = d[u][k];
          if constexpr (RELU) dz = o[u][k] > 0.f ? dz : 0.f;
          d[u][k] = dz;
          const float xhat = (v[u][k] - mean[k]) * invstd[k];
          v[u][k] = sc[k] * (dz - k1[k] - xhat * k2[k]);
        }
        store8<T>(dx + off, v[u]);
        if constexpr (RES) store8<T>(dres + off, d[u]);
      }
    }
    for (; r < M; r += stride) {
      float d[8], o[8], v[8];
      const int64_t off = r * C + coff;
      load8<T>(dy + off, d);
      load8<T>(x + off, v);
      if constexpr (RELU) load8<T>(y + off, o);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float dz = d[k];
        if constexpr (RELU) dz = o[k] > 0.f ? dz : 0.f;
        d[k] = dz;
        const float xhat = (v[k] - mean[k]) * invstd[k];
        v[k] = sc[k] * (dz - k1[k] - xhat * k2[k]);
      }
      store8<T>(dx + off, v);
      if constexpr (RES) store8<T>(dres + off, d);
    }
  }
}

// ------------------------------------------------------------------ host side
static int wdtype(const at::Tensor& t) {
  switch (t.scalar_type()) {
    case at::kFloat: return kF32;
    case at::kBFloat16: return kBF16;
    case at::kHalf: return kF16;
    default: TORCH_CHECK(false, "unsupported BN parameter dtype"); return 0;
  }
}

static void check_nhwc(const at::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.dim() == 4, name, " must be a 4-d CUDA tensor");
  TORCH_CHECK(t.is_contiguous(at::MemoryFormat::ChannelsLast), name, " must be channels_last contiguous");
}

struct Geometry { int64_t M; int C; int rpp; int sms; };
static Geometry geometry(const at::Tensor& x) {
  Geometry g;
  g.C = (int)x.size(1);
  g.M = x.numel() / g.C;
  TORCH_CHECK(g.C % 8 == 0 && g.C <= 8192, "fused BN needs C % 8 == 0 and C <= 8192 (got ", g.C, ")");
  const int tpr = std::min(g.C / 8, kBnThreads);
  g.rpp = kBnThreads / tpr;
  g.sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  return g;
}
static int reduce_grid(const Geometry& g, int* rows_per_block) {
  int64_t blocks = (g.M + (int64_t)g.rpp * 16 - 1) / ((int64_t)g.rpp * 16);
  blocks = std::max<int64_t>(std::min<int64_t>(blocks, (int64_t)g.sms * 8), 1);
  int64_t rpb = (g.M + blocks - 1) / blocks;
  rpb = (rpb + g.rpp - 1) / g.rpp * g.rpp;
  *rows_per_block = (int)rpb;
  return (int)((g.M + rpb - 1) / rpb);
}
static int apply_grid(const Geometry& g) {
  int64_t blocks = (g.M + (int64_t)g.rpp * 4 - 1) / ((int64_t)g.rpp * 4);
  return (int)std::max<int64_t>(std::min<int64_t>(blocks, (int64_t)g.sms * 8), 1);
}

// returns {y, saved(mean|invstd)} ; `work` = zeroed float[2C] accumulator supplied by the caller
template <typename T>
static void fwd_impl(const at::Tensor& x, const at::Tensor* res, at::Tensor& y, at::Tensor& work, at::Tensor& saved, const at::Tensor& w,
                     const at::Tensor& b, at::Tensor& rm, at::Tensor& rv, bool training, float momentum, float eps, bool relu) {
  const Geometry g = geometry(x);
  cudaStream_t st = at::cuda::getCurrentCUDAStream();
  const T* xp = reinterpret_cast<const T*>(x.data_ptr());
  T* yp = reinterpret_cast<T*>(y.data_ptr());
  const T* rp = res ? reinterpret_cast<const T*>(res->data_ptr()) : nullptr;
  float* wk = work.data_ptr<float>();
  if (training) {
    int rpb;
    const int grid = reduce_grid(g, &rpb);
    bn_stats_kernel<T><<<grid, kBnThreads, 2 * g.C * sizeof(float), st>>>(xp, wk, g.M, g.C, rpb);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
  }
  const int grid = apply_grid(g);
  float* rmp = rm.defined() ? rm.data_ptr<float>() : nullptr;
  float* rvp = rv.defined() ? rv.data_ptr<float>() : nullptr;
  float* sv = saved.defined() ? saved.data_ptr<float>() : nullptr;
  const int wdt = wdtype(w);
#define APPLY(R, S) \
  bn_apply_kernel<T, R, S><<<grid, kBnThreads, 0, st>>>(xp, rp, yp, wk, w.data_ptr(), b.data_ptr(), wdt, rmp, rvp, sv, g.M, g.C, eps, momentum, training ? 1 : 0)
  if (relu) { if (res) APPLY(true, true); else APPLY(true, false); }
  else      { if (res) APPLY(false, true); else APPLY(false, false); }
#undef APPLY
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

std::vector<at::Tensor> bn_act_forward(const at::Tensor& x, const c10::optional<at::Tensor>& residual, const at::Tensor& weight,
                                       const at::Tensor& bias, at::Tensor running_mean, at::Tensor running_var, bool training, double momentum,
                                       double eps, bool relu, at::Tensor work) {
  check_nhwc(x, "x");
  TORCH_CHECK(weight.scalar_type() == bias.scalar_type() && weight.is_contiguous() && bias.is_contiguous());
  if (running_mean.defined()) TORCH_CHECK(running_mean.scalar_type() == at::kFloat && running_var.scalar_type() == at::kFloat, "running stats must be fp32");
  TORCH_CHECK(training || running_mean.defined(), "eval mode needs running statistics");
  const at::Tensor* res = nullptr;
  if (residual.has_value() && residual->defined()) {
    check_nhwc(*residual, "residual");
    TORCH_CHECK(residual->sizes() == x.sizes() && residual->scalar_type() == x.scalar_type(), "residual must match x");
    res = &*residual;
  }
  c10::cuda::CUDAGuard guard(x.device());
  const int C = (int)x.size(1);
  at::Tensor y = at::empty_like(x, x.options().memory_format(at::MemoryFormat::ChannelsLast));
  at::Tensor saved;
  if (training) {
    TORCH_CHECK(work.defined() && work.scalar_type() == at::kFloat && work.numel() >= 2 * C, "work buffer too small");
    saved = at::empty({2 * C}, x.options().dtype(at::kFloat));
  } else {
    work = running_mean;  // unused pointer
  }
  switch (x.scalar_type()) {
    case at::kBFloat16: fwd_impl<__nv_bfloat16>(x, res, y, work, saved, weight, bias, running_mean, running_var, training, (float)momentum, (float)eps, relu); break;
    case at::kHalf: fwd_impl<__half>(x, res, y, work, saved, weight, bias, running_mean, running_var, training, (float)momentum, (float)eps, relu); break;
    case at::kFloat: fwd_impl<float>(x, res, y, work, saved, weight, bias, running_mean, running_var, training, (float)momentum, (float)eps, relu); break;
    default: TORCH_CHECK(false, "unsupported activation dtype");
  }
  return {y, saved};
}

template <typename T>
static void bwd_impl(const at::Tensor& dy, const at::Tensor& y, const at::Tensor& x, const at::Tensor& saved, at::Tensor& work, const at::Tensor& w,
                     at::Tensor& dx, at::Tensor& dres, at::Tensor& dw, at::Tensor& db, bool relu, bool has_res) {
  const Geometry g = geometry(x);
  cudaStream_t st = at::cuda::getCurrentCUDAStream();
  const T* dyp = reinterpret_cast<const T*>(dy.data_ptr());
  const T* yp = relu ? reinterpret_cast<const T*>(y.data_ptr()) : nullptr;
  const T* xp = reinterpret_cast<const T*>(x.data_ptr());
  float* wk = work.data_ptr<float>();
  const float* sv = saved.data_ptr<float>();
  int rpb;
  const int rgrid = reduce_grid(g, &rpb);
  if (relu) bn_bwd_reduce_kernel<T, true><<<rgrid, kBnThreads, 2 * g.C * sizeof(float), st>>>(dyp, yp, xp, sv, wk, g.M, g.C, rpb);
  else      bn_bwd_reduce_kernel<T, false><<<rgrid, kBnThreads, 2 * g.C * sizeof(float), st>>>(dyp, yp, xp, sv, wk, g.M, g.C, rpb);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  const int grid = apply_grid(g);
  T* dxp = reinterpret_cast<T*>(dx.data_ptr());
  T* drp = has_res ? reinterpret_cast<T*>(dres.data_ptr()) : nullptr;
  const int wdt = wdtype(w);
#define BAPPLY(R, S) \
  bn_bwd_apply_kernel<T, R, S><<<grid, kBnThreads, 0, st>>>(dyp, yp, xp, sv, wk, w.data_ptr(), wdt, dxp, drp, dw.data_ptr(), db.data_ptr(), g.M, g.C)
  if (relu) { if (has_res) BAPPLY(true, true); else BAPPLY(true, false); }
  else      { if (has_res) BAPPLY(false, true); else BAPPLY(false, false); }
#undef BAPPLY
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// returns {dx, dres (undefined if !has_residual), dweight, dbias}
std::vector<at::Tensor> bn_act_backward(const at::Tensor& dy_in, const at::Tensor& x, const at::Tensor& y, const at::Tensor& weight,
                                        const at::Tensor& saved, bool relu, bool has_residual, at::Tensor work) {
  check_nhwc(x, "x");
  at::Tensor dy = dy_in.is_contiguous(at::MemoryFormat::ChannelsLast) ? dy_in : dy_in.contiguous(at::MemoryFormat::ChannelsLast);
  TORCH_CHECK(dy.scalar_type() == x.scalar_type() && dy.sizes() == x.sizes());
  if (relu) check_nhwc(y, "y");
  const int C = (int)x.size(1);
  TORCH_CHECK(work.defined() && work.scalar_type() == at::kFloat && work.numel() >= 2 * C, "work buffer too small");
  c10::cuda::CUDAGuard guard(x.device());
  at::Tensor dx = at::empty_like(x, x.options().memory_format(at::MemoryFormat::ChannelsLast));
  at::Tensor dres;
  if (has_residual) dres = (!relu) ? dy : at::empty_like(x, x.options().memory_format(at::MemoryFormat::ChannelsLast));
  at::Tensor dw = at::empty_like(weight), db = at::empty_like(weight);
  const bool write_res = has_residual && relu;  // without ReLU the residual gradient IS dy: no copy
  switch (x.scalar_type()) {
    case at::kBFloat16: bwd_impl<__nv_bfloat16>(dy, y, x, saved, work, weight, dx, dres, dw, db, relu, write_res); break;
    case at::kHalf: bwd_impl<__half>(dy, y, x, saved, work, weight, dx, dres, dw, db, relu, write_res); break;
    case at::kFloat: bwd_impl<float>(dy, y, x, saved, work, weight, dx, dres, dw, db, relu, write_res); break;
    default: TORCH_CHECK(false, "unsupported activation dtype");
  }
  return {dx, dres, dw, db};
}

}  // namespace ptd
