// Shared helpers for the uglad_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/uglad_b200.h"

#define UGLAD_NS_ITERS 10  // torch_sqrtm.py:13,33 (itr_TH)

namespace uglad {

void set_error(const char* fmt, ...);
void count_launch();  // every kernel launch of the library is counted (uglad_launch_count)

#define UGLAD_CHECK_LAUNCH(name)                                                      \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      uglad::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));       \
      return 1;                                                                       \
    }                                                                                 \
    uglad::count_launch();                                                            \
  } while (0)

#define UGLAD_CUDA(call)                                                              \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      uglad::set_error("%s failed: %s", #call, cudaGetErrorString(e__));              \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

// Packed parameter vector layout (GladParams.state_dict() order).
struct ParamLayout {
  int H;
  int t0, rW1, rb1, rW2, rb2, rW3, rb3, lW1, lb1, lW2, lb2, total;
};
__host__ __device__ inline ParamLayout param_layout(int H) {
  ParamLayout p;
  p.H = H;
  int o = 0;
  p.t0 = o;  o += 1;
  p.rW1 = o; o += H * UGLAD_NF;
  p.rb1 = o; o += H;
  p.rW2 = o; o += H * H;
  p.rb2 = o; o += H;
  p.rW3 = o; o += H;
  p.rb3 = o; o += 1;
  p.lW1 = o; o += H * 2;
  p.lb1 = o; o += H;
  p.lW2 = o; o += H;
  p.lb2 = o; o += 1;
  p.total = o;
  return p;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; `red` is shared scratch of >= 32 floats; result valid in every thread.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ double block_sum_d(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum_d(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double r = (lane < nw) ? red[lane] : 0.0;
  r = warp_sum_d(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
// MUFU-based activations of the per-entry rho_l1 MLP (six tanh and one sigmoid per matrix entry: the
// libm versions made z_update instruction-bound).  Absolute error ~1e-7 (ex2.approx + rcp.approx;
// tanh(x) = sign(x) (1 - 2 / (e^{2|x|} + 1)) saturates cleanly), two orders below the parity budget.
__device__ __forceinline__ float fast_tanh(float x) {
  const float e = __expf(2.f * fabsf(x));
  const float t = 1.f - __fdividef(2.f, e + 1.f);
  return copysignf(t, x);
}
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// ---------------------------------------------------------------------------------------
// Newton-Schulz iteration collapsed onto eigenvalues (torch_sqrtm.py:12-45).
// Forward: Y_t = V diag(y_t) V^T, so the matrix recurrence is this scalar one per eigenvalue.
__device__ __forceinline__ float ns_forward_scalar(float mu, float nrm) {
  float y = mu / nrm, z = 1.f;
#pragma unroll
  for (int t = 0; t < UGLAD_NS_ITERS; ++t) {
    const float T = 0.5f * (3.f - z * y);
    y = y * T;
    z = T * z;
  }
  return y * sqrtf(nrm);
}
// Backward: entry (i,j) of Q (in the eigenbasis of the saved root) is scaled every step by
// 0.5*(3 - a_i^2 - a_j^2 + a_i a_j) while a <- 0.5 a (3 - a^2).  Returns that product.
__device__ __forceinline__ float ns_backward_factor(float ai, float aj) {
  float c = 1.f;
#pragma unroll
  for (int t = 0; t < UGLAD_NS_ITERS; ++t) {
    c *= 0.5f * (3.f - ai * ai - aj * aj + ai * aj);
    ai = 0.5f * ai * (3.f - ai * ai);
    aj = 0.5f * aj * (3.f - aj * aj);
  }
  return c;
}

// ---------------------------------------------------------------------------------------
// rho_l1 MLP (glad_params.py:33-44) evaluated per matrix entry, weights in shared memory.
// HT > 0: compile-time hidden width (fully unrolled); HT == 0: runtime width <= UGLAD_MAX_H.
template <int HT>
struct RhoMLP {
  static constexpr int HM = HT ? HT : UGLAD_MAX_H;
  const float* w;  // shared-memory copy of the packed parameter vector
  ParamLayout pl;
  int H;
  __device__ RhoMLP(const float* w_, int h) : w(w_), pl(param_layout(HT ? HT : h)), H(HT ? HT : h) {}

  __device__ __forceinline__ float forward(float x, float s, float f, float* h1, float* h2) const {
#pragma unroll
    for (int i = 0; i < HM; ++i) {
      if (i < H) {
        const float* r = w + pl.rW1 + i * UGLAD_NF;
        h1[i] = fast_tanh(fmaf(r[0], x, fmaf(r[1], s, fmaf(r[2], f, w[pl.rb1 + i]))));
      }
    }
#pragma unroll
    for (int i = 0; i < HM; ++i) {
      if (i < H) {
        float a = w[pl.rb2 + i];
#pragma unroll
        for (int j = 0; j < HM; ++j)
          if (j < H) a = fmaf(w[pl.rW2 + i * H + j], h1[j], a);
        h2[i] = fast_tanh(a);
      }
    }
    float o = w[pl.rb3];
#pragma unroll
    for (int j = 0; j < HM; ++j)
      if (j < H) o = fmaf(w[pl.rW3 + j], h2[j], o);
    return fast_sigmoid(o);
  }
};

// soft threshold  sign(x) * max(0, |x| - rho)   (glad_params.py:81), with torch's NaN behaviour: torch.max and torch.sign
// both propagate NaN (fmaxf would launder it into 0 and a diverged fit would sail past the reference's NaN stop)
__device__ __forceinline__ float soft_threshold(float x, float rho) {
  const float d = fabsf(x) - rho;
  const float m = (d > 0.f) ? d : ((d != d) ? d : 0.f);
  return (x > 0.f) ? m : ((x < 0.f) ? -m : x * m);
}

}  // namespace uglad
