// Batched FP32 GEMM used by the spectral stages:
//   reconstruction  X = V diag(f) V^T          (transA=1, transB=0, kscale=f)
//   eigenbasis in   Gt = V^T G V               (NT then NN)
//   eigenbasis out  Gb = V W V^T               (NN then TN)
//   covariance      S = (X-m)^T (X-m) / M      (transA=1, transB=0, centring, alpha=1/M)
// FP32 FMA accumulation: the parity budget (1e-4 on theta after 15 layers) leaves no room
// for single-pass TF32 (1e-3), so the contraction stays on the FP32 pipe here.
// 64x64x16 tiles, 256 threads, 4x4 register micro-tiles, register-staged double buffering.
#include "kernels.cuh"

namespace uglad {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

__global__ void __launch_bounds__(256) gemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int bz = blockIdx.z;
  const float* __restrict__ A = g.A + (size_t)bz * g.sA;
  const float* __restrict__ Bm = g.Bm + (size_t)bz * g.sB;
  const float* __restrict__ ks = g.kscale ? g.kscale + (size_t)bz * g.sK : nullptr;
  const float* __restrict__ mA = g.meanA ? g.meanA + (size_t)bz * g.sMean : nullptr;
  const float* __restrict__ mB = g.meanB ? g.meanB + (size_t)bz * g.sMean : nullptr;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (g.lower_only && n0 >= m0 + BM) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int M = g.M, N = g.N, K = g.K;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + 256 * r;
      int kk, mm;
      if (g.transA) { kk = idx >> 6; mm = idx & 63; } else { kk = idx & 15; mm = idx >> 4; }
      const int gk = k0 + kk, gm = m0 + mm;
      float v = 0.f;
      if (gk < K && gm < M) {
        v = g.transA ? A[(size_t)gk * g.lda + gm] : A[(size_t)gm * g.lda + gk];
        if (mA) v -= mA[gm];
      }
      ra[r] = v;
      int kb, nn;
      if (g.transB) { kb = idx & 15; nn = idx >> 4; } else { kb = idx >> 6; nn = idx & 63; }
      const int gkb = k0 + kb, gn = n0 + nn;
      float w = 0.f;
      if (gkb < K && gn < N) {
        w = g.transB ? Bm[(size_t)gn * g.ldb + gkb] : Bm[(size_t)gkb * g.ldb + gn];
        if (mB) w -= mB[gn];
        if (ks) w *= ks[gkb];
      }
      rb[r] = w;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + 256 * r;
      int kk, mm;
      if (g.transA) { kk = idx >> 6; mm = idx & 63; } else { kk = idx & 15; mm = idx >> 4; }
      As[kk][mm] = ra[r];
      int kb, nn;
      if (g.transB) { kb = idx & 15; nn = idx >> 4; } else { kb = idx >> 6; nn = idx & 63; }
      Bs[kb][nn] = rb[r];
    }
  };

  fetch(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    stash();
    __syncthreads();
    if (k0 + BK < K) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  float* C = g.C + (size_t)bz * g.sC;
  const float* E1 = g.E1 ? g.E1 + (size_t)bz * g.sE1 : nullptr;
  const float alpha = g.alpha_dev ? g.alpha * g.alpha_dev[(size_t)bz * g.sAlpha] : g.alpha;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = alpha * acc[i][j];
      if (E1) v = fmaf(g.beta, E1[(size_t)gm * g.lde1 + gn], v);
      if (gm == gn) v += g.diag;
      C[(size_t)gm * g.ldc + gn] = v;
    }
  }
}

int launch_gemm(const GemmArgs& g, int batch, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || batch <= 0) return 0;
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, batch);
  if (grid.z > 65535 || grid.y > 65535) {
    set_error("gemm: grid too large (batch=%d)", batch);
    return 1;
  }
  gemm_kernel<<<grid, 256, 0, st>>>(g);
  UGLAD_CHECK_LAUNCH("gemm_kernel");
  return 0;
}

// dispatch point for the dense D^3 products of the large-D path
int launch_gemm_auto(const GemmArgs& g, int batch, cudaStream_t st) { return launch_gemm(g, batch, st); }

}  // namespace uglad
