// Large-D (D > UGLAD_SMALL_D_MAX) theta update: the reference's Newton-Schulz square root
// (torch_sqrtm.py:12-45) kept in its matrix form, because at D ~ 1000 a chain of dense D^3
// products on the tensor pipe is far cheaper than any eigensolver ("large D").
//
// forward  (glad.py:139-142, torch_sqrtm.py:12-28), per graph:
//   b = S/lam - Theta ;  A = b b + 4/lam I ;  n = ||A||_F ;  Y0 = A/n, Z0 = I
//   10x { T = (3I - Z Y)/2 ; Y <- Y T ; Z <- T Z }         (28 products: Z0 = I and the last Z
//   X = (sqrt(n) Y - b)/2                                    are never formed)
// backward (torch_sqrtm.py:31-45): R = 2X + b is the saved root, r = ||R||_F, A = R/r,
//   Q = (GX/2)/r ; 10x { Q <- (Q(3I - AA) - A^T(A^T Q - Q A))/2 ; A <- A(3I - AA)/2 } (59 products)
//   H = Q/2 = dL/d(b b + 4/lam I) ;  Gb = b (H + H^T) - GX/2 ;  tr(H) feeds dL/dlam.
// Every product goes through mm(): the tcgen05 3xTF32 kernel when the shape allows it,
// the FP32 SIMT kernel otherwise.
#include <string.h>
#include "kernels.cuh"

namespace uglad {

constexpr int NS_THREADS = 256;

// per-graph reduction finish: the last block of graph b (ticket counter) sums that graph's
// partials in index order in double and publishes {sqrt(sum), 1/sqrt(sum), sqrt(sqrt(sum))}
__device__ __forceinline__ void finish_norm(float tot, float* part, unsigned* counter, float* scal, int B,
                                            double* redd, bool* s_last) {
  const int b = blockIdx.y, nb = gridDim.x;
  if (threadIdx.x == 0) {
    part[(size_t)b * nb + blockIdx.x] = tot;
    __threadfence();
    *s_last = (atomicAdd(counter + b, 1u) == (unsigned)nb - 1u);
  }
  __syncthreads();
  if (*s_last) {
    __threadfence();
    double s = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s += (double)((volatile float*)part)[(size_t)b * nb + i];
    s = block_sum_d(s, redd);
    if (threadIdx.x == 0) {
      const double n = sqrt(s);
      scal[b] = (float)n;
      scal[B + b] = (float)(1.0 / n);
      scal[2 * B + b] = (float)sqrt(n);
      counter[b] = 0u;
    }
  }
}

// b = S/lam - Theta
__global__ void __launch_bounds__(NS_THREADS) ns_build_b_kernel(const float* __restrict__ S, long long sS,
                                                               const float* __restrict__ Theta,
                                                               const float* __restrict__ lam, int n,
                                                               float* __restrict__ bout) {
  const float il = 1.f / lam[0];
  const float* Sb = S + (size_t)blockIdx.y * sS;
  const size_t base = (size_t)blockIdx.y * n;
  for (int i = blockIdx.x * NS_THREADS + threadIdx.x; i < n; i += gridDim.x * NS_THREADS)
    bout[base + i] = fmaf(il, Sb[i], -Theta[base + i]);
}

// A_ii += 4/lam ; scal <- ||A||_F
__global__ void __launch_bounds__(NS_THREADS) ns_diag_fro_kernel(float* __restrict__ A, const float* __restrict__ lam,
                                                                int D, float* part, unsigned* counter,
                                                                float* scal, int B) {
  __shared__ float red[32];
  __shared__ double redd[32];
  __shared__ bool s_last;
  const float c = 4.f / lam[0];
  const int n = D * D;
  const size_t base = (size_t)blockIdx.y * n;
  float acc = 0.f;
  for (int i = blockIdx.x * NS_THREADS + threadIdx.x; i < n; i += gridDim.x * NS_THREADS) {
    float v = A[base + i];
    if (i / D == i % D) {
      v += c;
      A[base + i] = v;
    }
    acc = fmaf(v, v, acc);
  }
  const float tot = block_sum(acc, red);
  finish_norm(tot, part, counter, scal, B, redd, &s_last);
}

// Z1 = T0 = (3I - A/n)/2
__global__ void __launch_bounds__(NS_THREADS) ns_t0_kernel(const float* __restrict__ A, const float* __restrict__ scal,
                                                          int B, int D, float* __restrict__ Z) {
  const int n = D * D;
  const size_t base = (size_t)blockIdx.y * n;
  const float h = -0.5f * scal[B + blockIdx.y];
  for (int i = blockIdx.x * NS_THREADS + threadIdx.x; i < n; i += gridDim.x * NS_THREADS)
    Z[base + i] = fmaf(h, A[base + i], (i / D == i % D) ? 1.5f : 0.f);
}

// backward prologue: b = S/lam - Theta ; R = 2X + b ; scal <- ||R||_F
__global__ void __launch_bounds__(NS_THREADS) ns_build_r_kernel(const float* __restrict__ S, long long sS,
                                                               const float* __restrict__ Theta,
                                                               const float* __restrict__ X,
                                                               const float* __restrict__ lam, int n,
                                                               float* __restrict__ bout, float* __restrict__ R,
                                                               float* part, unsigned* counter, float* scal, int B) {
  __shared__ float red[32];
  __shared__ double redd[32];
  __shared__ bool s_last;
  const float il = 1.f / lam[0];
  const float* Sb = S + (size_t)blockIdx.y * sS;
  const size_t base = (size_t)blockIdx.y * n;
  float acc = 0.f;
  for (int i = blockIdx.x * NS_THREADS + threadIdx.x; i < n; i += gridDim.x * NS_THREADS) {
    const float bv = fmaf(il, Sb[i], -Theta[base + i]);
    const float r = fmaf(2.f, X[base + i], bv);
    bout[base + i] = bv;
    R[base + i] = r;
    acc = fmaf(r, r, acc);
  }
  const float tot = block_sum(acc, red);
  finish_norm(tot, part, counter, scal, B, redd, &s_last);
}

// A = R/r (in place) ; Q = (GX/2)/r
__global__ void __launch_bounds__(NS_THREADS) ns_scale_aq_kernel(float* __restrict__ A, const float* __restrict__ GX,
                                                                const float* __restrict__ scal, int B, int n,
                                                                float* __restrict__ Q) {
  const size_t base = (size_t)blockIdx.y * n;
  const float inv = scal[B + blockIdx.y];
  for (int i = blockIdx.x * NS_THREADS + threadIdx.x; i < n; i += gridDim.x * NS_THREADS) {
    A[base + i] *= inv;
    Q[base + i] = GX[base + i] * (0.5f * inv);
  }
}

// Hs = H + H^T = (Q + Q^T)/2 with H = Q/2 ; per-block partial of tr(H).  32x32 tiles, block (32, 8).
__global__ void ns_hsym_kernel(const float* __restrict__ Q, int D, float* __restrict__ Hs, float* trh_part,
                               int part_stride) {
  __shared__ float t[32][33];
  __shared__ float red[32];
  const size_t base = (size_t)blockIdx.z * D * D;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int gi = bx + r, gj = by + threadIdx.x;  // transposed tile: Q[bx + r][by + c]
    t[r][threadIdx.x] = (gi < D && gj < D) ? Q[base + (size_t)gi * D + gj] : 0.f;
  }
  __syncthreads();
  float tr = 0.f;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int gi = by + r, gj = bx + threadIdx.x;
    if (gi < D && gj < D) {
      const float q = Q[base + (size_t)gi * D + gj];
      Hs[base + (size_t)gi * D + gj] = 0.5f * (q + t[threadIdx.x][r]);
      if (gi == gj) tr += 0.5f * q;
    }
  }
  // block reduction of the trace partial (only diagonal tiles contribute)
  const int tid = threadIdx.y * 32 + threadIdx.x;
  tr = warp_sum(tr);
  if ((tid & 31) == 0) red[tid >> 5] = tr;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    if (blockIdx.x == blockIdx.y) trh_part[(size_t)blockIdx.z * part_stride + blockIdx.x] = s;
  }
}

__global__ void ns_zero_kernel(float* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.f;
}

// ---------------------------------------------------------------------------------------------
struct NsBuf {
  float* M[8];
  float* scal;        // [3][B]
  float* part;        // [B][nblk]
  unsigned* counter;  // [B]
};
static inline size_t al4(size_t x) { return (x + 3) & ~(size_t)3; }

static int g_use_tc = 1;
bool ns_use_tc() { return g_use_tc != 0; }
int ns_tune(const char* key, int value) {
  if (!strcmp(key, "use_tc")) { g_use_tc = value ? 1 : 0; return 0; }
  if (!strcmp(key, "tc_bn")) return tc_tune_bn(value);
  if (!strcmp(key, "tc_pdl")) return tc_tune_pdl(value);
  if (!strcmp(key, "tc_dual")) return tc_tune_dual(value);
  if (!strcmp(key, "tc_raw")) return tc_tune_raw(value);
  if (!strcmp(key, "tc_tma_store")) return tc_tune_tma_store(value);
  if (!strcmp(key, "tc_exp")) return tc_tune_exp(value);
  if (!strcmp(key, "tc_atm")) return tc_tune_atm(value);
  if (!strcmp(key, "tc_chain")) return tc_tune_chain(value);
  if (!strcmp(key, "tc_chain_bn")) return tc_tune_chain_bn(value);
  return 1;
}

static size_t ns_simt_scratch_floats(int B, int D) {
  const size_t n2 = (size_t)B * D * D;
  return 8 * al4(n2) + al4(3 * (size_t)B) + al4((size_t)B * elem_blocks_per_graph(D)) + al4(B);
}
// one region serves both product back-ends (the "use_tc" knob may flip between calls)
size_t ns_scratch_floats(int B, int D) {
  const size_t a = ns_simt_scratch_floats(B, D), b = ns_tc_scratch_floats(B, D);
  return a > b ? a : b;
}
static NsBuf ns_carve(float* scratch, int B, int D) {
  NsBuf s;
  const size_t n2 = al4((size_t)B * D * D);
  for (int i = 0; i < 8; ++i) s.M[i] = scratch + i * n2;
  s.scal = scratch + 8 * n2;
  s.part = s.scal + al4(3 * (size_t)B);
  s.counter = reinterpret_cast<unsigned*>(s.part + al4((size_t)B * elem_blocks_per_graph(D)));
  return s;
}
int ns_scratch_init(float* scratch, int B, int D, cudaStream_t st) {
  NsBuf s = ns_carve(scratch, B, D);
  UGLAD_CUDA(cudaMemsetAsync(s.counter, 0, (size_t)B * sizeof(unsigned), st));
  return ns_tc_scratch_init(scratch, B, D, st);
}

struct MM {
  const float* A; int tA;
  const float* Bm; int tB;
  float* C;
  float alpha = 1.f; const float* alpha_dev = nullptr;
  float beta = 0.f; const float* E1 = nullptr;
  float diag = 0.f;
};
static int mm(const MM& m, int B, int D, cudaStream_t st) {
  GemmArgs g;
  g.A = m.A; g.Bm = m.Bm; g.C = m.C;
  g.M = g.N = g.K = D;
  g.lda = g.ldb = g.ldc = g.lde1 = D;
  g.sA = g.sB = g.sC = g.sE1 = (long long)D * D;
  g.transA = m.tA; g.transB = m.tB;
  g.alpha = m.alpha; g.alpha_dev = m.alpha_dev; g.sAlpha = 1;
  g.beta = m.beta; g.E1 = m.E1; g.diag = m.diag;
  return launch_gemm_auto(g, B, st);
}

int ns_theta_update_forward(const float* S, long long sS, const float* Theta, const float* lam, int B, int D,
                            float* X, float* scratch, cudaStream_t st) {
  if (g_use_tc) return ns_tc_theta_update_forward(S, sS, Theta, lam, B, D, X, scratch, st);
  const NsBuf s = ns_carve(scratch, B, D);
  const int n = D * D;
  const dim3 grid(elem_blocks_per_graph(D), B);
  float *b = s.M[0], *A = s.M[1], *Z = s.M[2], *T = s.M[3], *Y = s.M[4], *Y2 = s.M[5], *Z2 = s.M[6];
  ns_build_b_kernel<<<grid, NS_THREADS, 0, st>>>(S, sS, Theta, lam, n, b);
  UGLAD_CHECK_LAUNCH("ns_build_b_kernel");
  { MM m{b, 1, b, 0, A}; if (mm(m, B, D, st)) return 1; }  // b^T b (glad.py:140)
  ns_diag_fro_kernel<<<grid, NS_THREADS, 0, st>>>(A, lam, D, s.part, s.counter, s.scal, B);
  UGLAD_CHECK_LAUNCH("ns_diag_fro_kernel");
  ns_t0_kernel<<<grid, NS_THREADS, 0, st>>>(A, s.scal, B, D, Z);
  UGLAD_CHECK_LAUNCH("ns_t0_kernel");
  { MM m{A, 0, Z, 0, Y}; m.alpha_dev = s.scal + B; if (mm(m, B, D, st)) return 1; }  // Y1 = Y0 T0
  for (int t = 1; t < UGLAD_NS_ITERS; ++t) {
    { MM m{Z, 0, Y, 0, T}; m.alpha = -0.5f; m.diag = 1.5f; if (mm(m, B, D, st)) return 1; }
    if (t + 1 < UGLAD_NS_ITERS) {
      { MM m{Y, 0, T, 0, Y2}; if (mm(m, B, D, st)) return 1; }
      { MM m{T, 0, Z, 0, Z2}; if (mm(m, B, D, st)) return 1; }
      float* tmp = Y; Y = Y2; Y2 = tmp;
      tmp = Z; Z = Z2; Z2 = tmp;
    } else {  // X = (sqrt(n) Y T - b)/2
      MM m{Y, 0, T, 0, X};
      m.alpha = 0.5f; m.alpha_dev = s.scal + 2 * B; m.beta = -0.5f; m.E1 = b;
      if (mm(m, B, D, st)) return 1;
    }
  }
  return 0;
}

int ns_theta_update_backward(const float* S, long long sS, const float* Theta, const float* X, const float* lam,
                             const float* GX, int B, int D, float* Gb, float* trh_part, int nblk,
                             float* scratch, cudaStream_t st) {
  if (g_use_tc) return ns_tc_theta_update_backward(S, sS, Theta, X, lam, GX, B, D, Gb, trh_part, nblk, scratch, st);
  const NsBuf s = ns_carve(scratch, B, D);
  const int n = D * D;
  const dim3 grid(elem_blocks_per_graph(D), B);
  float *b = s.M[0], *A = s.M[1], *A2 = s.M[2], *Q = s.M[3], *Q2 = s.M[4], *B3 = s.M[5], *QB = s.M[6], *W = s.M[7];
  ns_build_r_kernel<<<grid, NS_THREADS, 0, st>>>(S, sS, Theta, X, lam, n, b, A, s.part, s.counter, s.scal, B);
  UGLAD_CHECK_LAUNCH("ns_build_r_kernel");
  ns_scale_aq_kernel<<<grid, NS_THREADS, 0, st>>>(A, GX, s.scal, B, n, Q);
  UGLAD_CHECK_LAUNCH("ns_scale_aq_kernel");
  for (int t = 0; t < UGLAD_NS_ITERS; ++t) {
    { MM m{A, 0, A, 0, B3}; m.alpha = -1.f; m.diag = 3.f; if (mm(m, B, D, st)) return 1; }        // 3I - AA
    { MM m{Q, 0, B3, 0, QB}; if (mm(m, B, D, st)) return 1; }                                       // Q(3I - AA)
    { MM m{Q, 0, A, 0, W}; if (mm(m, B, D, st)) return 1; }                                         // QA
    { MM m{A, 1, Q, 0, W}; m.beta = -1.f; m.E1 = W; if (mm(m, B, D, st)) return 1; }                // A^T Q - QA
    { MM m{A, 1, W, 0, Q2}; m.alpha = -0.5f; m.beta = 0.5f; m.E1 = QB; if (mm(m, B, D, st)) return 1; }
    if (t + 1 < UGLAD_NS_ITERS) {
      MM m{A, 0, B3, 0, A2}; m.alpha = 0.5f;
      if (mm(m, B, D, st)) return 1;
      float* tmp = A; A = A2; A2 = tmp;
    }
    float* tmp = Q; Q = Q2; Q2 = tmp;
  }
  // H = Q/2 ; Hs = H + H^T ; tr(H) partials into the first ceil(D/32) slots of this layer's row
  {
    ns_zero_kernel<<<(B * nblk + 255) / 256, 256, 0, st>>>(trh_part, B * nblk);
    UGLAD_CHECK_LAUNCH("ns_zero_kernel");
    const int nt = (D + 31) / 32;
    if (nt > nblk) { set_error("ns backward: %d trace partials do not fit %d slots", nt, nblk); return 1; }
    dim3 g2(nt, nt, B), blk(32, 8);
    ns_hsym_kernel<<<g2, blk, 0, st>>>(Q, D, B3, trh_part, nblk);
    UGLAD_CHECK_LAUNCH("ns_hsym_kernel");
  }
  MM m{b, 0, B3, 0, Gb};
  m.beta = -0.5f; m.E1 = GX;
  return mm(m, B, D, st);
}

}  // namespace uglad
