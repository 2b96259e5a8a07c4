// Blocked Cholesky for the large-D path (D > UGLAD_SMALL_D_MAX), batched over graphs:
//   * log det Theta and Theta^-1 for the glasso loss and its gradient (main.py:306-315),
//   * theta_0 = (S + t I)^-1 (glad.py:115-117),
//   * the positive-definiteness test behind the covariance repair (prepare_data.py:345-355).
// Right-looking, block size CB = 64: the diagonal block is factored (and its triangular inverse
// formed) by one CTA in shared memory, the panel solve and the trailing update are GEMMs.
// A^-1 = W^T W with W = L^-1 built block row by block row from the diagonal inverses.
#include "kernels.cuh"

namespace uglad {

constexpr int CB = 64;
constexpr int CBP = CB + 1;  // padded leading dimension in shared memory

// Factor the nb x nb diagonal block at (j0, j0) of every graph: A_jj = L L^T in place (lower),
// Winv = L^-1 (dense CB x CB, zero above the diagonal and beyond nb), logdet partial
// sum(log L_ii), fail flag when a pivot is not positive.
__global__ void __launch_bounds__(256) chol_diag_kernel(float* A, int D, int j0, int nb, float shift,
                                                        const float* shift_dev, float* Winv, float* logdet_part, int nparts, int jblk,
                                                        int* fail) {
  __shared__ float Ls[CB][CBP];
  __shared__ float Ws[CB][CBP];
  __shared__ float piv_s[CB];
  __shared__ double redd[32];
  __shared__ int s_fail;
  const int b = blockIdx.x, tid = threadIdx.x;
  float* Ab = A + (size_t)b * D * D;
  if (tid == 0) s_fail = 0;
  if (shift_dev) shift += shift_dev[b];
  for (int idx = tid; idx < CB * CB; idx += 256) {
    const int i = idx / CB, j = idx % CB;
    float v = 0.f;
    if (i < nb && j < nb && j <= i) {
      v = Ab[(size_t)(j0 + i) * D + j0 + j];
      if (i == j) v -= shift;
    }
    Ls[i][j] = v;
    Ws[i][j] = 0.f;
  }
  __syncthreads();
  if (tid < CB) piv_s[tid] = 1.f;
  const int tx = tid & 15, ty = tid >> 4;   // 16 x 16 threads over the trailing triangle
  for (int j = 0; j < nb; ++j) {
    const float piv = Ls[j][j];
    __syncthreads();
    if (!(piv > 0.f)) {
      if (tid == 0) s_fail = 1;
      break;  // uniform: every thread read the same pivot
    }
    const float r = sqrtf(piv), ir = 1.f / r;
    if (tid == 0) piv_s[j] = r;
    // scale column j
    for (int i = j + tid; i < nb; i += 256) Ls[i][j] = (i == j) ? r : Ls[i][j] * ir;
    __syncthreads();
    // rank-1 update of the trailing lower triangle
    for (int i = j + 1 + ty; i < nb; i += 16) {
      const float lij = Ls[i][j];
      for (int k = j + 1 + tx; k <= i; k += 16) Ls[i][k] = fmaf(-lij, Ls[k][j], Ls[i][k]);
    }
    __syncthreads();
  }
  __syncthreads();
  const bool bad = s_fail != 0;
  // log det of the block: sum log L_jj, in double, one pivot per thread
  const double ld = block_sum_d((tid < nb && !bad) ? log((double)piv_s[tid]) : 0.0, redd);
  // W = L^-1 by forward substitution, one thread per column c: W[i][c] = (d_ic - sum_k L[i][k] W[k][c]) / L[i][i]
  if (!bad && tid < nb) {
    const int c = tid;
    for (int i = c; i < nb; ++i) {
      float acc = (i == c) ? 1.f : 0.f;
      for (int k = c; k < i; ++k) acc = fmaf(-Ls[i][k], Ws[k][c], acc);
      Ws[i][c] = acc / Ls[i][i];
    }
  }
  __syncthreads();
  for (int idx = tid; idx < CB * CB; idx += 256) {
    const int i = idx / CB, j = idx % CB;
    if (i < nb && j < nb && j <= i) Ab[(size_t)(j0 + i) * D + j0 + j] = Ls[i][j];
    Winv[((size_t)b * CB + i) * CB + j] = bad ? 0.f : Ws[i][j];
  }
  if (tid == 0) {
    logdet_part[(size_t)b * nparts + jblk] = (float)ld;
    if (bad) fail[b] = 1;
  }
}

// logdet[b] = 2 sum_j part[b][j], NaN when the factorisation broke down (torch.logdet of a
// matrix with a negative determinant is NaN; main.py:307)
__global__ void chol_logdet_kernel(const float* part, int nparts, const int* fail, float* logdet, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double s = 0.0;
  for (int j = 0; j < nparts; ++j) s += (double)part[(size_t)b * nparts + j];
  logdet[b] = fail[b] ? __int_as_float(0x7fc00000) : (float)(2.0 * s);
}

// copy the CB x CB diagonal inverses onto the block diagonal of W and zero the rest
__global__ void chol_winit_kernel(const float* Winv, int D, int nblkc, float* W) {
  const int b = blockIdx.y;
  const size_t n = (size_t)D * D;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / D), j = (int)(idx % D);
    float v = 0.f;
    if (i / CB == j / CB) v = Winv[(((size_t)(i / CB) * gridDim.y + b) * CB + (i % CB)) * CB + (j % CB)];
    W[(size_t)b * n + idx] = v;
  }
  (void)nblkc;
}

__global__ void chol_copy_shift_kernel(const float* __restrict__ src, long long sSrc, int D, float shift,
                                       const float* shift_dev, float* __restrict__ dst) {
  const size_t n = (size_t)D * D;
  if (shift_dev) shift += shift_dev[0];
  const float* s = src + (size_t)blockIdx.y * sSrc;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    float v = s[idx];
    if (idx / D == idx % D) v += shift;
    dst[(size_t)blockIdx.y * n + idx] = v;
  }
}

static inline size_t al4c(size_t x) { return (x + 3) & ~(size_t)3; }
static inline int nblocks_c(int D) { return (D + CB - 1) / CB; }

// padded order of the recursive triangular inverse: CB * 2^levels >= D
static inline int padded_order(int D) {
  int dp = CB;
  while (dp < D) dp *= 2;
  return dp;
}

// scratch: Winv [nblk][B][CB][CB] | logdet parts [B][nblk] | fail [B] | Tpanel [B][CB][D] |
//          recursive inverse on the tensor pipe: Lp, W, Wt, Tt, each [B][Dp][Dp]
size_t chol_scratch_floats(int B, int D) {
  const size_t nb = nblocks_c(D);
  const size_t dp = padded_order(D);
  return al4c(nb * B * CB * CB) + al4c((size_t)B * nb) + al4c(B) + al4c((size_t)B * CB * D) + 4 * al4c((size_t)B * dp * dp);
}

struct CholBuf { float* Winv; float* ldpart; int* fail; float* Tp; float* R[4]; };
static CholBuf chol_carve(float* scratch, int B, int D) {
  const size_t nb = nblocks_c(D);
  CholBuf c;
  c.Winv = scratch;
  c.ldpart = c.Winv + al4c(nb * B * CB * CB);
  c.fail = reinterpret_cast<int*>(c.ldpart + al4c((size_t)B * nb));
  c.Tp = reinterpret_cast<float*>(c.fail) + al4c(B);
  const size_t dp = padded_order(D);
  for (int i = 0; i < 4; ++i) c.R[i] = c.Tp + al4c((size_t)B * CB * D) + (size_t)i * al4c((size_t)B * dp * dp);
  return c;
}

// In-place lower Cholesky factor of (A - shift I) for every graph; logdet (may be null).
// fail[b] (device, inside scratch) is 1 where the matrix was not positive definite.
int chol_factor(float* A, int B, int D, float shift, const float* shift_dev, float* logdet, float* scratch,
                cudaStream_t st) {
  const CholBuf c = chol_carve(scratch, B, D);
  const int nb = nblocks_c(D);
  const long long n2 = (long long)D * D;
  UGLAD_CUDA(cudaMemsetAsync(c.fail, 0, (size_t)B * sizeof(int), st));
  for (int j = 0; j < nb; ++j) {
    const int j0 = j * CB, w = (D - j0 < CB) ? D - j0 : CB;
    float* Wj = c.Winv + (size_t)j * B * CB * CB;
    chol_diag_kernel<<<B, 256, 0, st>>>(A, D, j0, w, shift, shift_dev, Wj, c.ldpart, nb, j, c.fail);
    UGLAD_CHECK_LAUNCH("chol_diag_kernel");
    const int rest = D - j0 - w;
    if (rest <= 0) break;
    // panel: L[i, j] = A[i, j] L_jj^-T   (in place: one CTA column covers all of K = N = w)
    GemmArgs p;
    p.A = A + (size_t)(j0 + w) * D + j0; p.lda = D; p.sA = n2;
    p.Bm = Wj; p.ldb = CB; p.sB = (long long)CB * CB; p.transB = 1;
    p.C = A + (size_t)(j0 + w) * D + j0; p.ldc = D; p.sC = n2;
    p.M = rest; p.N = w; p.K = w;
    if (launch_gemm(p, B, st)) return 1;
    // trailing update (lower tiles only): A[i, i'] -= L[i, j] L[i', j]^T
    GemmArgs u;
    u.A = p.C; u.lda = D; u.sA = n2;
    u.Bm = p.C; u.ldb = D; u.sB = n2; u.transB = 1;
    u.C = A + (size_t)(j0 + w) * D + (j0 + w); u.ldc = D; u.sC = n2;
    u.E1 = u.C; u.lde1 = D; u.sE1 = n2; u.beta = 1.f; u.alpha = -1.f;
    u.M = rest; u.N = rest; u.K = w; u.lower_only = 1;
    if (launch_gemm(u, B, st)) return 1;
  }
  if (logdet) {
    chol_logdet_kernel<<<(B + 127) / 128, 128, 0, st>>>(c.ldpart, nb, c.fail, logdet, B);
    UGLAD_CHECK_LAUNCH("chol_logdet_kernel");
  }
  return 0;
}
const int* chol_fail_flags(float* scratch, int B, int D) { return chol_carve(scratch, B, D).fail; }


// ---- triangular inverse by recursive doubling on the tensor pipe -----------------------------------
// inv([L11 0; L21 L22]) = [W11 0; -W22 L21 W11, W22]: starting from the CB x CB diagonal inverses of
// chol_diag_kernel, every level doubles the block size with three tcgen05 products per block pair
// (all pairs of a level in one batched launch when B == 1).  W and its transpose Wt are carried
// together so that every operand is K-major as stored:
//   Tt  = Wt11 L21^T   (= (L21 W11)^T)      A = Wt11, B rows = L21 rows
//   W21 = -W22 T                            A = W22,  B rows = Tt rows
//   Wt12 = -Tt W22^T   (= W21^T)            A = Tt,   B rows = W22 rows
// The matrices are padded to Dp = CB 2^levels (zero rows/columns: they never reach the D x D part).

// Lp = lower triangle of the factor, zero elsewhere and in the padding
__global__ void chol_pad_kernel(const float* __restrict__ Lf, int D, int Dp, float* __restrict__ Lp) {
  const size_t n = (size_t)Dp * Dp;
  const float* L = Lf + (size_t)blockIdx.y * D * D;
  float* out = Lp + (size_t)blockIdx.y * n;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / Dp), j = (int)(idx % Dp);
    out[idx] = (i < D && j <= i) ? L[(size_t)i * D + j] : 0.f;
  }
}
// W, Wt <- block diagonal of the CB x CB inverses (and their transposes), zero elsewhere
__global__ void chol_winit2_kernel(const float* __restrict__ Winv, int nblk, int Dp, float* __restrict__ W,
                                   float* __restrict__ Wt) {
  const int b = blockIdx.y;
  const size_t n = (size_t)Dp * Dp;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / Dp), j = (int)(idx % Dp);
    float v = 0.f, vt = 0.f;
    if (i / CB == j / CB && i / CB < nblk) {
      const float* blk = Winv + ((size_t)(i / CB) * gridDim.y + b) * CB * CB;
      v = blk[(i % CB) * CB + (j % CB)];
      vt = blk[(j % CB) * CB + (i % CB)];
    }
    W[(size_t)b * n + idx] = v;
    Wt[(size_t)b * n + idx] = vt;
  }
}

static TcGemm tri_gemm(const float* A, const float* Bop, float* C, int s, int Dp, long long stride, float alpha) {
  TcGemm g;
  g.A_hi = A; g.B_hi = Bop; g.C_hi = C;
  g.M = g.N = g.K = s;
  g.lda = g.ldb = g.ldc = Dp;
  g.sA = g.sB = g.sC = stride;
  g.alpha = alpha;
  return g;
}

static int chol_inverse_tc(const float* Lf, int B, int D, float* Ainv, float alpha, const float* E1, long long sE1,
                           float beta, const CholBuf& c, cudaStream_t st) {
  const int Dp = padded_order(D), nb = nblocks_c(D);
  const long long n2p = (long long)Dp * Dp;
  float *Lp = c.R[0], *W = c.R[1], *Wt = c.R[2], *Tt = c.R[3];
  dim3 grid(128, B);
  chol_pad_kernel<<<grid, 256, 0, st>>>(Lf, D, Dp, Lp);
  UGLAD_CHECK_LAUNCH("chol_pad_kernel");
  chol_winit2_kernel<<<grid, 256, 0, st>>>(c.Winv, nb, Dp, W, Wt);
  UGLAD_CHECK_LAUNCH("chol_winit2_kernel");
  for (int s = CB; s < Dp; s *= 2) {
    const int npairs = Dp / (2 * s);
    // B == 1: all pairs of the level in one batch (stride 2s rows + 2s columns); else one pair per
    // launch, batched over the graphs
    const int outer = (B == 1) ? 1 : npairs;
    const int batch = (B == 1) ? npairs : B;
    const long long stride = (B == 1) ? 2LL * s * (Dp + 1) : n2p;
    for (int p = 0; p < outer; ++p) {
      const size_t o11 = (size_t)(2 * s * p) * (Dp + 1);          // block (2sp, 2sp)
      const size_t o22 = o11 + (size_t)s * (Dp + 1);               // block (2sp + s, 2sp + s)
      const size_t o21 = o11 + (size_t)s * Dp;                     // block (2sp + s, 2sp)
      const size_t o12 = o11 + (size_t)s;                          // block (2sp, 2sp + s)
      if (launch_tc_gemm(tri_gemm(Wt + o11, Lp + o21, Tt + o12, s, Dp, stride, 1.f), batch, st)) return 1;
      if (launch_tc_gemm2(tri_gemm(W + o22, Tt + o12, W + o21, s, Dp, stride, -1.f),
                          tri_gemm(Tt + o12, W + o22, Wt + o12, s, Dp, stride, -1.f), batch, st)) return 1;
    }
  }
  // Ainv = alpha W^T W + beta E1 = alpha Wt Wt^T + beta E1
  TcGemm g;
  g.A_hi = Wt; g.B_hi = Wt;
  g.M = g.N = g.K = D;
  g.lda = g.ldb = Dp; g.sA = g.sB = n2p;
  g.alpha = alpha; g.beta = beta; g.E1_hi = E1; g.sE1 = sE1; g.lde1 = D;
  g.C_hi = Ainv; g.ldc = D; g.sC = (long long)D * D;
  return launch_tc_gemm(g, B, st);
}

// Ainv = (L L^T)^-1 from the factor left in Lf by chol_factor (same scratch).  W: [B][D][D] work.
int chol_inverse(const float* Lf, int B, int D, float* W, float* Ainv, float alpha, const float* E1,
                 long long sE1, float beta, float* scratch, cudaStream_t st) {
  const CholBuf c = chol_carve(scratch, B, D);
  if (ns_use_tc() && tc_raw_enabled()) return chol_inverse_tc(Lf, B, D, Ainv, alpha, E1, sE1, beta, c, st);
  const int nb = nblocks_c(D);
  const long long n2 = (long long)D * D;
  dim3 grid(64, B);
  chol_winit_kernel<<<grid, 256, 0, st>>>(c.Winv, D, nb, W);
  UGLAD_CHECK_LAUNCH("chol_winit_kernel");
  for (int i = 1; i < nb; ++i) {
    const int i0 = i * CB, w = (D - i0 < CB) ? D - i0 : CB;
    // T = L[i, 0:i0] W[0:i0, 0:i0]
    GemmArgs t;
    t.A = Lf + (size_t)i0 * D; t.lda = D; t.sA = n2;
    t.Bm = W; t.ldb = D; t.sB = n2;
    t.C = c.Tp; t.ldc = D; t.sC = (long long)CB * D;
    t.M = w; t.N = i0; t.K = i0;
    if (launch_gemm(t, B, st)) return 1;
    // W[i, 0:i0] = -Winv_i T
    GemmArgs v;
    v.A = c.Winv + (size_t)i * B * CB * CB; v.lda = CB; v.sA = (long long)CB * CB;
    v.Bm = c.Tp; v.ldb = D; v.sB = (long long)CB * D;
    v.C = W + (size_t)i0 * D; v.ldc = D; v.sC = n2;
    v.M = w; v.N = i0; v.K = w; v.alpha = -1.f;
    if (launch_gemm(v, B, st)) return 1;
  }
  // Ainv = alpha W^T W + beta E1
  GemmArgs g;
  g.A = W; g.Bm = W; g.C = Ainv;
  g.M = g.N = g.K = D;
  g.lda = g.ldb = g.ldc = g.lde1 = D;
  g.sA = g.sB = g.sC = n2;
  g.transA = 1;
  g.alpha = alpha; g.beta = beta; g.E1 = E1; g.sE1 = sE1;
  return launch_gemm_auto(g, B, st);
}

// ---- float64 positive-definiteness test (covariance repair decision, prepare_data.py:345-355) ------
// The reference decides "min eig <= 1e-6" on float64 eigenvalues.  An FP32 factorisation carries a
// backward error of ~sqrt(D) eps32 ||S|| -- above the threshold itself -- so whenever the FP32 test is
// not conclusive (api.cu: condition_large) the decision is taken here: left-looking Cholesky of
// S - shift[b] I in double, one CTA per graph, the factor kept transposed (Lt[k][i] = L[i][k]) so that
// the threads' reads are coalesced.  D^3/3 double FMAs through one SM: ~10 ms at D = 1000; it only runs
// for ill-conditioned covariances, once per fit.  fail[b] = 1 when a pivot is not positive.
__global__ void __launch_bounds__(1024) chol_pd_test_f64_kernel(const float* __restrict__ S, const double* __restrict__ S64, int D,
                                                                const double* __restrict__ shift,
                                                                const int* __restrict__ active, double* __restrict__ work,
                                                                int* __restrict__ fail) {
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  if (active && !active[b]) return;
  __shared__ double s_piv;
  const float* Sb = S ? S + (size_t)b * D * D : nullptr;
  const double* Sd = S64 ? S64 + (size_t)b * D * D : nullptr;
  double* Lt = work + (size_t)b * D * D;
  const double sh = shift[b];
  for (int j = 0; j < D; ++j) {
    // column j of L for the rows i >= j owned by this thread
    for (int i = j + tid; i < D; i += nt) {
      double acc = (Sd ? Sd[(size_t)i * D + j] : (double)Sb[(size_t)i * D + j]) - ((i == j) ? sh : 0.0);
      const double* lj = Lt + j;   // Lt[k][j], stride D
      const double* li = Lt + i;
      int k = 0;
      for (; k + 4 <= j; k += 4) {
        const double a0 = li[(size_t)k * D], a1 = li[(size_t)(k + 1) * D], a2 = li[(size_t)(k + 2) * D], a3 = li[(size_t)(k + 3) * D];
        const double c0 = lj[(size_t)k * D], c1 = lj[(size_t)(k + 1) * D], c2 = lj[(size_t)(k + 2) * D], c3 = lj[(size_t)(k + 3) * D];
        acc -= a0 * c0 + a1 * c1 + a2 * c2 + a3 * c3;
      }
      for (; k < j; ++k) acc -= li[(size_t)k * D] * lj[(size_t)k * D];
      if (i == j) s_piv = acc;
      Lt[(size_t)j * D + i] = acc;   // unscaled for now
    }
    __syncthreads();
    const double piv = s_piv;
    if (!(piv > 0.0)) {
      if (tid == 0) fail[b] = 1;
      return;   // uniform
    }
    const double ir = 1.0 / sqrt(piv);
    for (int i = j + tid; i < D; i += nt) Lt[(size_t)j * D + i] *= ir;
    __syncthreads();
  }
  if (tid == 0) fail[b] = 0;
}
int chol_pd_test_f64(const float* S, const double* S64, int B, int D, const double* shift_dev, const int* active_dev,
                     double* work, int* fail_dev, cudaStream_t st) {
  int threads = ((D + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  if (threads < 64) threads = 64;
  chol_pd_test_f64_kernel<<<B, threads, 0, st>>>(S, S64, D, shift_dev, active_dev, work, fail_dev);
  UGLAD_CHECK_LAUNCH("chol_pd_test_f64_kernel");
  return 0;
}
// float64 covariance of the samples for the graphs marked active (the matrix the reference decides on):
// S64[b] = (X_b - mean_b)^T (X_b - mean_b) / M.  16 x 16 output tile per block, 16 samples per step.
__global__ void __launch_bounds__(256) cov_f64_kernel(const float* __restrict__ X, const float* __restrict__ mean, int M, int D,
                                                      const int* __restrict__ active, double* __restrict__ S64) {
  const int b = blockIdx.z;
  if (active && !active[b]) return;
  if (blockIdx.x > blockIdx.y) return;   // lower tiles; mirrored below
  __shared__ double ta[16][17], tb[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
  const float* Xb = X + (size_t)b * M * D;
  const float* mb = mean + (size_t)b * D;
  const double mi = (i0 + tx < D) ? (double)mb[i0 + tx] : 0.0, mj = (j0 + tx < D) ? (double)mb[j0 + tx] : 0.0;
  double acc = 0.0;
  for (int m0 = 0; m0 < M; m0 += 16) {
    const int m = m0 + ty;
    ta[ty][tx] = (m < M && i0 + tx < D) ? (double)Xb[(size_t)m * D + i0 + tx] - mi : 0.0;
    tb[ty][tx] = (m < M && j0 + tx < D) ? (double)Xb[(size_t)m * D + j0 + tx] - mj : 0.0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += ta[k][ty] * tb[k][tx];
    __syncthreads();
  }
  const int i = i0 + ty, j = j0 + tx;
  if (i < D && j < D) {
    const double v = acc / (double)M;
    S64[(size_t)b * D * D + (size_t)i * D + j] = v;
    S64[(size_t)b * D * D + (size_t)j * D + i] = v;
  }
}
int launch_cov_f64(const float* X, const float* mean, int B, int M, int D, const int* active_dev, double* S64,
                   cudaStream_t st) {
  dim3 grid((D + 15) / 16, (D + 15) / 16, B);
  cov_f64_kernel<<<grid, 256, 0, st>>>(X, mean, M, D, active_dev, S64);
  UGLAD_CHECK_LAUNCH("cov_f64_kernel");
  return 0;
}
// trace of every graph (the scale of the FP32 factorisation's backward error)
__global__ void trace_kernel(const float* __restrict__ S, int D, float* __restrict__ out) {
  __shared__ float red[32];
  const float* Sb = S + (size_t)blockIdx.x * D * D;
  float t = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) t += Sb[(size_t)i * D + i];
  t = block_sum(t, red);
  if (threadIdx.x == 0) out[blockIdx.x] = t;
}
int launch_trace(const float* S, int B, int D, float* out, cudaStream_t st) {
  trace_kernel<<<B, 256, 0, st>>>(S, D, out);
  UGLAD_CHECK_LAUNCH("trace_kernel");
  return 0;
}

int launch_copy_shift(const float* src, long long sSrc, int B, int D, float shift, const float* shift_dev,
                      float* dst, cudaStream_t st) {
  dim3 grid(64, B);
  chol_copy_shift_kernel<<<grid, 256, 0, st>>>(src, sSrc, D, shift, shift_dev, dst);
  UGLAD_CHECK_LAUNCH("chol_copy_shift_kernel");
  return 0;
}

// S_ii += add[b]  (prepare_data.py:349-350)
__global__ void add_diag_kernel(float* S, int D, const float* add) {
  const float a = add[blockIdx.y];
  float* Sb = S + (size_t)blockIdx.y * D * D;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D; i += gridDim.x * blockDim.x) Sb[(size_t)i * D + i] += a;
}
int launch_add_diag(float* S, int B, int D, const float* add_dev, cudaStream_t st) {
  dim3 grid((D + 255) / 256, B);
  add_diag_kernel<<<grid, 256, 0, st>>>(S, D, add_dev);
  UGLAD_CHECK_LAUNCH("add_diag_kernel");
  return 0;
}

}  // namespace uglad
