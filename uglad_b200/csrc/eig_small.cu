// One-CTA-per-graph symmetric eigensolver for D <= UGLAD_SMALL_D_MAX (sm_100a).
//
// Algorithm: one-sided (Hestenes) Jacobi on U = A + sigma*I held column-major in shared
// memory.  With sigma chosen so that A + sigma*I is positive definite (indefinite inputs
// such as b = S/lambda - theta of glad.py:139), orthogonalising the columns of U gives
// U -> V diag(lambda + sigma): eigenvalues are column norms minus sigma, eigenvectors the
// normalised columns, so ONE D x D matrix in shared memory is the whole state (D <= 232
// fits the 227 KB of a B200 SM).
//
// Mapping: a group of LP lanes (8/16/32) owns one column pair per step of a round-robin
// tournament, each lane holding CH float4 chunks of both columns in registers, so a round
// is (nearly always) a single pass; the only reduction per pair is the dot product
// (column norms are cached and updated analytically, refreshed every sweep); the rotation
// is computed with MUFU rsqrt/rcp + one Newton step and applied in the small-angle-accurate
// form a' = a - s(b + tau a), b' = b + s(a - tau b), whose rounding error scales with the
// rotation, not with the column (late sweeps add almost no error).  One __syncthreads per
// round.
//
// sigma: 1.35 x a 10-step power-iteration estimate of ||A||_2; the result is validated
// by sum(norms) == trace + D sigma and redone with the guaranteed min(Gershgorin,
// Frobenius) bound if that fails.  A tight sigma matters: eigenvalue error ~ eps (|A|+sigma).
//
// When two D x D buffers fit (D <= 166) the unshifted-then-shifted input G is kept beside U:
//   * warm start: U_0 = G V_prev (V_prev = the same layer's eigenvectors from the previous
//     epoch, whose b differs by one Adam step) is already nearly orthogonal, so ~2 sweeps
//     replace ~9;
//   * eigenvalues are Rayleigh quotients v^T G v of the final vectors (second-order
//     accurate) instead of accumulated column norms.
//
// Tails:
//   TAIL_LAYER: f_k = (s_k - beta_k)/2 with s_k the reference's 10-step Newton-Schulz
//               square root of beta_k^2 + 4/lambda collapsed onto the eigenvalues
//               (torch_sqrtm.py:12-28, glad.py:140-142);
//   TAIL_LOSS : logdet and -1/eig for main.py:307 (torch.logdet) and its gradient.
#include <string.h>
#include "common.cuh"
#include "kernels.cuh"

namespace uglad {

__device__ __forceinline__ void rr_pair(int n, int r, int i, int& p, int& q) {
  const int m = n - 1;  // circle method: player m fixed, the others rotate
  if (i == 0) {
    p = m;
    q = r;
  } else {
    p = r + i;
    if (p >= m) p -= m;
    q = r - i;
    if (q < 0) q += m;
  }
}

template <int LP>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// same, but only the LP lanes of this group take part (divergent trip counts across groups)
template <int LP>
__device__ __forceinline__ float group_sum_masked(float v, unsigned mask) {
#pragma unroll
  for (int o = LP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

// MUFU approximations without the denormal fix-up sequences of the library wrappers: the rotation
// angle only steers the convergence (Jacobi corrects any O(ulp) error in the next sweep)
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// packed FP32 pairs (sm_100 FFMA2: one issue slot, two FMAs): the sweeps are issue-bound, and the dot
// products and rotations are nothing but FMAs on float4 chunks that already sit in aligned register pairs
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
}

// t = G u for one column u of U (both in shared memory); lane `gl` of the group keeps the
// float4 chunks gl, gl+LP, ... of t in registers.
template <int LP, int CH>
__device__ __forceinline__ void group_matvec(const float* __restrict__ G, const float* __restrict__ u,
                                             int D, int ld, int nch, int gl, float4* tv) {
#pragma unroll
  for (int c = 0; c < CH; ++c) tv[c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
  for (int j = 0; j < D; ++j) {
    const float vj = u[j];
    const float* gj = G + (size_t)j * ld;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int ch = gl + LP * c;
      if (ch < nch) {
        const float4 g4 = *reinterpret_cast<const float4*>(gj + 4 * ch);
        tv[c].x = fmaf(g4.x, vj, tv[c].x);
        tv[c].y = fmaf(g4.y, vj, tv[c].y);
        tv[c].z = fmaf(g4.z, vj, tv[c].z);
        tv[c].w = fmaf(g4.w, vj, tv[c].w);
      }
    }
  }
}

// two columns at once: the G chunks are loaded once for both (half the shared-memory traffic per FMA)
template <int LP, int CH>
__device__ __forceinline__ void group_matvec2(const float* __restrict__ G, const float* __restrict__ u0,
                                              const float* __restrict__ u1, int D, int ld, int nch, int gl,
                                              float4* t0, float4* t1) {
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    t0[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    t1[c] = t0[c];
  }
#pragma unroll 4
  for (int j = 0; j < D; ++j) {
    const float v0 = u0[j], v1 = u1[j];
    const float* gj = G + (size_t)j * ld;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int ch = gl + LP * c;
      if (ch < nch) {
        const float4 g4 = *reinterpret_cast<const float4*>(gj + 4 * ch);
        t0[c].x = fmaf(g4.x, v0, t0[c].x);  t1[c].x = fmaf(g4.x, v1, t1[c].x);
        t0[c].y = fmaf(g4.y, v0, t0[c].y);  t1[c].y = fmaf(g4.y, v1, t1[c].y);
        t0[c].z = fmaf(g4.z, v0, t0[c].z);  t1[c].z = fmaf(g4.z, v1, t1[c].z);
        t0[c].w = fmaf(g4.w, v0, t0[c].w);  t1[c].w = fmaf(g4.w, v1, t1[c].w);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// G * [8 columns] on the warp-level tensor path (mma.sync m16n8k8 TF32, 3xTF32 split in registers):
// the D^3 products of the solver (warm start U0 = G V_prev, Rayleigh quotients) are the only
// GEMM-shaped work of this kernel; as SIMT matvecs they were shared-memory-bandwidth bound.
//   A[m][k] = G[m][k] from shared memory (G is symmetric, so the column-major copy serves),
//   B[k][n] = Bsrc[(c0 + n) * ldb + k]   (shared or global memory), n = 0..7,
//   acc[rt][0..3] = rows 16*(tile0 + rt) + {g, g, g + 8, g + 8}, columns c0 + {2t, 2t + 1, 2t, 2t + 1}.
__device__ __forceinline__ void split2(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
constexpr int MMA_RT = 4;  // row tiles (of 16) accumulated per pass: 16 accumulator registers
__device__ __forceinline__ void mma_g_times_cols(const float* __restrict__ G, int ld, int D,
                                                 const float* __restrict__ Bsrc, int ldb, int c0, int tile0,
                                                 int lane, float (*acc)[4]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int rt = 0; rt < MMA_RT; ++rt)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[rt][e] = 0.f;
  const int col = c0 + g;
  const bool colok = col < D;
  const float* bp = Bsrc + (size_t)(colok ? col : 0) * ldb;
  for (int k0 = 0; k0 < D; k0 += 8) {
    const int ka = k0 + t, kb = ka + 4;
    const float bx0 = (colok && ka < D) ? bp[ka] : 0.f;
    const float bx1 = (colok && kb < D) ? bp[kb] : 0.f;
    uint32_t bh0, bl0, bh1, bl1;
    split2(bx0, bh0, bl0);
    split2(bx1, bh1, bl1);
#pragma unroll
    for (int rt = 0; rt < MMA_RT; ++rt) {
      const int m0 = (tile0 + rt) * 16 + g, m1 = m0 + 8;
      if ((tile0 + rt) * 16 >= D) break;   // warp-uniform
      const float ax0 = (m0 < D && ka < D) ? G[(size_t)m0 * ld + ka] : 0.f;
      const float ax1 = (m1 < D && ka < D) ? G[(size_t)m1 * ld + ka] : 0.f;
      const float ax2 = (m0 < D && kb < D) ? G[(size_t)m0 * ld + kb] : 0.f;
      const float ax3 = (m1 < D && kb < D) ? G[(size_t)m1 * ld + kb] : 0.f;
      uint32_t ah0, al0, ah1, al1, ah2, al2, ah3, al3;
      split2(ax0, ah0, al0); split2(ax1, ah1, al1); split2(ax2, ah2, al2); split2(ax3, ah3, al3);
      mma_tf32(acc[rt], al0, al1, al2, al3, bh0, bh1);
      mma_tf32(acc[rt], ah0, ah1, ah2, ah3, bl0, bl1);
      mma_tf32(acc[rt], ah0, ah1, ah2, ah3, bh0, bh1);
    }
  }
}

constexpr int eig_max_threads(int LP, int CH) {
  const int want = (116 * LP + 31) / 32 * 32, cap = (CH >= 8) ? 512 : 1024;  // CH=8 needs >64 registers
  return want > cap ? cap : want;
}

// PAD: the launcher padded the shared-memory columns to ld = LP * CH * 4 floats (zero rows beyond
// D), so every lane's CH chunks exist and the chunk loops carry no predicates.
template <int LP, int CH, bool PAD>
__global__ void __launch_bounds__(eig_max_threads(LP, CH), 1) eig_jacobi_small_kernel(EigArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int D = a.D, ld = a.ld;
  float* U = smem;                                // [D][ld] column-major
  float* Gk = a.keepG ? U + (size_t)ld * D : U;   // [D][ld] A + sigma I (kept) or alias of U
  float* wv = Gk + (size_t)ld * D;                // [ld] eigenvalues / power-iteration x
  float* nrm2 = wv + ld;                          // [ld] cached squared column norms / y
  double* redd = reinterpret_cast<double*>(nrm2 + ld);  // [32]
  float* red = reinterpret_cast<float*>(redd + 32);     // [32]
  __shared__ unsigned s_flag;
  __shared__ float s_sigma;
  constexpr int MAXFIX = 64;              // column pairs the post-sweep check may hand to the fix-up pass
  __shared__ unsigned s_nfix;
  __shared__ unsigned s_fix[MAXFIX];
  __shared__ unsigned s_fixs[MAXFIX];     // the list in ascending (p, q) order: the atomic slots are not reproducible
  __shared__ int s_stat[3];               // developer knob "eig_timing" = 2: fix-up rotations, full-sweep fallbacks, re-checks
  __shared__ unsigned s_rot, s_dots;      // profiling (a.work != nullptr): applied rotations, column-pair dot products

  const int b = blockIdx.x;
  // retry pass behind the cluster kernel (eig_cluster.cu): only the graphs it flagged (info[0] >= 1000: the warm
  // start was not positive definite) are solved, from scratch with the guaranteed shift
  if (a.retry_only && !(a.info[(size_t)b * 4] >= 1000.f)) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  const int nch = ld >> 2;
  const int grp = tid / LP, gl = tid % LP, ngroups = nthreads / LP;
  const int n = D + (D & 1), npairs = n >> 1, m = n - 1;
  const float tol = a.tol;

  const size_t base = (size_t)b * D * D;
  float inv_lam = 0.f;
  const float* Sb = nullptr;
  const float* Tb = nullptr;
  const float* Ab = nullptr;
  if (a.build) {
    inv_lam = 1.0f / a.lam[0];
    Sb = a.S + (size_t)b * a.strideS;
    Tb = a.Theta + base;
  } else {
    Ab = a.A + base;
  }

  int sweeps = 0;
  if (tid < 3) s_stat[tid] = 0;
  if (tid == 0) { s_rot = 0u; s_dots = 0u; }

  float sigma = 0.f, trace = 0.f, wsum = 0.f;
  long long t_start = clock64(), t_sweeps0 = 0, t_sweeps1 = 0;
  for (int attempt = a.retry_only ? 1 : 0; attempt < 2; ++attempt) {
    if (attempt == 0 && a.U0 != nullptr) {
      // pre-multiplied warm start: U_0 = (A + sigma I) V_prev was formed by a tcgen05 GEMM; nothing else to set up
      const float* U0b = a.U0 + (size_t)b * D * a.ldu;
      for (int idx = tid; idx < D * ld; idx += nthreads) {
        const int col = idx / ld, row = idx - col * ld;
        U[idx] = (row < D) ? U0b[(size_t)col * a.ldu + row] : 0.f;
      }
      if (tid == 0) { s_flag = 0u; s_nfix = 0u; }
      sigma = a.pre_sigma[b];
      trace = a.pre_trace[b];
      __syncthreads();
    } else {
    // ---- load (coalesced along the contiguous global dimension) -------------------------
    float trace_part = 0.f, fro_part = 0.f;
    for (int idx = tid; idx < D * ld; idx += nthreads) {
      const int col = idx / ld, row = idx - col * ld;
      float v = 0.f;
      if (row < D) {
        const size_t g = (size_t)col * D + row;
        v = a.build ? (inv_lam * Sb[g] - Tb[g]) : Ab[g];
        fro_part = fmaf(v, v, fro_part);
        if (row == col) trace_part += v;
      }
      Gk[idx] = v;
    }
    if (tid == 0) { s_flag = 0u; s_nfix = 0u; }
    trace = block_sum(trace_part, red);
    if (a.shift_mode == 1) {
      float bound;
      if (attempt == 0 && a.warm_w != nullptr && a.warmVt != nullptr) {
        // warm start: the previous epoch's eigenvalues of the same layer ARE its spectral norm
        // (the matrix moved by one optimiser step, far inside the 35 % margin)
        float mx = 0.f;
        for (int i = tid; i < D; i += nthreads) mx = fmaxf(mx, fabsf(a.warm_w[(size_t)b * D + i]));
        bound = 1.35f * block_max(mx, red);
      } else if (attempt == 0) {
        // power iteration on the symmetric U: ||A||_2 estimate (a lower bound, hence x1.35)
        for (int i = tid; i < ld; i += nthreads) {
          const unsigned h = (unsigned)(i + 1) * 2654435761u;
          wv[i] = (i < D) ? (((h >> 8) & 0xffff) * (1.f / 32768.f) - 1.f) : 0.f;
        }
        __syncthreads();
        float est = 0.f;
        for (int it = 0; it < 10; ++it) {
          float y = 0.f;
          if (tid < D) {
#pragma unroll 4
            for (int j = 0; j < D; ++j) y = fmaf(Gk[(size_t)j * ld + tid], wv[j], y);
          }
          const float xn = block_sum((tid < D) ? wv[tid] * wv[tid] : 0.f, red);
          const float yn = block_sum(y * y, red);
          est = sqrtf(yn / fmaxf(xn, 1e-30f));
          __syncthreads();
          if (tid < D) wv[tid] = y * rsqrtf(fmaxf(yn, 1e-30f));
          __syncthreads();
        }
        bound = 1.35f * est;
      } else {
        const float fro = sqrtf(block_sum(fro_part, red));
        float gmax = 0.f;
        for (int col = warp; col < D; col += nwarps) {
          float s = 0.f;
          for (int r = lane; r < D; r += 32) s += fabsf(Gk[(size_t)col * ld + r]);
          s = warp_sum(s);
          gmax = fmaxf(gmax, s);
        }
        gmax = block_max(gmax, red);
        bound = 1.25f * fminf(gmax, fro);
      }
      if (tid == 0) s_sigma = (bound > 0.f) ? bound : 1.0f;
    } else if (tid == 0) {
      s_sigma = 0.f;
    }
    __syncthreads();
    sigma = s_sigma;
    if (sigma != 0.f)
      for (int i = tid; i < D; i += nthreads) Gk[(size_t)i * ld + i] += sigma;
    __syncthreads();
    if (a.keepG) {
      if (a.warmVt != nullptr && attempt == 0) {
        // U_0 = G V_prev, column by column in place: U[:,k] = V_prev[k][:] then U[:,k] <- G U[:,k]
        const float* Vp = a.warmVt + base;
        if (a.use_mma) {
          // tensor path: B fragments straight from global V_prev, the product lands in U
          for (int idx = tid; idx < D * ld; idx += nthreads) {   // zero the padding rows of U
            const int row = idx % ld;
            if (row >= D) U[idx] = 0.f;
          }
          const int ntile = (D + 15) / 16;
          for (int cg = warp; cg * 8 < D; cg += nwarps) {
            for (int tile0 = 0; tile0 < ntile; tile0 += MMA_RT) {
              float acc[MMA_RT][4];
              mma_g_times_cols(Gk, ld, D, Vp, D, cg * 8, tile0, lane, acc);
              const int g = lane >> 2, t = lane & 3;
#pragma unroll
              for (int rt = 0; rt < MMA_RT; ++rt) {
                const int m0 = (tile0 + rt) * 16 + g, m1 = m0 + 8;
                const int cA = cg * 8 + 2 * t, cB = cA + 1;
                if (cA < D) {
                  if (m0 < D) U[(size_t)cA * ld + m0] = acc[rt][0];
                  if (m1 < D) U[(size_t)cA * ld + m1] = acc[rt][2];
                }
                if (cB < D) {
                  if (m0 < D) U[(size_t)cB * ld + m0] = acc[rt][1];
                  if (m1 < D) U[(size_t)cB * ld + m1] = acc[rt][3];
                }
              }
            }
          }
        } else {
        for (int idx = tid; idx < D * ld; idx += nthreads) {
          const int col = idx / ld, row = idx - col * ld;
          U[idx] = (row < D) ? Vp[(size_t)col * D + row] : 0.f;
        }
        __syncthreads();
        if constexpr (CH <= 4) {
          for (int cb = 0; cb < D; cb += 2 * ngroups) {
            const int c0 = cb + 2 * grp, c1 = c0 + 1;
            float4 t0[CH], t1[CH];
            float* u0 = U + (size_t)(c0 < D ? c0 : 0) * ld;
            float* u1 = U + (size_t)(c1 < D ? c1 : 0) * ld;
            group_matvec2<LP, CH>(Gk, u0, u1, D, ld, nch, gl, t0, t1);
            __syncwarp();
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              const int ch = gl + LP * c;
              if (ch < nch) {
                if (c0 < D) *reinterpret_cast<float4*>(u0 + 4 * ch) = t0[c];
                if (c1 < D) *reinterpret_cast<float4*>(u1 + 4 * ch) = t1[c];
              }
            }
          }
        } else {
          for (int cb = 0; cb < D; cb += ngroups) {
            const int col = cb + grp;
            float4 tv[CH];
            float* u = U + (size_t)(col < D ? col : 0) * ld;
            group_matvec<LP, CH>(Gk, u, D, ld, nch, gl, tv);
            __syncwarp();
            if (col < D) {
#pragma unroll
              for (int c = 0; c < CH; ++c) {
                const int ch = gl + LP * c;
                if (ch < nch) *reinterpret_cast<float4*>(u + 4 * ch) = tv[c];
              }
            }
          }
        }
        }  // !use_mma
      } else {
        for (int idx = tid; idx < D * ld; idx += nthreads) U[idx] = Gk[idx];
      }
      __syncthreads();
    }
    }  // !pre

    // ---- Jacobi sweeps -------------------------------------------------------------------
    t_sweeps0 = clock64();
    for (int sweep = 0; sweep < a.max_sweeps; ++sweep) {
      // refresh the cached squared norms
      for (int cb = 0; cb < D; cb += ngroups) {
        const int col = cb + grp;
        float s = 0.f;
        if (col < D) {
          const float* u = U + (size_t)col * ld;
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const int ch = gl + LP * c;
            if (ch < nch) {
              const float4 v = *reinterpret_cast<const float4*>(u + 4 * ch);
              s = dot4(v, v, s);
            }
          }
        }
        s = group_sum<LP>(s);
        if (col < D && gl == 0) nrm2[col] = s;
      }
      __syncthreads();
      float wmax = 0.f;
      const float tol2 = tol * tol;
      float* const Ul = U + 4 * gl;
      // one column pair: dot product, convergence test, rotation.  `valid` pairs only are written.
      auto do_pair = [&](int p, int q, bool valid) {
        float* up = Ul + (size_t)p * ld;
        float* uq = Ul + (size_t)q * ld;
        float4 av[CH], bv[CH];
        unsigned long long g01 = 0ull, g23 = 0ull;   // four independent chains in two packed accumulators
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          if (PAD || gl + LP * c < nch) {
            av[c] = *reinterpret_cast<const float4*>(up + 4 * LP * c);
            bv[c] = *reinterpret_cast<const float4*>(uq + 4 * LP * c);
            g01 = ffma2(pk2(av[c].x, av[c].y), pk2(bv[c].x, bv[c].y), g01);
            g23 = ffma2(pk2(av[c].z, av[c].w), pk2(bv[c].z, bv[c].w), g23);
          } else {
            av[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            bv[c] = av[c];
          }
        }
        float g0, g1, g2, g3;
        upk2(g01, g0, g1);
        upk2(g23, g2, g3);
        const float ga = group_sum<LP>((g0 + g1) + (g2 + g3));
        const float al = nrm2[p], be = nrm2[q];
        const float den = al * be;
        const float g2s = ga * ga;
        // |cos(u_p, u_q)| > tol  <=>  ga^2 > tol^2 |u_p|^2 |u_q|^2  (no rsqrt on the common path)
        if (valid && g2s > tol2 * den) {
          wmax = 1.f;   // only "did anything rotate" is consumed
          const float d = be - al, g2 = 2.f * ga;
          const float h = fast_sqrt(fmaf(d, d, g2 * g2));
          float t = fabsf(g2) * fast_rcp(fabsf(d) + h);
          t = ((d < 0.f) != (g2 < 0.f)) ? -t : t;
          const float x = fmaf(t, t, 1.f);
          float cs = fast_rsqrt(x);
          cs = cs * fmaf(-0.5f * x, cs * cs, 1.5f);  // one Newton step: ~0.5 ulp
          const float sn = t * cs;
          const float tau = sn * fast_rcp(1.f + cs);
          const unsigned long long tau2 = pk2(tau, tau), ntau2 = pk2(-tau, -tau), sn2 = pk2(sn, sn), nsn2 = pk2(-sn, -sn);
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            if (PAD || gl + LP * c < nch) {
              const unsigned long long a01 = pk2(av[c].x, av[c].y), a23 = pk2(av[c].z, av[c].w);
              const unsigned long long b01 = pk2(bv[c].x, bv[c].y), b23 = pk2(bv[c].z, bv[c].w);
              // a' = a - s (b + tau a), b' = b + s (a - tau b): four packed FMAs per pair of rows
              const unsigned long long na01 = ffma2(nsn2, ffma2(tau2, a01, b01), a01), na23 = ffma2(nsn2, ffma2(tau2, a23, b23), a23);
              const unsigned long long nb01 = ffma2(sn2, ffma2(ntau2, b01, a01), b01), nb23 = ffma2(sn2, ffma2(ntau2, b23, a23), b23);
              float4 na, nb;
              upk2(na01, na.x, na.y); upk2(na23, na.z, na.w);
              upk2(nb01, nb.x, nb.y); upk2(nb23, nb.z, nb.w);
              *reinterpret_cast<float4*>(up + 4 * LP * c) = na;
              *reinterpret_cast<float4*>(uq + 4 * LP * c) = nb;
            }
          }
          if (gl == 0) {
            nrm2[p] = fmaxf(fmaf(-t, ga, al), 0.f);
            nrm2[q] = fmaf(t, ga, be);
            if (a.work != nullptr) atomicAdd(&s_rot, 1u);
          }
        }
      };
      if (ngroups >= npairs) {
        // every pair of a round has its own lane group: the pair indices advance incrementally
        const bool act = grp < npairs;
        int p = 0, q = 0;
        if (act) rr_pair(n, 0, grp, p, q);
        for (int r = 0; r < m; ++r) {
          const bool valid = act && p < D && q < D;  // padding player of an odd D
          do_pair(valid ? p : 0, valid ? q : 0, valid);
          __syncthreads();
          if (grp == 0) {
            q = r + 1;
          } else {
            p = (p + 1 == m) ? 0 : p + 1;
            q = (q + 1 == m) ? 0 : q + 1;
          }
        }
      } else {
        for (int r = 0; r < m; ++r) {
          for (int pb = 0; pb < npairs; pb += ngroups) {
            const int pi = pb + grp;
            int p = 0, q = 0;
            bool valid = pi < npairs;
            if (valid) {
              rr_pair(n, r, pi, p, q);
              valid = (p < D) && (q < D);
            }
            do_pair(valid ? p : 0, valid ? q : 0, valid);
          }
          __syncthreads();
        }
      }
      ++sweeps;
      if (a.work != nullptr && tid == 0) s_dots += (unsigned)(D * (D - 1) / 2);
      wmax = warp_max(wmax);
      if (lane == 0) atomicMax(&s_flag, __float_as_uint(wmax));
      __syncthreads();
      const float fmaxoff = __uint_as_float(s_flag);   // largest squared cosine above tol^2, 0 if none
      __syncthreads();
      if (tid == 0) s_flag = 0u;
      if (fmaxoff <= 0.f) break;
      // Barrier-free check of all D(D-1)/2 cosines (no writes): when the sweep above already
      // converged the columns this replaces a whole no-rotation sweep of m synchronised rounds.
      // The few pairs that are still above the tolerance after a sweep (typically a handful of
      // 4950; they sent a fifth of the warm solves into a full extra sweep, and every launch lasts
      // as long as its slowest graph) are listed and rotated one at a time in a fix-up pass, then
      // the check runs again; a long list, or a second failed re-check, falls back to a full sweep.
      bool converged = false;
      for (int fixrounds = 0;; ++fixrounds) {
        float cmax = 0.f;
        const unsigned gmask = (LP >= 32) ? 0xffffffffu : (((1u << LP) - 1u) << (lane & ~(LP - 1)));
        for (int pp = grp; pp < (D + 1) / 2; pp += ngroups) {
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {   // rows pp and D-1-pp together: balanced work
            const int p = half ? D - 1 - pp : pp;
            if (half && p == pp) break;
            const float* up = Ul + (size_t)p * ld;
            float4 av[CH];
#pragma unroll
            for (int c = 0; c < CH; ++c)
              av[c] = (PAD || gl + LP * c < nch) ? *reinterpret_cast<const float4*>(up + 4 * LP * c)
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            const float al = nrm2[p];
            for (int q = p + 1; q < D; ++q) {
              const float* uq = Ul + (size_t)q * ld;
              unsigned long long g01 = 0ull, g23 = 0ull;
#pragma unroll
              for (int c = 0; c < CH; ++c) {
                if (PAD || gl + LP * c < nch) {
                  const float4 bv = *reinterpret_cast<const float4*>(uq + 4 * LP * c);
                  g01 = ffma2(pk2(av[c].x, av[c].y), pk2(bv.x, bv.y), g01);
                  g23 = ffma2(pk2(av[c].z, av[c].w), pk2(bv.z, bv.w), g23);
                }
              }
              float g0, g1, g2, g3;
              upk2(g01, g0, g1);
              upk2(g23, g2, g3);
              const float ga = group_sum_masked<LP>((g0 + g1) + (g2 + g3), gmask);
              const float den = al * nrm2[q];
              if (ga * ga > tol2 * den) {
                cmax = fmaxf(cmax, ga * ga / den);   // largest squared cosine above the tolerance
                if (gl == 0) {
                  const unsigned slot = atomicAdd(&s_nfix, 1u);
                  if (slot < (unsigned)MAXFIX) s_fix[slot] = ((unsigned)p << 16) | (unsigned)q;
                }
              }
            }
          }
        }
        if (cmax > 0.f) atomicMax(&s_flag, __float_as_uint(cmax));
        if (a.work != nullptr && tid == 0) s_dots += (unsigned)(D * (D - 1) / 2);
        __syncthreads();
        const float worst2 = __uint_as_float(s_flag);
        const bool more = worst2 > 0.f;
        const unsigned nfix = s_nfix;
        __syncthreads();
        if (tid == 0) { s_flag = 0u; s_nfix = 0u; }
        if (!more) { converged = true; break; }
        if (tid == 0) {
          if (nfix > (unsigned)MAXFIX || fixrounds >= 2) ++s_stat[1];
          else { s_stat[0] += (int)nfix; s_stat[2] += fixrounds > 0 ? 1 : 0; }
        }
        if (nfix > (unsigned)MAXFIX || fixrounds >= 2) { __syncthreads(); break; }   // full sweep
        // fix-up: one listed pair per round (they may share columns), in ascending (p, q) order so
        // that the result does not depend on the order in which the groups found them; every group
        // executes the same rotation arithmetic, group 0 alone writes
        for (unsigned e = tid; e < nfix; e += nthreads) {   // small D: fewer threads than list entries
          const unsigned mine = s_fix[e];
          unsigned rank = 0;
          for (unsigned j = 0; j < nfix; ++j) rank += (s_fix[j] < mine) ? 1u : 0u;
          s_fixs[rank] = mine;
        }
        __syncthreads();
        for (unsigned f = 0; f < nfix; ++f) {
          const unsigned pq = s_fixs[f];
          do_pair((int)(pq >> 16), (int)(pq & 0xffffu), grp == 0);
          __syncthreads();
        }
        // A rotation by theta in the (p, q) plane moves cos(p, r) by ~theta cos(q, r) <= theta tol: with
        // every listed |cos| < 1e-3 the untouched pairs stay within tol (1 + 1e-3) and the re-check
        // (one more pass over all D^2/2 dot products) is skipped.
        if (worst2 < 1e-6f) { converged = true; break; }
      }
      if (converged) break;
    }

    t_sweeps1 = clock64();
    // ---- column norms ----------------------------------------------------------------------
    float wsum_part = 0.f;
    for (int col = warp; col < D; col += nwarps) {
      const float* u = U + (size_t)col * ld;
      float s = 0.f;
      for (int r = lane; r < D; r += 32) s = fmaf(u[r], u[r], s);
      s = warp_sum(s);
      const float nrm = sqrtf(s);
      if (lane == 0) {
        wv[col] = nrm;
        wsum_part += nrm;
      }
    }
    wsum = block_sum(wsum_part, red);
    __syncthreads();
    if (a.shift_mode != 1 || attempt == 1) break;
    const float expect = trace + (float)D * sigma;
    if (fabsf(wsum - expect) <= 2e-3f * fabsf(expect)) break;  // A + sigma I was positive definite
    sweeps += 1000;                                            // mark the retry in info[0]
  }

  // ---- eigenvalues = column norms - sigma, eigenvectors = normalised columns ---------------
  float* Vb = a.Vt + (size_t)b * D * D;
  for (int col = warp; col < D; col += nwarps) {
    const float* u = U + (size_t)col * ld;
    const float nrm = wv[col];
    const float inv = (nrm > 0.f) ? 1.f / nrm : 0.f;
    for (int r = lane; r < D; r += 32) Vb[(size_t)col * D + r] = u[r] * inv;
  }
  __syncthreads();
  if (a.keepG) {
    // Rayleigh quotients with the kept G = A + sigma I: lambda_k = u^T G u / u^T u - sigma
    if (a.use_mma) {
      const int ntile = (D + 15) / 16;
      const int g = lane >> 2, t = lane & 3;
      for (int cg = warp; cg * 8 < D; cg += nwarps) {
        const int cA = cg * 8 + 2 * t, cB = cA + 1;
        const float* uA = U + (size_t)(cA < D ? cA : 0) * ld;
        const float* uB = U + (size_t)(cB < D ? cB : 0) * ld;
        float nA = 0.f, dA = 0.f, nB = 0.f, dB = 0.f;
        for (int tile0 = 0; tile0 < ntile; tile0 += MMA_RT) {
          float acc[MMA_RT][4];
          mma_g_times_cols(Gk, ld, D, U, ld, cg * 8, tile0, lane, acc);
#pragma unroll
          for (int rt = 0; rt < MMA_RT; ++rt) {
            const int m0 = (tile0 + rt) * 16 + g, m1 = m0 + 8;
            if (m0 < D) {
              const float a0 = uA[m0], b0 = uB[m0];
              nA = fmaf(acc[rt][0], a0, nA); dA = fmaf(a0, a0, dA);
              nB = fmaf(acc[rt][1], b0, nB); dB = fmaf(b0, b0, dB);
            }
            if (m1 < D) {
              const float a1 = uA[m1], b1 = uB[m1];
              nA = fmaf(acc[rt][2], a1, nA); dA = fmaf(a1, a1, dA);
              nB = fmaf(acc[rt][3], b1, nB); dB = fmaf(b1, b1, dB);
            }
          }
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {   // reduce over the 8 row groups g (lanes with equal t)
          nA += __shfl_xor_sync(0xffffffffu, nA, o); dA += __shfl_xor_sync(0xffffffffu, dA, o);
          nB += __shfl_xor_sync(0xffffffffu, nB, o); dB += __shfl_xor_sync(0xffffffffu, dB, o);
        }
        if (g == 0) {
          if (cA < D) nrm2[cA] = (dA > 0.f) ? nA / dA : 0.f;
          if (cB < D) nrm2[cB] = (dB > 0.f) ? nB / dB : 0.f;
        }
      }
    } else if constexpr (CH <= 4) {
      for (int cb = 0; cb < D; cb += 2 * ngroups) {
        const int c0 = cb + 2 * grp, c1 = c0 + 1;
        const float* u0 = U + (size_t)(c0 < D ? c0 : 0) * ld;
        const float* u1 = U + (size_t)(c1 < D ? c1 : 0) * ld;
        float4 t0[CH], t1[CH];
        group_matvec2<LP, CH>(Gk, u0, u1, D, ld, nch, gl, t0, t1);
        float n0 = 0.f, d0 = 0.f, n1 = 0.f, d1 = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int ch = gl + LP * c;
          if (ch < nch) {
            const float4 a0 = *reinterpret_cast<const float4*>(u0 + 4 * ch);
            const float4 a1 = *reinterpret_cast<const float4*>(u1 + 4 * ch);
            n0 = dot4(t0[c], a0, n0);  d0 = dot4(a0, a0, d0);
            n1 = dot4(t1[c], a1, n1);  d1 = dot4(a1, a1, d1);
          }
        }
        n0 = group_sum<LP>(n0);  d0 = group_sum<LP>(d0);
        n1 = group_sum<LP>(n1);  d1 = group_sum<LP>(d1);
        if (gl == 0) {
          if (c0 < D) nrm2[c0] = (d0 > 0.f) ? n0 / d0 : 0.f;
          if (c1 < D) nrm2[c1] = (d1 > 0.f) ? n1 / d1 : 0.f;
        }
      }
    } else {
      for (int cb = 0; cb < D; cb += ngroups) {
        const int col = cb + grp;
        const float* u = U + (size_t)(col < D ? col : 0) * ld;
        float4 tv[CH];
        group_matvec<LP, CH>(Gk, u, D, ld, nch, gl, tv);
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int ch = gl + LP * c;
          if (ch < nch) {
            const float4 uv = *reinterpret_cast<const float4*>(u + 4 * ch);
            num = dot4(tv[c], uv, num);
            den = dot4(uv, uv, den);
          }
        }
        num = group_sum<LP>(num);
        den = group_sum<LP>(den);
        if (col < D && gl == 0) nrm2[col] = (den > 0.f) ? num / den : 0.f;
      }
    }
    __syncthreads();
    for (int i = tid; i < D; i += nthreads) wv[i] = nrm2[i];
    __syncthreads();
  }
  for (int i = tid; i < D; i += nthreads) {
    const float ev = wv[i] - sigma;
    wv[i] = ev;
    a.w[(size_t)b * D + i] = ev;
  }
  wsum -= (float)D * sigma;
  if (a.work != nullptr && tid == 0) {
    atomicAdd(a.work, (unsigned long long)s_dots);
    atomicAdd(a.work + 1, (unsigned long long)s_rot);
  }
  if (a.info && tid == 0) {
    float* o = a.info + (size_t)b * 4;
    if (a.retry_only) sweeps += 1000;
    o[0] = (float)sweeps;
    o[1] = sigma;
    o[2] = trace;
    o[3] = wsum;
    if (a.timing == 1) {  // developer knob "eig_timing": phase cycle counts instead of shift / trace / sum
      o[1] = (float)(t_sweeps0 - t_start);
      o[2] = (float)(t_sweeps1 - t_sweeps0);
      o[3] = (float)(clock64() - t_sweeps1);
    } else if (a.timing == 2) {  // convergence branches: fix-up rotations, full-sweep fallbacks, re-checks after a fix-up
      o[1] = (float)s_stat[0];
      o[2] = (float)s_stat[1];
      o[3] = (float)s_stat[2];
    }
  }
  __syncthreads();

  // ---- tails -----------------------------------------------------------------------------
  if (a.tail == TAIL_LAYER) {
    const double c4 = 4.0 / (double)a.lam[0];
    double part = 0.0;
    for (int i = tid; i < D; i += nthreads) {
      const double be = wv[i];
      const double mu = be * be + c4;
      part += mu * mu;
    }
    const double nrm = sqrt(block_sum_d(part, redd));
    double part2 = 0.0;
    for (int i = tid; i < D; i += nthreads) {
      const double be = wv[i];
      const double mu = be * be + c4;
      double s;
      if (a.exact_sqrt) {
        s = sqrt(mu);
      } else {
        double y = mu / nrm, z = 1.0;
#pragma unroll
        for (int t = 0; t < UGLAD_NS_ITERS; ++t) {
          const double T = 0.5 * (3.0 - z * y);
          y = y * T;
          z = T * z;
        }
        s = y * sqrt(nrm);
      }
      a.sroot[(size_t)b * D + i] = (float)s;
      a.f[(size_t)b * D + i] = (float)(0.5 * (s - be));
      part2 += s * s;
    }
    const double sn = sqrt(block_sum_d(part2, redd));
    if (tid == 0) a.snorm[b] = (float)sn;
  } else if (a.tail == TAIL_LOSS) {
    // torch.logdet: log(det) -> NaN when det < 0.  With sigma == 0 the solver returns |eig|;
    // sum|eig| == trace iff no eigenvalue is negative.
    double part = 0.0;
    for (int i = tid; i < D; i += nthreads) {
      const float ev = wv[i];
      part += log((double)ev);
      a.f[(size_t)b * D + i] = -1.f / ev;
    }
    const double ld_ = block_sum_d(part, redd);
    if (tid == 0) {
      const bool pd = fabsf(wsum - trace) <= 1e-4f * fabsf(wsum);
      a.snorm[b] = pd ? (float)ld_ : __int_as_float(0x7fc00000);
    }
  }
}

static int g_tune_lp = 0;      // 0 = auto; otherwise force lanes per column pair (4/8/16/32)
static int g_tune_tol_1e7 = 0;  // 0 = EigArgs::tol; otherwise the cosine tolerance in units of 1e-7
static int g_tune_keepg = -1;  // -1 = auto (keep G when two buffers fit); 0 / 1 force
static int g_tune_timing = 0;
static int g_tune_mma = 1;     // 1 = warm-start product and Rayleigh quotients on the mma.sync tensor path
static int g_tune_pad = 1;     // 1 = pad shared-memory columns to LP*CH*4 floats when it fits
int eig_small_timing() { return g_tune_timing; }
int eig_small_tune(const char* key, int value) {
  if (!eig_cluster_tune(key, value)) return 0;
  if (!strcmp(key, "eig_lp")) { g_tune_lp = value; return 0; }
  if (!strcmp(key, "eig_keepg")) { g_tune_keepg = value; return 0; }
  if (!strcmp(key, "eig_tol_1e7")) { g_tune_tol_1e7 = value; return 0; }
  if (!strcmp(key, "eig_pad")) { g_tune_pad = value; return 0; }
  if (!strcmp(key, "eig_timing")) { g_tune_timing = value; return 0; }
  if (!strcmp(key, "eig_mma")) { g_tune_mma = value; return 0; }
  return 1;
}

template <int LP, int CH, bool PAD>
static int launch_cfg(const EigArgs& a, int B, size_t smem, cudaStream_t st) {
  const int npairs = (a.D + 1) / 2;
  int threads = ((npairs * LP + 31) / 32) * 32;
  if (threads > eig_max_threads(LP, CH)) threads = eig_max_threads(LP, CH);
  if (threads < 128) threads = 128;
  if (threads < a.D) threads = ((a.D + 31) / 32) * 32;  // the power iteration wants a thread per row
  UGLAD_CUDA(cudaFuncSetAttribute(eig_jacobi_small_kernel<LP, CH, PAD>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (!a.retry_only) profile_begin(st, 0, (double)B * (4.0 * a.D * a.D + 3.0 * a.D) * 4.0);   // (the retry pass is not a solve)
  eig_jacobi_small_kernel<LP, CH, PAD><<<B, threads, smem, st>>>(a);
  if (!a.retry_only) profile_end(st);
  UGLAD_CHECK_LAUNCH("eig_jacobi_small_kernel");
  return 0;
}

int launch_eig_small(const EigArgs& a_in, int B, cudaStream_t st) {
  EigArgs a = a_in;
  a.timing = g_tune_timing;
  a.work = profile_eig_counters();
  a.use_mma = g_tune_mma;
  a.ld = (a.D + 3) & ~3;
  if (a.D > UGLAD_SMALL_D_MAX) {
    set_error("eig_small: D=%d exceeds the shared-memory solver limit %d", a.D, UGLAD_SMALL_D_MAX);
    return 1;
  }
  if (a.U0 != nullptr && a.info != nullptr && !a.retry_only) {
    // small batches: the graph's column pairs spread over a cluster of CTAs (eig_cluster.cu), then a retry pass of
    // this kernel for the graphs whose warm start failed the positive-definiteness check (normally none: the
    // CTAs of the pass exit at once)
    const int nc = eig_cluster_size(B, a.D);
    if (nc > 0) {
      const int rc = launch_eig_cluster(a, B, nc, st);
      if (rc == 1) return 1;
      if (rc == 0) {
        EigArgs r = a_in;
        r.retry_only = 1;
        return launch_eig_small(r, B, st);
      }
    }
  }
  int nch = a.ld / 4;
  // lanes per column pair: narrow groups replicate the rotation scalar math less (the kernel
  // is issue-bound), wide groups shorten the per-round dependent chain.
  int lp = g_tune_lp;
  if (lp == 0) lp = (nch <= 8) ? 4 : (nch <= 32 ? 8 : 16);
  while (lp < 32 && lp * 8 < nch) lp *= 2;
  const int ch = (nch + lp - 1) / lp;
  // the template CH that the case list below will pick for (lp, ch)
  int chT = ch <= 1 ? 1 : (ch <= 2 ? 2 : (ch <= 4 ? 4 : 8));
  if (lp == 4 && chT < 2) chT = 2;
  auto fits = [&](int ld, int nbuf) {
    const size_t mat = (size_t)ld * a.D * sizeof(float);
    const size_t extra = 2 * ld * sizeof(float) + 32 * sizeof(double) + 32 * sizeof(float) + 16;
    return nbuf * mat + extra <= 227 * 1024;
  };
  // padded columns (branch-free chunk loops) when that does not cost the second buffer
  const int ld_pad = lp * chT * 4 + 4;   // +4: consecutive columns start 4 banks apart (no conflicts on scalar loads)
  const bool want2 = fits(a.ld, 2);
  bool pad = g_tune_pad != 0 && chT <= 4 && fits(ld_pad, want2 ? 2 : 1);
  if (pad) a.ld = ld_pad;
  const size_t mat = (size_t)a.ld * a.D * sizeof(float);
  const size_t extra = 2 * a.ld * sizeof(float) + 32 * sizeof(double) + 32 * sizeof(float) + 16;
  const bool fits2 = 2 * mat + extra <= 227 * 1024;
  a.keepG = (g_tune_keepg < 0) ? (fits2 ? 1 : 0) : (g_tune_keepg && fits2 ? 1 : 0);
  if (a.U0 != nullptr) { a.keepG = 0; a.warmVt = nullptr; a.warm_w = nullptr; }   // pre-multiplied start: one buffer
  if (g_tune_tol_1e7 > 0) a.tol = 1e-7f * (float)g_tune_tol_1e7;
  if (!a.keepG) a.warmVt = nullptr;
  const size_t smem = (a.keepG ? 2 : 1) * mat + extra;
#define UGLAD_EIG_CASE(LP_, CH_)                                                    \
  if (lp == LP_ && ch <= CH_)                                                       \
    return pad ? launch_cfg<LP_, CH_, true>(a, B, smem, st) : launch_cfg<LP_, CH_, false>(a, B, smem, st)
  UGLAD_EIG_CASE(4, 2); UGLAD_EIG_CASE(4, 4); UGLAD_EIG_CASE(4, 8);
  UGLAD_EIG_CASE(8, 1); UGLAD_EIG_CASE(8, 2); UGLAD_EIG_CASE(8, 4); UGLAD_EIG_CASE(8, 8);
  UGLAD_EIG_CASE(16, 1); UGLAD_EIG_CASE(16, 2); UGLAD_EIG_CASE(16, 4);
  UGLAD_EIG_CASE(32, 1); UGLAD_EIG_CASE(32, 2);
#undef UGLAD_EIG_CASE
  set_error("eig_small: no kernel configuration for D=%d lp=%d", a.D, lp);
  return 1;
}

}  // namespace uglad
