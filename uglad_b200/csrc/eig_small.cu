// One-CTA-per-graph symmetric eigensolver for D <= UGLAD_SMALL_D_MAX (sm_100a).
//
// Algorithm: one-sided (Hestenes) Jacobi on U = A + sigma*I held column-major in shared
// memory.  With sigma chosen so that A + sigma*I is positive definite (indefinite inputs
// such as b = S/lambda - theta of glad.py:139), orthogonalising the columns of U gives
// U -> V diag(lambda + sigma): eigenvalues are column norms minus sigma, eigenvectors the
// normalised columns, so ONE D x D matrix in shared memory is the whole state (D <= 232
// fits the 227 KB of a B200 SM).  A warp owns a column pair per step of a round-robin
// tournament; the three dot products are warp-shuffle reductions; one __syncthreads per
// round.  The tail turns eigenvalues into what the caller needs:
//   TAIL_LAYER: f_k = (s_k - beta_k)/2 with s_k the reference's 10-step Newton-Schulz
//               square root of beta_k^2 + 4/lambda collapsed onto the eigenvalues
//               (torch_sqrtm.py:12-28, glad.py:140-142);
//   TAIL_LOSS : logdet and -1/eig for main.py:307 (torch.logdet) and its gradient.
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

template <int CH>
__global__ void __launch_bounds__(1024, 1) eig_jacobi_small_kernel(EigArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int D = a.D, ld = a.ld;
  float* U = smem;                                // [D][ld] column-major
  float* wv = U + (size_t)ld * D;                 // [ld] eigenvalues
  double* redd = reinterpret_cast<double*>(wv + ld + (ld & 1));  // [32] reduction scratch
  float* red = reinterpret_cast<float*>(redd + 32);              // [32]
  __shared__ unsigned s_flag;
  __shared__ float s_sigma;

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  const int nch = ld >> 2;

  // ---- load (coalesced along the contiguous global dimension) ---------------------------
  float trace_part = 0.f, fro_part = 0.f;
  {
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
    for (int idx = tid; idx < D * ld; idx += nthreads) {
      const int col = idx / ld, row = idx - col * ld;
      float v = 0.f;
      if (row < D) {
        const size_t g = (size_t)col * D + row;
        v = a.build ? (inv_lam * Sb[g] - Tb[g]) : Ab[g];
        fro_part += v * v;
        if (row == col) trace_part += v;
      }
      U[idx] = v;
    }
  }
  if (tid == 0) s_flag = 0u;
  const float trace = block_sum(trace_part, red);
  const float fro = sqrtf(block_sum(fro_part, red));
  // Gershgorin: max column abs-sum (symmetric input)
  float gmax = 0.f;
  for (int col = warp; col < D; col += nwarps) {
    float s = 0.f;
    for (int r = lane; r < D; r += 32) s += fabsf(U[(size_t)col * ld + r]);
    s = warp_sum(s);
    gmax = fmaxf(gmax, s);
  }
  gmax = block_max(gmax, red);
  if (tid == 0) {
    float sg = 0.f;
    if (a.shift_mode == 1) {
      const float bound = fminf(gmax, fro);
      sg = (bound > 0.f) ? 1.25f * bound : 1.0f;
    }
    s_sigma = sg;
  }
  __syncthreads();
  const float sigma = s_sigma;
  if (sigma != 0.f)
    for (int i = tid; i < D; i += nthreads) U[(size_t)i * ld + i] += sigma;
  __syncthreads();

  // ---- Jacobi sweeps ---------------------------------------------------------------------
  const int n = D + (D & 1), npairs = n >> 1, m = n - 1;
  const float tol = a.tol;
  int sweeps = 0;
  for (int sweep = 0; sweep < a.max_sweeps; ++sweep) {
    float wmax = 0.f;
    for (int r = 0; r < m; ++r) {
      for (int pi = warp; pi < npairs; pi += nwarps) {
        int p, q;
        rr_pair(n, r, pi, p, q);
        if (p >= D || q >= D) continue;  // padding player of an odd D
        float* up = U + (size_t)p * ld;
        float* uq = U + (size_t)q * ld;
        float4 av[CH], bv[CH];
        float al = 0.f, be = 0.f, ga = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int ch = lane + 32 * c;
          if (ch < nch) {
            av[c] = *reinterpret_cast<const float4*>(up + 4 * ch);
            bv[c] = *reinterpret_cast<const float4*>(uq + 4 * ch);
          } else {
            av[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            bv[c] = av[c];
          }
          al = fmaf(av[c].x, av[c].x, fmaf(av[c].y, av[c].y, fmaf(av[c].z, av[c].z, fmaf(av[c].w, av[c].w, al))));
          be = fmaf(bv[c].x, bv[c].x, fmaf(bv[c].y, bv[c].y, fmaf(bv[c].z, bv[c].z, fmaf(bv[c].w, bv[c].w, be))));
          ga = fmaf(av[c].x, bv[c].x, fmaf(av[c].y, bv[c].y, fmaf(av[c].z, bv[c].z, fmaf(av[c].w, bv[c].w, ga))));
        }
        al = warp_sum(al);
        be = warp_sum(be);
        ga = warp_sum(ga);
        const float den = al * be;
        const float off = (den > 0.f) ? fabsf(ga) / sqrtf(den) : 0.f;
        wmax = fmaxf(wmax, off);
        if (off > tol) {
          const float zeta = (be - al) / (2.f * ga);
          const float az = fabsf(zeta);
          float t = 1.f / (az + sqrtf(fmaf(az, az, 1.f)));
          t = (zeta < 0.f) ? -t : t;
          const float cs = 1.f / sqrtf(fmaf(t, t, 1.f));
          const float sn = cs * t;
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const int ch = lane + 32 * c;
            if (ch < nch) {
              float4 na, nb;
              na.x = cs * av[c].x - sn * bv[c].x;  nb.x = sn * av[c].x + cs * bv[c].x;
              na.y = cs * av[c].y - sn * bv[c].y;  nb.y = sn * av[c].y + cs * bv[c].y;
              na.z = cs * av[c].z - sn * bv[c].z;  nb.z = sn * av[c].z + cs * bv[c].z;
              na.w = cs * av[c].w - sn * bv[c].w;  nb.w = sn * av[c].w + cs * bv[c].w;
              *reinterpret_cast<float4*>(up + 4 * ch) = na;
              *reinterpret_cast<float4*>(uq + 4 * ch) = nb;
            }
          }
        }
      }
      __syncthreads();
    }
    ++sweeps;
    if (lane == 0) atomicMax(&s_flag, __float_as_uint(wmax));
    __syncthreads();
    const float fmaxoff = __uint_as_float(s_flag);
    __syncthreads();
    if (tid == 0) s_flag = 0u;
    if (fmaxoff <= tol) break;
  }

  // ---- eigenvalues = column norms - sigma, eigenvectors = normalised columns -------------
  float* Vb = a.Vt + (size_t)b * D * D;
  for (int col = warp; col < D; col += nwarps) {
    const float* u = U + (size_t)col * ld;
    float s = 0.f;
    for (int r = lane; r < D; r += 32) s = fmaf(u[r], u[r], s);
    s = warp_sum(s);
    const float nrm = sqrtf(s);
    const float inv = (nrm > 0.f) ? 1.f / nrm : 0.f;
    for (int r = lane; r < D; r += 32) Vb[(size_t)col * D + r] = u[r] * inv;
    if (lane == 0) wv[col] = nrm - sigma;
  }
  __syncthreads();
  float wsum_part = 0.f;
  for (int i = tid; i < D; i += nthreads) {
    a.w[(size_t)b * D + i] = wv[i];
    wsum_part += wv[i];
  }
  const float wsum = block_sum(wsum_part, red);
  if (a.info && tid == 0) {
    float* o = a.info + (size_t)b * 4;
    o[0] = (float)sweeps;
    o[1] = sigma;
    o[2] = trace;
    o[3] = wsum;
  }

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

int launch_eig_small(const EigArgs& a_in, int B, cudaStream_t st) {
  EigArgs a = a_in;
  a.ld = (a.D + 3) & ~3;
  if (a.D > UGLAD_SMALL_D_MAX) {
    set_error("eig_small: D=%d exceeds the shared-memory solver limit %d", a.D, UGLAD_SMALL_D_MAX);
    return 1;
  }
  const size_t smem = ((size_t)a.ld * a.D + a.ld + 2) * sizeof(float) + 32 * sizeof(double) + 32 * sizeof(float) + 16;
  const int npairs = (a.D + 1) / 2;
  int nwarps = npairs < 4 ? 4 : (npairs > 32 ? 32 : npairs);
  if (B >= 148 && nwarps > 16) nwarps = 16;  // throughput mode: more CTAs per SM
  const int threads = nwarps * 32;
  if (a.ld <= 128) {
    UGLAD_CUDA(cudaFuncSetAttribute(eig_jacobi_small_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    eig_jacobi_small_kernel<1><<<B, threads, smem, st>>>(a);
  } else {
    UGLAD_CUDA(cudaFuncSetAttribute(eig_jacobi_small_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    eig_jacobi_small_kernel<2><<<B, threads, smem, st>>>(a);
  }
  UGLAD_CHECK_LAUNCH("eig_jacobi_small_kernel");
  return 0;
}

}  // namespace uglad
