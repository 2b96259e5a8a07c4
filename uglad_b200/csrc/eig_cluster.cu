// Warm-started symmetric eigensolver spread over a thread-block CLUSTER (sm_100a): one graph per
// cluster of 1 / 2 / 4 CTAs, for the batches that leave SMs idle with one CTA per graph
// (configs[1]: ONE graph at D = 100 used one SM of 148; configs[2] sharded over 8 GPUs: 32 graphs).
//
// Same mathematics as eig_small.cu (one-sided Jacobi on U0 = (A + sigma I) V_prev, the pre-multiplied
// warm start of uglad_glad_layer_forward; glad.py:139-142 is the caller this replaces), different
// schedule:
//   * odd-even ordering instead of the round-robin tournament: the columns sit on a line of
//     positions 0 .. D-1, step s pairs (2g + (s&1), 2g + 1 + (s&1)), rotates and SWAPS them, so that
//     after D steps every column has met every other one (the order is reversed);
//   * the lane group g keeps the column at the odd position 2g+1 in REGISTERS for the whole sweep
//     and exchanges only the even-position column with its neighbours through shared memory: one
//     column read + one column written per pair and step instead of two + two;
//   * the groups of one graph are split into contiguous ranges over the CTAs of the cluster.  Only
//     the two boundary columns of a CTA cross to a neighbour per step; the writer PUSHES its result
//     into the shared memory of the CTA whose group reads it next (st.shared::cluster), so every read
//     is local.  One cluster barrier per step (arrive.release right after the stores, wait.acquire
//     before the next loads);
//   * after a sweep every CTA pushes its columns to all others (each holds the full matrix), checks
//     a 1/nc share of the D(D-1)/2 cosines, the lists of pairs still above the tolerance are
//     exchanged and merged in sorted order, and every CTA applies the same fix-up rotations to its
//     own copy (identical arithmetic: the copies stay bit-identical), as eig_small.cu does in one CTA.
// A start that turns out not to be positive definite (sum of column norms != trace + D sigma) is
// flagged in info[0] (+1000); the launcher then runs the one-CTA kernel in retry mode for those graphs.
#include <string.h>
#include <type_traits>
#include "common.cuh"
#include "kernels.cuh"

namespace uglad {

namespace oe {

__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
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
template <int LP>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int LP>
__device__ __forceinline__ float group_sum_masked(float v, unsigned mask) {
#pragma unroll
  for (int o = LP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in the CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_rank(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, const float4& v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_f1(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_u1(uint32_t addr, unsigned v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// asynchronous remote store that reports its bytes to an mbarrier of the destination CTA (both shared::cluster
// addresses): data and signal travel together, the reader waits on its LOCAL barrier -- no fence, no round trip
__device__ __forceinline__ void st_async_f4(uint32_t addr, const float4& v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(mbar) : "memory");
}
__device__ __forceinline__ void st_async_f1(uint32_t addr, float v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];"
               ::"r"(addr), "f"(v), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arm(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  for (int spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && (spin & 1023) == 1023) {  // a lost arrival must fault, never hang the GPU
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// Jacobi rotation of the column pair (a, b) with squared norms (al, be) and dot product ga, in the
// small-angle-accurate form of eig_small.cu.  Returns false (columns untouched) below the tolerance.
struct Rot {
  float t, sn, tau;
};
__device__ __forceinline__ bool rotation(float al, float be, float ga, float tol2, Rot& r) {
  if (!(ga * ga > tol2 * (al * be))) return false;
  const float d = be - al, g2 = 2.f * ga;
  const float h = fast_sqrt(fmaf(d, d, g2 * g2));
  float t = fabsf(g2) * fast_rcp(fabsf(d) + h);
  t = ((d < 0.f) != (g2 < 0.f)) ? -t : t;
  const float x = fmaf(t, t, 1.f);
  float cs = fast_rsqrt(x);
  cs = cs * fmaf(-0.5f * x, cs * cs, 1.5f);
  r.t = t;
  r.sn = t * cs;
  r.tau = r.sn * fast_rcp(1.f + cs);
  return true;
}
// a' = a - s (b + tau a), b' = b + s (a - tau b) on one float4 chunk of both columns
__device__ __forceinline__ void rotate4(const Rot& r, float4& a, float4& b) {
  const unsigned long long tau2 = pk2(r.tau, r.tau), ntau2 = pk2(-r.tau, -r.tau), sn2 = pk2(r.sn, r.sn), nsn2 = pk2(-r.sn, -r.sn);
  const unsigned long long a01 = pk2(a.x, a.y), a23 = pk2(a.z, a.w), b01 = pk2(b.x, b.y), b23 = pk2(b.z, b.w);
  const unsigned long long na01 = ffma2(nsn2, ffma2(tau2, a01, b01), a01), na23 = ffma2(nsn2, ffma2(tau2, a23, b23), a23);
  const unsigned long long nb01 = ffma2(sn2, ffma2(ntau2, b01, a01), b01), nb23 = ffma2(sn2, ffma2(ntau2, b23, a23), b23);
  upk2(na01, a.x, a.y); upk2(na23, a.z, a.w);
  upk2(nb01, b.x, b.y); upk2(nb23, b.z, b.w);
}
// The sweep's rotation scalars, branch-free and with three MUFU operations on the dependent chain (the step loop is
// one long dependent chain: its instruction count is its time): with h = sqrt(d^2 + g2^2), d = be - al, g2 = 2 ga,
//   cos^2 = (1 + |d| / h) / 2,  sin = sign(d) g2 / (2 h cos),  tau = sin / (1 + cos),  tan = sin / cos
// (the same angle as rotation(): tan 2 theta = g2 / d, smaller root).  MUFU accuracy (2^-22) is enough here: in the
// small-angle-accurate update form the error of the applied rotation scales with the rotation itself, and the
// fix-up pass polishes with rotation().  `rot` false leaves sn = tau = t = 0: the update is then the identity.
__device__ __forceinline__ void rotation_fast(float al, float be, float ga, float tol2, bool valid, Rot& r, bool& rot) {
  const float d = be - al, g2 = ga + ga;
  rot = valid && (ga * ga > tol2 * (al * be));
  const float rh = fast_rsqrt(fmaf(d, d, g2 * g2));
  const float c2 = fmaf(0.5f * fabsf(d), rh, 0.5f);
  const float rcs = fast_rsqrt(c2);
  const float cs = c2 * rcs;
  float sn = (0.5f * g2) * (rh * rcs);
  sn = (d < 0.f) ? -sn : sn;
  const float tau = sn * fast_rcp(1.f + cs);
  r.sn = rot ? sn : 0.f;
  r.tau = rot ? tau : 0.f;
  r.t = rot ? sn * rcs : 0.f;
}
__device__ __forceinline__ float dot4acc(const float4& a, const float4& b, unsigned long long& g01, unsigned long long& g23) {
  g01 = ffma2(pk2(a.x, a.y), pk2(b.x, b.y), g01);
  g23 = ffma2(pk2(a.z, a.w), pk2(b.z, b.w), g23);
  return 0.f;
}
__device__ __forceinline__ float fold(unsigned long long g01, unsigned long long g23) {
  float g0, g1, g2, g3;
  upk2(g01, g0, g1);
  upk2(g23, g2, g3);
  return (g0 + g1) + (g2 + g3);
}

constexpr int MAXFIX = 192;
constexpr int MAXNC = 4;

}  // namespace oe

constexpr int oe_max_threads(int, int) { return 512; }   // 128 registers: the check pass keeps LP partial sums per lane

// ld = LP * CH * 4 + 4 floats per column (rows beyond D are zero): every lane owns CH float4 chunks of a
// column, no predicates in the chunk loops; +4 keeps consecutive columns 4 banks apart.
// MINB = 2: the one-CTA-per-graph form for full batches (64 registers, two CTAs per SM)
// MAXT = 448 with MINB = 2: 72 registers (with 64 the step loop re-forms its pointers from special registers and
// constants at the top of every step, on the dependent chain)
template <int LP, int CH, int MINB, int MAXT = 512>
__global__ void __launch_bounds__(MAXT, MINB) eig_jacobi_oe_kernel(EigArgs a, int nc, int gpc) {
  using namespace oe;
  extern __shared__ __align__(16) float smem[];
  const int D = a.D, ld = a.ld;
  const int Dp = (D + 3) & ~3;
  float* F = smem;                           // [D][ld] column at every position (this CTA's copy)
  float* nrm2 = F + (size_t)ld * (D + 1);    // [Dp + 4] squared column norms by position (+ the scratch column's, at [D])
  float* wv = nrm2 + Dp + 4;                 // [Dp] final column norms; F[D] is a scratch column (the idle group's mirror)
  float* red = wv + Dp;                      // [32]
  __shared__ unsigned s_flag, s_nfix, s_rot;
  __shared__ unsigned s_fix[MAXFIX];
  __shared__ unsigned s_fixs[MAXFIX];
  __shared__ unsigned s_owner[256];   // fix-up rounds: lowest pending list index per column (D <= 200)
  __shared__ unsigned s_left;
  __shared__ unsigned x_cnt[2][MAXNC], x_flag[2][MAXNC];   // exchanged between the CTAs, double-buffered by exchange parity
  __shared__ unsigned x_fix[2][MAXNC][MAXFIX];
  __shared__ __align__(8) unsigned long long s_mbar[2];   // boundary columns arriving from rank - 1 / rank + 1

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  constexpr bool CL = (MINB == 1);   // the cluster forms; MINB = 2 is launched with one CTA per graph only
  if (!CL) nc = 1;
  const uint32_t rank = (CL && nc > 1) ? cluster_rank() : 0u;
  const int b = blockIdx.x / nc;
  const int ng = D >> 1;                     // lane groups of the graph (D is even on this path)
  const int lg = tid / LP, gl = tid % LP, lgroups = nthreads / LP;
  const int g0 = (int)rank * gpc;
  const int g = g0 + lg;
  const bool act = lg < gpc && g < ng;
  const int gc = act ? g : (ng - 1);         // clamped: inactive groups compute on valid addresses, never store
  const float tol = a.tol, tol2 = tol * tol;
  const long long t_start = clock64();

  auto block_barrier = [&]() {
    if (CL && nc > 1) { __syncwarp(); cluster_arrive(); cluster_wait(); }
    else __syncthreads();
  };

  // ---- load U0 (every CTA takes the whole matrix: 4 D^2 bytes from L2) -----------------------
  {
    const float* U0b = a.U0 + (size_t)b * D * a.ldu;
    const int nch = ld >> 2;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(U0b) & 15) == 0) && (a.ldu % 4 == 0);
    for (int idx = tid; idx < D * nch; idx += nthreads) {
      const int col = idx / nch, c4 = (idx - col * nch) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (vec_ok && c4 + 3 < D) v = *reinterpret_cast<const float4*>(U0b + (size_t)col * a.ldu + c4);
      else if (c4 < D) {
        const float* s = U0b + (size_t)col * a.ldu + c4;
        v.x = s[0];
        if (c4 + 1 < D) v.y = s[1];
        if (c4 + 2 < D) v.z = s[2];
        if (c4 + 3 < D) v.w = s[3];
      }
      *reinterpret_cast<float4*>(F + (size_t)col * ld + c4) = v;
    }
    if (tid == 0) {
      s_flag = 0u; s_nfix = 0u; s_rot = 0u;
      if (CL && nc > 1) {
        mbar_init(smem_u32(&s_mbar[0]), 1u);
        mbar_init(smem_u32(&s_mbar[1]), 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arm(smem_u32(&s_mbar[0]), (uint32_t)(LP * CH * 16 + 4));
        mbar_arm(smem_u32(&s_mbar[1]), (uint32_t)(LP * CH * 16 + 4));
      }
    }
  }
  const float sigma = a.pre_sigma[b], trace = a.pre_trace[b];
  block_barrier();   // (cluster: every barrier is initialised before a neighbour can signal it)

  // neighbour copies of F / nrm2 (shared::cluster addresses); rank itself maps to the local window
  const uint32_t F_u32 = smem_u32(F), N_u32 = smem_u32(nrm2);
  uint32_t Fr[MAXNC], Nr[MAXNC];
#pragma unroll
  for (int r = 0; r < MAXNC; ++r) {
    Fr[r] = (CL && nc > 1 && r < nc) ? map_rank(F_u32, (uint32_t)r) : F_u32;
    Nr[r] = (CL && nc > 1 && r < nc) ? map_rank(N_u32, (uint32_t)r) : N_u32;
  }
  // destination CTA of this group's step output: the group that reads the even position next
  //   even step: position 2g is read by group g-1 in the odd step (group 0 keeps it)
  //   odd step : position 2g+2 is read by group g+1 in the even step
  const int r_even = (gc > 0) ? (gc - 1) / gpc : 0, r_odd = (gc + 1 < ng) ? (gc + 1) / gpc : (int)rank;
  uint32_t F_even = F_u32, F_odd = F_u32, N_even = N_u32, N_odd = N_u32;
#pragma unroll
  for (int r = 0; r < MAXNC; ++r) {
    if (r == r_even) { F_even = Fr[r]; N_even = Nr[r]; }
    if (r == r_odd) { F_odd = Fr[r]; N_odd = Nr[r]; }
  }

  // the two groups at the ends of this CTA's range trade one column per step with the neighbour CTAs
  const int last_lg = min(gpc, ng - g0) - 1;
  const bool edge_lo = CL && nc > 1 && act && lg == 0 && rank > 0;             // even steps: reads from / pushes to rank - 1
  const bool edge_hi = CL && nc > 1 && act && lg == last_lg && g + 1 < ng;     // odd steps : reads from / pushes to rank + 1
  const uint32_t mb_lo = smem_u32(&s_mbar[0]), mb_hi = smem_u32(&s_mbar[1]);
  const uint32_t mb_peer_lo = edge_lo ? map_rank(mb_hi, rank - 1) : 0u;  // my downward push lands in rank-1's "from above" barrier
  const uint32_t mb_peer_hi = edge_hi ? map_rank(mb_lo, rank + 1) : 0u;
  const uint32_t tx_bytes = (uint32_t)(LP * CH * 16 + 4);
  unsigned ph_lo = 0, ph_hi = 0;
  int sweeps = 0, xpar = 0;
  int st_fix = 0, st_over = 0, st_maxlist = 0, st_rounds = 0;   // developer knob eig_timing = 2: fix-up rotations, full-sweep reasons, longest list
  unsigned rot_count = 0;
  bool converged = false;
  long long t_sweep_cycles = 0, t_check_cycles = 0, t_fix_cycles = 0, t_load_end = 0;
  for (int sweep = 0; sweep < a.max_sweeps && !converged; ++sweep) {
    const long long t_s0 = clock64();
    if (sweep == 0) t_load_end = t_s0;
    // ---- the group's register column (odd position) and the norms of its two columns --------
    float4 av[CH];
    float na;
    {
      const float* uo = F + (size_t)(2 * gc + 1) * ld + 4 * gl;
      const float* ue = F + (size_t)(2 * gc) * ld + 4 * gl;
      unsigned long long s01 = 0ull, s23 = 0ull, e01 = 0ull, e23 = 0ull;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        av[c] = *reinterpret_cast<const float4*>(uo + 4 * LP * c);
        const float4 ev = *reinterpret_cast<const float4*>(ue + 4 * LP * c);
        dot4acc(av[c], av[c], s01, s23);
        dot4acc(ev, ev, e01, e23);
      }
      na = group_sum<LP>(fold(s01, s23));
      const float ne = group_sum<LP>(fold(e01, e23));
      __syncthreads();   // (a fallback sweep) every read of the old norms is over
      if (act && gl == 0) nrm2[2 * g] = ne;
    }
    __syncthreads();
    // ---- D steps of the odd-even ordering: D / 2 (even, odd) pairs.  The loop body is a dependent chain whose
    // instruction count IS its time, so: everything that depends only on the step's parity is formed once out
    // here; the two register arrays trade roles instead of being copied (a step loads the partner into `oth`,
    // rotates, stores the rotated `cur` and continues with `oth` as its own column); and the one group that has no
    // partner in the odd steps (the last one: position D - 1 waits there) pairs with a MIRROR of its own column in
    // the scratch slot F[D] with the rotation switched off, so that the role swap stays unconditional.
    {
      const bool is_last = act && (g == ng - 1);
      const bool valid_o = act && !is_last;
      float* const ue_e = F + (size_t)(2 * gc) * ld + 4 * gl;                                       // even step: position 2g
      float* const ue_o = F + (size_t)(valid_o ? 2 * gc + 2 : (is_last ? D : 0)) * ld + 4 * gl;     // odd step : position 2g + 2
      float* const ne_e = nrm2 + 2 * gc;
      float* const ne_o = nrm2 + (valid_o ? 2 * gc + 2 : (is_last ? D : 0));
      const uint32_t dstF_e = F_even + (uint32_t)(((size_t)(2 * gc) * ld + 4 * gl) * sizeof(float));
      const uint32_t dstF_o = F_odd + (uint32_t)(((size_t)(valid_o ? 2 * gc + 2 : 0) * ld + 4 * gl) * sizeof(float));
      const uint32_t dstN_e = N_even + (uint32_t)(2 * gc * sizeof(float));
      const uint32_t dstN_o = N_odd + (uint32_t)((valid_o ? 2 * gc + 2 : 0) * sizeof(float));
      auto step = [&](auto oddc, bool first, float4 (&cur)[CH], float4 (&oth)[CH]) {
        constexpr bool ODD = decltype(oddc)::value;
        const bool valid = ODD ? valid_o : act;
        const bool edge = ODD ? edge_hi : edge_lo;
        float* const ue = ODD ? ue_o : ue_e;
        float* const ne = ODD ? ne_o : ne_e;
        if (edge && !(first && !ODD)) {   // the column the neighbour CTA pushed in the previous step
          const uint32_t mb = ODD ? mb_hi : mb_lo;
          unsigned& ph = ODD ? ph_hi : ph_lo;
          mbar_wait(mb, ph & 1u);
          ++ph;
          if (gl == 0) mbar_arm(mb, tx_bytes);
        }
        unsigned long long g01 = 0ull, g23 = 0ull;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          oth[c] = *reinterpret_cast<const float4*>(ue + 4 * LP * c);
          dot4acc(cur[c], oth[c], g01, g23);
        }
        const float be = *ne;
        const float ga = group_sum<LP>(fold(g01, g23));
        const float al = na;
        Rot r;
        bool rot;
        rotation_fast(al, be, ga, tol2, valid, r, rot);
#pragma unroll
        for (int c = 0; c < CH; ++c) rotate4(r, cur[c], oth[c]);
        const float nal = fmaxf(fmaf(-r.t, ga, al), 0.f);
        na = fmaf(r.t, ga, be);        // the rotated partner is this group's column from here on
        rot_count += rot ? 1u : 0u;    // (every lane counts; lane 0 of the group reports)
        if (valid) {
          // swap: the rotated former own column goes to the even position (in the CTA of its next reader)
          if (edge) {
            const uint32_t dstF = ODD ? dstF_o : dstF_e, dstN = ODD ? dstN_o : dstN_e;
            const uint32_t mb = ODD ? mb_peer_hi : mb_peer_lo;
#pragma unroll
            for (int c = 0; c < CH; ++c) st_async_f4(dstF + (uint32_t)(16 * LP * c), cur[c], mb);
            if (gl == 0) st_async_f1(dstN, nal, mb);
          } else {   // the reader is a group of this CTA: plain shared-memory stores
#pragma unroll
            for (int c = 0; c < CH; ++c) *reinterpret_cast<float4*>(ue + 4 * LP * c) = cur[c];
            if (gl == 0) *ne = nal;
          }
        }
        if (!ODD && is_last) {   // mirror of the new own column for the partner-less odd step
#pragma unroll
          for (int c = 0; c < CH; ++c) *reinterpret_cast<float4*>(ue_o + 4 * LP * c) = oth[c];
          if (gl == 0) *ne_o = na;
        }
        __syncthreads();
      };
      float4 bv[CH];
      for (int s2 = 0; s2 < (D >> 1); ++s2) {
        step(std::false_type{}, s2 == 0, av, bv);
        step(std::true_type{}, false, bv, av);
      }
    }
    if (edge_lo) {   // the last odd step's push from below (the final column at this CTA's first even position)
      mbar_wait(mb_lo, ph_lo & 1u);
      ++ph_lo;
      if (gl == 0) mbar_arm(mb_lo, tx_bytes);
    }
    block_barrier();   // every CTA is out of the step loop before the columns are broadcast
    ++sweeps;
    // ---- every CTA gets every column: odd positions from the registers, even ones from the owner
    if (act) {
      float* uo = F + (size_t)(2 * g + 1) * ld + 4 * gl;
      const float* ue = F + (size_t)(2 * g) * ld + 4 * gl;
      float4 ev[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        *reinterpret_cast<float4*>(uo + 4 * LP * c) = av[c];
        ev[c] = *reinterpret_cast<const float4*>(ue + 4 * LP * c);
      }
      const float ne = nrm2[2 * g];
      if (gl == 0) nrm2[2 * g + 1] = na;
#pragma unroll
      for (int r = 0; r < MAXNC; ++r) {
        if (CL && r < nc && r != (int)rank) {
          const uint32_t off_o = (uint32_t)(((size_t)(2 * g + 1) * ld + 4 * gl) * sizeof(float));
          const uint32_t off_e = (uint32_t)(((size_t)(2 * g) * ld + 4 * gl) * sizeof(float));
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            st_cluster_f4(Fr[r] + off_o + (uint32_t)(16 * LP * c), av[c]);
            st_cluster_f4(Fr[r] + off_e + (uint32_t)(16 * LP * c), ev[c]);
          }
          if (gl == 0) {
            st_cluster_f1(Nr[r] + (uint32_t)((2 * g + 1) * sizeof(float)), na);
            st_cluster_f1(Nr[r] + (uint32_t)((2 * g) * sizeof(float)), ne);
          }
        }
      }
    }
    block_barrier();
    t_sweep_cycles += clock64() - t_s0;

    // ---- check of all D(D-1)/2 cosines (this CTA: rows pp = rank, rank + nc, ...), fix-up pass ---
    for (int fixrounds = 0;; ++fixrounds) {
      const long long t_c0 = clock64();
      float cmax = 0.f;
      const unsigned gmask = (LP >= 32) ? 0xffffffffu : (((1u << LP) - 1u) << (lane & ~(LP - 1)));
      for (int pp = (int)rank + nc * lg; pp < ng; pp += nc * lgroups) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {   // rows pp and D-1-pp together: balanced work
          const int p = half ? D - 1 - pp : pp;
          const float* up = F + (size_t)p * ld + 4 * gl;
          float4 pv[CH];
#pragma unroll
          for (int c = 0; c < CH; ++c) pv[c] = *reinterpret_cast<const float4*>(up + 4 * LP * c);
          const float al = nrm2[p];
          // LP columns q per pass: every lane accumulates its partial dot product with each of them, then a
          // transposing butterfly (LP - 1 shuffles for LP sums instead of LP log2 LP) leaves the complete
          // dot product with column q0 + gl in lane gl -- independent chains, nothing serialised on a reduction
#pragma unroll 1
          for (int q0 = p + 1; q0 < D; q0 += LP) {
            float v[LP];
#pragma unroll
            for (int j = 0; j < LP; ++j) {
              const int qj = min(q0 + j, D - 1);
              const float* uq = F + (size_t)qj * ld + 4 * gl;
              unsigned long long g01 = 0ull, g23 = 0ull;
#pragma unroll
              for (int c = 0; c < CH; ++c) dot4acc(pv[c], *reinterpret_cast<const float4*>(uq + 4 * LP * c), g01, g23);
              v[j] = fold(g01, g23);
            }
#pragma unroll
            for (int o = LP / 2; o >= 1; o >>= 1) {
              const bool upper = (gl & o) != 0;
#pragma unroll
              for (int i = 0; i < o; ++i) {
                const float send = upper ? v[i] : v[i + o];
                const float keep = upper ? v[i + o] : v[i];
                v[i] = keep + __shfl_xor_sync(gmask, send, o);
              }
            }
            const int q = q0 + gl;
            if (q < D) {
              const float ga = v[0];
              const float den = al * nrm2[q];
              if (ga * ga > tol2 * den) {
                cmax = fmaxf(cmax, ga * ga / den);
                const unsigned slot = atomicAdd(&s_nfix, 1u);
                if (slot < (unsigned)MAXFIX) s_fix[slot] = ((unsigned)p << 16) | (unsigned)q;
              }
            }
          }
        }
      }
      if (cmax > 0.f) atomicMax(&s_flag, __float_as_uint(cmax));
      __syncthreads();
      // exchange {count, worst squared cosine, list} with every CTA of the cluster
      {
        const unsigned cnt = s_nfix, flg = s_flag;
        for (int i = tid; i < nc * (MAXFIX + 2); i += nthreads) {
          const int r = i / (MAXFIX + 2), j = i - r * (MAXFIX + 2);
          if (CL && nc > 1) {
            if (j == 0) st_cluster_u1(map_rank(smem_u32(&x_cnt[xpar][rank]), (uint32_t)r), cnt);
            else if (j == 1) st_cluster_u1(map_rank(smem_u32(&x_flag[xpar][rank]), (uint32_t)r), flg);
            else if ((unsigned)(j - 2) < min(cnt, (unsigned)MAXFIX))
              st_cluster_u1(map_rank(smem_u32(&x_fix[xpar][rank][j - 2]), (uint32_t)r), s_fix[j - 2]);
          } else {
            if (j == 0) x_cnt[xpar][0] = cnt;
            else if (j == 1) x_flag[xpar][0] = flg;
            else if ((unsigned)(j - 2) < min(cnt, (unsigned)MAXFIX)) x_fix[xpar][0][j - 2] = s_fix[j - 2];
          }
        }
      }
      block_barrier();
      if (tid == 0) { s_flag = 0u; s_nfix = 0u; }
      unsigned total = 0, worstb = 0;
      bool overflow = false;
      for (int r = 0; r < nc; ++r) {
        const unsigned c = x_cnt[xpar][r];
        overflow |= c > (unsigned)MAXFIX;
        total += c;
        worstb = max(worstb, x_flag[xpar][r]);   // non-negative floats order like their bit patterns
      }
      const float worst2 = __uint_as_float(worstb);
      const long long t_c1 = clock64();
      t_check_cycles += t_c1 - t_c0;
      const int xp = xpar;
      xpar ^= 1;
      if (worstb == 0u) { converged = true; break; }
      st_maxlist = max(st_maxlist, (int)total);
      if (overflow || total > (unsigned)MAXFIX || fixrounds >= 2) { st_over += (fixrounds >= 2 && !overflow && total <= (unsigned)MAXFIX) ? 100 : 1; break; }   // full sweep
      st_fix += (int)total;
      // merged list in ascending (p, q) order: the result must not depend on who found what first
      for (unsigned ei = tid; ei < total; ei += nthreads) {
        unsigned mine = 0, acc = 0;
        for (int r = 0; r < nc; ++r) {
          const unsigned c = x_cnt[xp][r];
          if (ei >= acc && ei < acc + c) mine = x_fix[xp][r][ei - acc];
          acc += c;
        }
        unsigned rk = 0;
        for (int r = 0; r < nc; ++r)
          for (unsigned j = 0; j < x_cnt[xp][r]; ++j) rk += (x_fix[xp][r][j] < mine) ? 1u : 0u;
        s_fixs[rk] = mine;
      }
      __syncthreads();
      // Fix-up rotations in PARALLEL rounds with the result of the sequential pass: every pending pair claims its two
      // columns with its list index (atomicMin); a pair that holds both rotates now, the others wait.  Two pairs that
      // share a column therefore run in list order, disjoint rotations commute exactly -- same numbers as one
      // pair after the other, in (longest chain through a shared column) rounds instead of `total`.  Every CTA of
      // the cluster does this on its own copy (identical lists, identical arithmetic).
      for (;;) {
        for (int i = tid; i < D; i += nthreads) s_owner[i] = 0xffffffffu;
        if (tid == 0) s_left = 0u;
        __syncthreads();
        for (unsigned e = tid; e < total; e += nthreads) {
          const unsigned pq = s_fixs[e];
          if (!(pq & 0x80000000u)) {
            atomicMin(&s_owner[pq >> 16], e);
            atomicMin(&s_owner[pq & 0xffffu], e);
          }
        }
        __syncthreads();
        for (unsigned e = lg; e < total; e += lgroups) {
          const unsigned pq = s_fixs[e];
          if (pq & 0x80000000u) continue;
          const int p = (int)(pq >> 16), q = (int)(pq & 0xffffu);
          if (s_owner[p] != e || s_owner[q] != e) {
            if (gl == 0) s_left = 1u;
            continue;
          }
          float* up = F + (size_t)p * ld + 4 * gl;
          float* uq = F + (size_t)q * ld + 4 * gl;
          float4 pa[CH], pb[CH];
          unsigned long long g01 = 0ull, g23 = 0ull;
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            pa[c] = *reinterpret_cast<const float4*>(up + 4 * LP * c);
            pb[c] = *reinterpret_cast<const float4*>(uq + 4 * LP * c);
            dot4acc(pa[c], pb[c], g01, g23);
          }
          const float ga = group_sum_masked<LP>(fold(g01, g23), gmask);
          const float al = nrm2[p], be = nrm2[q];
          Rot r;
          if (rotation(al, be, ga, tol2, r)) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              rotate4(r, pa[c], pb[c]);
              *reinterpret_cast<float4*>(up + 4 * LP * c) = pa[c];
              *reinterpret_cast<float4*>(uq + 4 * LP * c) = pb[c];
            }
            if (gl == 0) {
              nrm2[p] = fmaxf(fmaf(-r.t, ga, al), 0.f);
              nrm2[q] = fmaf(r.t, ga, be);
              ++rot_count;
            }
          }
          if (gl == 0) s_fixs[e] = pq | 0x80000000u;
        }
        __syncthreads();
        const bool more = s_left != 0u;
        ++st_rounds;
        __syncthreads();
        if (!more) break;
      }
      t_fix_cycles += clock64() - t_c1;
      // every listed |cos| < 1e-3: the untouched pairs moved by theta * tol at most, no re-check (eig_small.cu)
      if (worst2 < 1e-6f) { converged = true; break; }
    }
  }
  const long long t_tail0 = clock64();

  // ---- column norms (every CTA, all columns: the validation sum), eigenvectors of this CTA's share, one pass:
  // a warp per column, one or two float4 per lane (D % 4 == 0 and D <= 200 on this path), normalised rows stored
  // with 16-byte accesses
  float wsum_part = 0.f;
  float* Vb = a.Vt + (size_t)b * D * D;
  const int nq = D >> 2;
  for (int col = warp; col < D; col += nwarps) {
    const float4* u4 = reinterpret_cast<const float4*>(F + (size_t)col * ld);
    float4 v[2];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int q = lane + 32 * j;
      v[j] = (q < nq) ? u4[q] : make_float4(0.f, 0.f, 0.f, 0.f);
      s = fmaf(v[j].x, v[j].x, fmaf(v[j].y, v[j].y, fmaf(v[j].z, v[j].z, fmaf(v[j].w, v[j].w, s))));
    }
    s = warp_sum(s);
    const float nrm = sqrtf(s);
    const float inv = (nrm > 0.f) ? 1.f / nrm : 0.f;
    if (nc == 1 || (col % nc) == (int)rank) {
      float4* o4 = reinterpret_cast<float4*>(Vb + (size_t)col * D);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int q = lane + 32 * j;
        if (q < nq) o4[q] = make_float4(v[j].x * inv, v[j].y * inv, v[j].z * inv, v[j].w * inv);
      }
      if (lane == 0) a.w[(size_t)b * D + col] = nrm - sigma;
    }
    if (lane == 0) wsum_part += nrm;
  }
  const float wsum = block_sum(wsum_part, red);
  const float expect = trace + (float)D * sigma;
  const bool pd = fabsf(wsum - expect) <= 2e-3f * fabsf(expect);
  if (a.work != nullptr) {
    if (gl == 0 && rot_count) atomicAdd(&s_rot, rot_count);
    __syncthreads();
    if (tid == 0) {
      // dot products: the sweeps + every check pass (counted once per graph), rotations of this CTA's groups
      if (rank == 0) atomicAdd(a.work, (unsigned long long)(2 * sweeps) * (unsigned long long)(D * (D - 1) / 2));
      atomicAdd(a.work + 1, (unsigned long long)s_rot);
    }
  }
  if (a.info && tid == 0 && rank == 0) {
    float* o = a.info + (size_t)b * 4;
    o[0] = (float)(sweeps + (pd ? 0 : 1000));   // >= 1000: the launcher's retry pass redoes this graph from scratch
    o[1] = sigma;
    o[2] = trace;
    o[3] = wsum - (float)D * sigma;
    if (a.timing == 2 && pd) {
      o[1] = (float)st_fix;
      o[2] = (float)(st_over + 1000 * st_rounds);   // + 1000 x parallel fix-up rounds
      o[3] = (float)st_maxlist;
    }
    if (a.timing == 3 && pd) {
      o[1] = (float)(t_load_end - t_start);
      o[2] = (float)t_check_cycles;
      o[3] = (float)t_fix_cycles;
    }
    if (a.timing == 1 && pd) {
      o[1] = (float)(t_tail0 - t_start - t_sweep_cycles);   // load + checks + fix-ups
      o[2] = (float)t_sweep_cycles;
      o[3] = (float)(clock64() - t_tail0);
    }
  }
  if (CL && nc > 1) { __syncwarp(); cluster_arrive(); cluster_wait(); }   // no CTA leaves while a peer may still push into its memory
}

static int g_tune_cluster = -1;   // -1 auto; 0 off (one-CTA kernel only); 1 / 2 / 4: CTAs per graph of the odd-even kernel
int eig_cluster_tune(const char* key, int value) {
  if (!strcmp(key, "eig_cluster")) { g_tune_cluster = value; return 0; }
  return 1;
}

template <int LP, int CH, int MINB, int MAXT = 512>
static int launch_oe_cfg(const EigArgs& a, int B, int nc, int gpc, int threads, size_t smem, cudaStream_t st) {
  static bool attr_set[16] = {false};
  int dev = 0;
  UGLAD_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16 || !attr_set[dev]) {
    UGLAD_CUDA(cudaFuncSetAttribute(eig_jacobi_oe_kernel<LP, CH, MINB, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
    if (dev >= 0 && dev < 16) attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * nc), 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)nc;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (nc > 1) ? 1 : 0;
  profile_begin(st, 0, (double)B * (4.0 * a.D * a.D + 3.0 * a.D) * 4.0);
  cudaError_t e = cudaLaunchKernelEx(&cfg, eig_jacobi_oe_kernel<LP, CH, MINB, MAXT>, a, nc, gpc);
  profile_end(st);
  if (e != cudaSuccess) {
    set_error("eig_jacobi_oe_kernel: launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  UGLAD_CHECK_LAUNCH("eig_jacobi_oe_kernel");
  return 0;
}

// CTAs per graph for a batch of B: as many as keep every cluster resident at once; full batches run the same
// schedule with one CTA per graph (one column read + written per pair and step, the butterfly check)
int eig_cluster_size(int B, int D) {
  if (g_tune_cluster == 0) return 0;
  if (D % 4 != 0 || D < 16 || D > 200) return 0;
  if (g_tune_cluster > 0) return g_tune_cluster;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // clusters of 4 strand some SMs (GPC sizes 16 / 18 / 20): leave a margin instead of a second wave
  if (B * 4 <= sms - 16) return 4;
  // (clusters of 2 -- 16 lanes per pair to stay within 512 threads -- are available through the knob and tested, but
  // were not re-measured after the sweep rewrites, where 16-lane groups fell behind: batches of 34 .. 74 graphs take the
  // one-CTA form, one CTA per SM)
  return 1;
}

// warm solve (a.U0 set) with nc CTAs per graph; returns 2 when the shape has no configuration (caller falls back)
int launch_eig_cluster(const EigArgs& a_in, int B, int nc, cudaStream_t st) {
  EigArgs a = a_in;
  if (a.U0 == nullptr || a.info == nullptr || a.D % 2 != 0 || nc < 1 || nc > oe::MAXNC || nc == 3) return 2;
  const int D = a.D, ng = D / 2, gpc = (ng + nc - 1) / nc;
  int lp = 32;
  while (lp > 8 && (gpc * lp > 512 || lp * 4 >= 2 * D)) lp >>= 1;   // narrow groups for small D / many groups per CTA
  if (nc == 1) lp = (D <= 128) ? 8 : 16;                             // full batches: two CTAs per SM, as eig_small.cu
  int ch = (D + 4 * lp - 1) / (4 * lp);
  int chT = ch <= 1 ? 1 : (ch <= 2 ? 2 : (ch <= 4 ? 4 : 8));
  if (chT > 4) { lp = 32; ch = (D + 127) / 128; chT = ch <= 1 ? 1 : (ch <= 2 ? 2 : 4); if (ch > 4) return 2; }
  int threads = ((gpc * lp + 31) / 32) * 32;
  if (threads < 64) threads = 64;
  if (threads > oe_max_threads(lp, chT)) return 2;
  a.ld = lp * chT * 4 + 4;
  const int Dp = (D + 3) & ~3;
  const size_t smem = ((size_t)a.ld * (D + 1) + 2 * Dp + 4 + 32) * sizeof(float);
  if (smem > 216 * 1024) return 2;
  a.timing = eig_small_timing();
  a.work = profile_eig_counters();
#define UGLAD_OE_CASE(LP_, CH_)                                                                    \
  if (lp == LP_ && chT == CH_)                                                                    \
    return nc == 1 ? (threads <= 448 ? launch_oe_cfg<LP_, CH_, 2, 448>(a, B, nc, gpc, threads, smem, st)    \
                                     : launch_oe_cfg<LP_, CH_, 2>(a, B, nc, gpc, threads, smem, st))        \
                   : launch_oe_cfg<LP_, CH_, 1>(a, B, nc, gpc, threads, smem, st)
  UGLAD_OE_CASE(8, 1); UGLAD_OE_CASE(8, 2); UGLAD_OE_CASE(8, 4);
  UGLAD_OE_CASE(16, 1); UGLAD_OE_CASE(16, 2); UGLAD_OE_CASE(16, 4);
  UGLAD_OE_CASE(32, 1); UGLAD_OE_CASE(32, 2);
#undef UGLAD_OE_CASE
  return 2;
}

}  // namespace uglad
