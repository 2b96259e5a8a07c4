// Batched 3xTF32 GEMM on the 5th-generation tensor cores (sm_100a): C = epi(A * B^T).
//
// Every operand is used as hi + lo with hi a TF32 number and lo = x - hi (exact), so x = hi + lo
// carries full FP32 precision and
//   A B^T  ~=  A_hi B_hi^T + A_hi B_lo^T + A_lo B_hi^T          (error ~2^-21 |A||B|, FP32 accumulate)
// which is what the 1e-4 parity budget on theta after 15 unrolled layers needs (single-pass TF32
// is 1e-3).  Both operands are K-major (row-major [rows][K]); the callers only ever multiply by
// symmetric matrices or provide the transpose explicitly, so B^T costs nothing.
//
// Two MMAs per K granule instead of three: B_hi and B_lo sit next to each other in the stage, so
//   acc[:, 0:2BN] += A_hi * [B_hi ; B_lo]^T   (one instruction, N = 2 BN)
//   acc[:, 0:BN]  += A_lo * B_hi^T
// and the epilogue adds the two column halves.  A K = 8 TF32 instruction with both operands in
// shared memory is bound by the 4 KB A-operand read (~65-78 cycles measured for N = 64 ... 128), so
// the wide instruction costs no more than a narrow one until N/2 cycles of tensor work exceed it.
//
// Operand forms:
//   pre-split (RAW = false): hi = tf32(x) (round to nearest) and lo are two FP32 arrays written by the
//              producer (eigenvector products of the D <= 166 path); TMA loads all four slabs.
//   raw (RAW = true; lo == nullptr; the Newton-Schulz chain and the Cholesky helpers): plain FP32
//              matrices.  TMA brings ONE 4-byte word per element into the hi slot; the tensor core
//              reads a 32-bit container as TF32 by dropping the low 13 mantissa bits, so the raw
//              word IS the hi operand (hi = trunc(x)), and eight split warps only write
//              lo = x - trunc(x) (exact) into the neighbouring slot behind fence.proxy.async.  Half
//              the L2->SM bytes, and every producer of the chain writes 4 bytes per element.
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0       : TMA producer -- cp.async.bulk.tensor.3d of the 32-float-wide K slabs (A: 128 rows,
//                  B: BN rows), 128B swizzle, mbarrier complete_tx, 3-4 stages
//   warp 1       : MMA issuer   -- one elected lane (elect.sync: no per-instruction election loops)
//                  issues tcgen05.mma.kind::tf32 (M = 128, K = 8) with base + immediate descriptors,
//                  accumulating in TMEM; tcgen05.commit frees the slab
//   epilogue     : 8 warps (pre-split) / 4 warps (raw) -- tcgen05.ld of the accumulator (two TMEM
//                  stages, so the next tile's MMAs overlap), alpha/beta/diag epilogue, optional
//                  hi/lo split output; the 8-warp form coalesces its stores through a shared-memory patch
//   split (raw)  : 8 warps in two groups taking alternate slabs
// A launch carries one product or two independent products of the same shape (two sets of tensor
// maps and epilogue parameters): the second doubles the tiles the persistent grid can spread over
// 148 SMs.  Launches are chained with programmatic dependent launch: the prologue (barriers, TMEM
// allocation, descriptor prefetch) runs under the tail of the previous kernel, griddepcontrol.wait
// guards the first access to its results.
#include <cuda.h>
#include <string.h>
#include <mutex>
#include <unordered_map>
#include "kernels.cuh"

namespace uglad {

namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;                 // floats per K slab = one 128-byte swizzle row
constexpr int A_BYTES = BM * BK * 4;   // 16 KB
// epilogue warps: two per TMEM lane quarter for pre-split operands (a lone warp per scheduler is
// latency-bound); the raw-operand kernel spends its register budget on eight split warps instead
// (measured: 4 + 8 beats 8 + 4 and 8 + 8 there)
constexpr int EPI_WARPS_SPLIT = 8, EPI_WARPS_RAW = 4;
constexpr int THREADS = 64 + 32 * EPI_WARPS_SPLIT;   // pre-split operands: TMA warp, MMA warp, epilogue warps
constexpr int SPLIT_WARPS = 8;     // raw operands: + 8 warps that split the slabs in shared memory
constexpr int THREADS_RAW = 64 + 32 * EPI_WARPS_RAW + 32 * SPLIT_WARPS;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one leader lane of a converged warp; ptxas then treats the guarded region as single-threaded and
// issues the tcgen05 / TMA instructions without a per-instruction election loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t hb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
  return __uint_as_float(hb);
}
// the part of x that a TF32 read of its 32-bit container drops: lo = x - trunc_tf32(x) (exact)
__device__ __forceinline__ float lo1(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ float4 lo4(const float4& x) { return make_float4(lo1(x.x), lo1(x.y), lo1(x.z), lo1(x.w)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  for (int spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"   // (a suspend-time hint was measured:
        "selp.u32 %0, 1, 0, p;\n\t}"                                    // slower wake-ups, no gain in the chain)
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && (spin & 1023) == 1023) {  // a lost arrival must fault, never hang the GPU
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// one [32 cols x 128 rows] box from shared memory (128B swizzle) to global; rows/columns beyond the
// tensor are clipped by the TMA unit
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128B-swizzled operand slab: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                   // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                   // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// the two column halves of one accumulator chunk (columns c .. c+31 and c+off .. c+off+31), one wait
__device__ __forceinline__ void tmem_ld32x2(uint32_t taddr, uint32_t off, uint32_t* r, uint32_t* r2) {
  tmem_ld32_nowait(taddr, r);
  tmem_ld32_nowait(taddr + off, r2);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 32 consecutive 32-bit columns of this thread's TMEM lane (the warp owns lanes 32 (warp % 4) ...)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand (128 lanes x 8 columns of 32-bit TF32 containers) comes from
// tensor memory, so it costs no shared-memory bandwidth
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

template <int BN>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = (BN <= 64) ? 4 : 3;
  static constexpr int TMEM_COLS = (BN <= 64) ? 256 : 512;  // two accumulator stages of 2 BN columns ([hi*hi + lo*hi | hi*lo])
  static constexpr int ACC_STRIDE = TMEM_COLS / 2;
  static constexpr int EPI_BYTES = EPI_WARPS_SPLIT * 32 * 20 * 4;   // one 32 x 20 float patch per epilogue warp
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ + EPI_BYTES;
  // raw kernel: two 128 x 32 float staging tiles (128B swizzle, 1024-byte aligned) for the TMA stores of C
  static constexpr int CST_BYTES = 2 * BM * 128;
  static constexpr int SMEM_RAW = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 1024 /*barriers*/ + CST_BYTES;
  static_assert(SMEM_RAW <= 232448, "raw-operand kernel exceeds 227 KB of shared memory");
  static constexpr int RAW_TX_BYTES = A_BYTES + B_BYTES;   // raw mode: one word per element arrives by TMA
  // kind::tf32, FP32 accumulate, K-major A and B, M = 128, N = BN
  static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  static constexpr uint32_t IDESC2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);   // N = 2 BN
};

}  // namespace tc

struct TcEpi {
  float alpha, beta, diag;
  const float* alpha_dev;
  const float* E1_hi;
  const float* E1_lo;   // null: E1 is a plain FP32 matrix
  long long sE1;
  int lde1;
  float* C_hi;
  float* C_lo;          // null: C is written as plain FP32
  long long sC;
  int ldc;
};
struct TcParams {
  long long* dbg;   // developer timeline: [grid][16] clock64 stamps and per-role phase sums (null = off)
  int M, N, K, batch, tiles_m, tiles_n;
  int nprob;        // 1, or 2: two independent products of the same shape share one launch
  int exp;          // developer experiment (wrong results): 1 = pre-split kernel skips its TMA loads
  int tma_c[2];     // raw kernel: C of problem i leaves through TMA stores (plain output, 16-byte aligned rows)
  TcEpi e[2];
};

// epilogue of one tile by the four epilogue warps (warp quarter q owns TMEM lanes 32q .. 32q+31 =
// tile rows): C = alpha acc + beta E1 + diag I, optional hi/lo split, vectorised stores
constexpr int EPI_LD = 20;   // floats per row of an epilogue warp's 32 x 16 patch (16-byte aligned rows, conflict-free row writes)
template <int BN, int EPI_WARPS>
__device__ __forceinline__ void tc_epilogue_tile(const TcParams& p, const TcEpi& ep, uint32_t tmem_acc, int tm, int tn,
                                                 int b, int q, int half, int lane, float* stg, long long* dbg) {
  long long t_ld = 0, t_rest = 0;
  const bool vec_c = (ep.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.C_hi) & 15) == 0) &&
                     (ep.C_lo == nullptr || (reinterpret_cast<uintptr_t>(ep.C_lo) & 15) == 0) && (ep.sC % 4 == 0);
  const bool vec_e = ep.E1_hi != nullptr && (ep.lde1 % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.E1_hi) & 15) == 0) &&
                     (ep.E1_lo == nullptr || (reinterpret_cast<uintptr_t>(ep.E1_lo) & 15) == 0) && (ep.sE1 % 4 == 0);
  const int row = tm * tc::BM + q * 32 + lane;
  const float alpha = ep.alpha_dev ? ep.alpha * ep.alpha_dev[b] : ep.alpha;
  const float beta = ep.beta, diag = ep.diag;
  float* Ch = ep.C_hi + (size_t)b * ep.sC + (size_t)row * ep.ldc;
  float* Cl = ep.C_lo ? ep.C_lo + (size_t)b * ep.sC + (size_t)row * ep.ldc : nullptr;
  const float* Eh = ep.E1_hi ? ep.E1_hi + (size_t)b * ep.sE1 + (size_t)row * ep.lde1 : nullptr;
  const float* El = ep.E1_lo ? ep.E1_lo + (size_t)b * ep.sE1 + (size_t)row * ep.lde1 : nullptr;
  const int row0 = tm * tc::BM + q * 32;   // first row of this warp
  float* Cb = ep.C_hi + (size_t)b * ep.sC;
  float* Clb = ep.C_lo ? ep.C_lo + (size_t)b * ep.sC : nullptr;
#pragma unroll 1
  for (int c0 = 32 * half; c0 < BN; c0 += 32 * EPI_WARPS / 4) {   // the warps of a quarter interleave the chunks
    const int colbase = tn * BN + c0;
    if (colbase >= p.N) break;
    uint32_t v[32], v2[32];
    const long long e0 = dbg ? clock64() : 0;
    // interior chunk (the common case): 32 full columns inside the tile and the matrix, 16-byte
    // aligned rows.  One lean instruction stream per 4-column group; the generic path below
    // handles ragged edges element by element.
    // (with one warp per quarter the direct-store path below measured faster than the patch)
    const bool fast = EPI_WARPS > 4 && c0 + 32 <= BN && colbase + 32 <= p.N && vec_c && (Eh == nullptr || vec_e);
    if (fast) {
      tc::tmem_ld32x2(tmem_acc + ((uint32_t)(q * 32) << 16) + c0, BN, v, v2);
      const long long e1 = dbg ? clock64() : 0;
      t_ld += e1 - e0;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));   // frees v2
      float4 e[8];
      if (Eh && row < p.M) {
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = *reinterpret_cast<const float4*>(Eh + colbase + 4 * j);
        if (El) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 l = *reinterpret_cast<const float4*>(El + colbase + 4 * j);
            e[j].x += l.x; e[j].y += l.y; e[j].z += l.z; e[j].w += l.w;
          }
        }
      }
      // tcgen05.ld hands every thread 32 consecutive columns of ITS row, so direct stores touch 32
      // different lines per instruction and the LSU retires one line per cycle (measured: 4096
      // cycles for a 128 x 128 tile).  Each 32 x 16 half chunk goes through a padded shared-memory
      // patch instead and leaves as eight 64-byte row segments per instruction.
      float* mine = stg + lane * EPI_LD;
      const int cr = lane >> 2, cc = (lane & 3) * 4;   // coalesced pass: rows cr + 8 i, columns cc .. cc + 3
      const int dcol = row - colbase;                  // this row's diagonal column inside the chunk, if any
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const int j = 4 * hh + j4;
          float4 o;
          o.x = alpha * __uint_as_float(v[4 * j + 0]);
          o.y = alpha * __uint_as_float(v[4 * j + 1]);
          o.z = alpha * __uint_as_float(v[4 * j + 2]);
          o.w = alpha * __uint_as_float(v[4 * j + 3]);
          if (Eh && row < p.M) {
            o.x = fmaf(beta, e[j].x, o.x); o.y = fmaf(beta, e[j].y, o.y);
            o.z = fmaf(beta, e[j].z, o.z); o.w = fmaf(beta, e[j].w, o.w);
          }
          *reinterpret_cast<float4*>(mine + 4 * j4) = o;
        }
        // diagonal term: one read-modify-write of the thread's own patch row (cheaper than a select
        // per element: the epilogue warps are latency-bound, every instruction counts)
        if (diag != 0.f && dcol >= 16 * hh && dcol < 16 * hh + 16) mine[dcol - 16 * hh] += diag;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = cr + 8 * i;
          if (row0 + r < p.M) {
            const float4 o = *reinterpret_cast<const float4*>(stg + r * EPI_LD + cc);
            const size_t off = (size_t)(row0 + r) * ep.ldc + colbase + 16 * hh + cc;
            if (Clb) {
              float4 h, l;
              h.x = tc::rna_tf32(o.x); l.x = o.x - h.x; h.y = tc::rna_tf32(o.y); l.y = o.y - h.y;
              h.z = tc::rna_tf32(o.z); l.z = o.z - h.z; h.w = tc::rna_tf32(o.w); l.w = o.w - h.w;
              *reinterpret_cast<float4*>(Cb + off) = h;
              *reinterpret_cast<float4*>(Clb + off) = l;
            } else {
              *reinterpret_cast<float4*>(Cb + off) = o;
            }
          }
        }
        __syncwarp();
      }
      if (dbg) t_rest += clock64() - e1;
      continue;
    }
    tc::tmem_ld32x2(tmem_acc + ((uint32_t)(q * 32) << 16) + c0, BN, v, v2);
    const long long e1 = dbg ? clock64() : 0;
    t_ld += e1 - e0;
    if (row < p.M) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int col = tn * BN + c0 + j;
        if (c0 + j >= BN || col >= p.N) break;   // BN = 112: the last 32-column read overhangs the tile
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = alpha * (__uint_as_float(v[j + e]) + __uint_as_float(v2[j + e]));
        const bool full4 = col + 3 < p.N;
        if (Eh) {
          if (full4 && vec_e) {
            const float4 eh = *reinterpret_cast<const float4*>(Eh + col);
            o[0] = fmaf(beta, eh.x, o[0]); o[1] = fmaf(beta, eh.y, o[1]);
            o[2] = fmaf(beta, eh.z, o[2]); o[3] = fmaf(beta, eh.w, o[3]);
            if (El) {
              const float4 el = *reinterpret_cast<const float4*>(El + col);
              o[0] = fmaf(beta, el.x, o[0]); o[1] = fmaf(beta, el.y, o[1]);
              o[2] = fmaf(beta, el.z, o[2]); o[3] = fmaf(beta, el.w, o[3]);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (col + e < p.N) {
                float ev = Eh[col + e];
                if (El) ev += El[col + e];
                o[e] = fmaf(beta, ev, o[e]);
              }
          }
        }
        if (diag != 0.f) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (col + e == row) o[e] += diag;
        }
        if (Cl) {
          float h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            uint32_t hb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(o[e]));
            h[e] = __uint_as_float(hb);
            l[e] = o[e] - h[e];
          }
          if (full4 && vec_c) {
            *reinterpret_cast<float4*>(Ch + col) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(Cl + col) = make_float4(l[0], l[1], l[2], l[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (col + e < p.N) { Ch[col + e] = h[e]; Cl[col + e] = l[e]; }
          }
        } else {
          if (full4 && vec_c) {
            *reinterpret_cast<float4*>(Ch + col) = make_float4(o[0], o[1], o[2], o[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (col + e < p.N) Ch[col + e] = o[e];
          }
        }
      }
    }
    if (dbg) t_rest += clock64() - e1;
  }
  if (dbg && q == 2 && lane == 0) { dbg[14] = t_ld; dbg[15] = t_rest; }
}

// Epilogue of the raw-operand kernel through TMA stores: tcgen05.ld hands every thread 32 consecutive
// columns of ITS row, and direct stores then touch 32 different lines per instruction (the LSU
// retires about one line per cycle: 4096 cycles for a 128 x 128 tile, as long as a short-K main
// loop).  The four epilogue warps instead write each 128 x 32 chunk into a 128B-swizzled staging tile
// (conflict-free 16-byte row writes) and one thread hands it to the TMA unit, which writes full
// lines asynchronously and clips the ragged edges; two staging tiles alternate.  The trailing partial
// chunk (BN = 112: 16 columns) goes through a second, dense 16-column box.  Plain FP32 output and
// addend only (the host checks).
template <int BN>
__device__ __forceinline__ void tc_epilogue_tile_tma(const TcParams& p, const TcEpi& ep, const CUtensorMap* mC,
                                                     const CUtensorMap* mCp, uint32_t tmem_acc, int tm, int tn, int b, int q, int lane,
                                                     uint8_t* cst, uint32_t cst_s, int& buf, bool leader) {
  const int rit = q * 32 + lane;                    // row inside the tile
  const int row = tm * tc::BM + rit;
  const float alpha = ep.alpha_dev ? ep.alpha * ep.alpha_dev[b] : ep.alpha;
  const float beta = ep.beta, diag = ep.diag;
  const float* Eh = ep.E1_hi ? ep.E1_hi + (size_t)b * ep.sE1 + (size_t)row * ep.lde1 : nullptr;
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    const int colbase = tn * BN + c0;
    if (colbase >= p.N) break;                       // uniform
    uint32_t v[32], v2[32];
    tc::tmem_ld32x2(tmem_acc + ((uint32_t)(q * 32) << 16) + c0, BN, v, v2);
    const bool full = c0 + 32 <= BN;                 // uniform
    const int dcol = row - colbase;                  // this row's diagonal column inside the chunk, if any
    if (full) {
      float4 o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j].x = alpha * (__uint_as_float(v[4 * j + 0]) + __uint_as_float(v2[4 * j + 0]));
        o[j].y = alpha * (__uint_as_float(v[4 * j + 1]) + __uint_as_float(v2[4 * j + 1]));
        o[j].z = alpha * (__uint_as_float(v[4 * j + 2]) + __uint_as_float(v2[4 * j + 2]));
        o[j].w = alpha * (__uint_as_float(v[4 * j + 3]) + __uint_as_float(v2[4 * j + 3]));
      }
      if (Eh && row < p.M) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = colbase + 4 * j;
          if (col + 3 < p.N) {
            const float4 e = *reinterpret_cast<const float4*>(Eh + col);
            o[j].x = fmaf(beta, e.x, o[j].x); o[j].y = fmaf(beta, e.y, o[j].y);
            o[j].z = fmaf(beta, e.z, o[j].z); o[j].w = fmaf(beta, e.w, o[j].w);
          } else {
            if (col < p.N) o[j].x = fmaf(beta, Eh[col], o[j].x);
            if (col + 1 < p.N) o[j].y = fmaf(beta, Eh[col + 1], o[j].y);
            if (col + 2 < p.N) o[j].z = fmaf(beta, Eh[col + 2], o[j].z);
          }
        }
      }
      if (leader) tc::tma_store_wait_read<1>();      // the store that read this staging tile two chunks ago is done
      asm volatile("bar.sync 1, 128;" ::: "memory");
      uint8_t* mine = cst + buf * (tc::BM * 128) + rit * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(mine + ((j ^ (rit & 7)) << 4)) = o[j];
      if (diag != 0.f && dcol >= 0 && dcol < 32)     // diagonal term: one read-modify-write of the own row
        *reinterpret_cast<float*>(mine + ((((dcol >> 2) ^ (rit & 7)) << 4) | ((dcol & 3) << 2))) += diag;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> TMA (async proxy) reads
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (leader) {
        tc::tma_store_3d(mC, cst_s + buf * (tc::BM * 128), colbase, tm * tc::BM, b);
        tc::tma_store_commit();
      }
      buf ^= 1;
    } else {                                         // trailing partial chunk (BN % 32 columns): dense unswizzled box
      constexpr int W = BN % 32 ? BN % 32 : 32;
      float4 o[W / 4];
#pragma unroll
      for (int j = 0; j < W / 4; ++j) {
        o[j].x = alpha * (__uint_as_float(v[4 * j + 0]) + __uint_as_float(v2[4 * j + 0]));
        o[j].y = alpha * (__uint_as_float(v[4 * j + 1]) + __uint_as_float(v2[4 * j + 1]));
        o[j].z = alpha * (__uint_as_float(v[4 * j + 2]) + __uint_as_float(v2[4 * j + 2]));
        o[j].w = alpha * (__uint_as_float(v[4 * j + 3]) + __uint_as_float(v2[4 * j + 3]));
      }
      if (Eh && row < p.M) {
#pragma unroll
        for (int j = 0; j < W / 4; ++j) {
          const int col = colbase + 4 * j;
          if (col + 3 < p.N) {
            const float4 e = *reinterpret_cast<const float4*>(Eh + col);
            o[j].x = fmaf(beta, e.x, o[j].x); o[j].y = fmaf(beta, e.y, o[j].y);
            o[j].z = fmaf(beta, e.z, o[j].z); o[j].w = fmaf(beta, e.w, o[j].w);
          } else {
            if (col < p.N) o[j].x = fmaf(beta, Eh[col], o[j].x);
            if (col + 1 < p.N) o[j].y = fmaf(beta, Eh[col + 1], o[j].y);
            if (col + 2 < p.N) o[j].z = fmaf(beta, Eh[col + 2], o[j].z);
          }
        }
      }
      if (leader) tc::tma_store_wait_read<1>();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      uint8_t* mine = cst + buf * (tc::BM * 128) + rit * (W * 4);
#pragma unroll
      for (int j = 0; j < W / 4; ++j) *reinterpret_cast<float4*>(mine + (j << 4)) = o[j];
      if (diag != 0.f && dcol >= 0 && dcol < W) *reinterpret_cast<float*>(mine + (dcol << 2)) += diag;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (leader) {
        tc::tma_store_3d(mCp, cst_s + buf * (tc::BM * 128), colbase, tm * tc::BM, b);
        tc::tma_store_commit();
      }
      buf ^= 1;
    }
  }
}

// the MMAs of one K slab: operand set (Ah | Al | Bh | Bl) OFF bytes behind the slab descriptor d0.
// OFF is a compile-time constant, so every descriptor is d0 + immediate (the 14-bit address field
// cannot carry: shared-memory addresses stay below 256 KB).
template <int BN, int OFF>
__device__ __forceinline__ void tc_issue_slab(uint64_t d0, uint32_t tmem_d, int ngran, bool first_slab) {
  using C = tc::Cfg<BN>;
  constexpr uint64_t oAh = (uint64_t)(OFF >> 4), oAl = (uint64_t)((OFF + tc::A_BYTES) >> 4),
                     oBh = (uint64_t)((OFF + 2 * tc::A_BYTES) >> 4);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (g < ngran) {
      const uint64_t adv = (uint64_t)(g * 2);                 // 32 bytes >> 4 inside the swizzle row
      // [B_hi ; B_lo] is one K-major operand of 2 BN rows (8-row groups 1024 bytes apart throughout)
      tc::umma_tf32(tmem_d, d0 + oAh + adv, d0 + oBh + adv, C::IDESC2, (g != 0 || !first_slab) ? 1u : 0u);
      tc::umma_tf32(tmem_d, d0 + oAl + adv, d0 + oBh + adv, C::IDESC, 1u);
    }
  }
}

// the same with the A operand in tensor memory (ATM kernels): hi at tmem_a, lo 32 columns further
template <int BN, int OFF>
__device__ __forceinline__ void tc_issue_slab_ts(uint64_t d0, uint32_t tmem_d, uint32_t tmem_a, int ngran, bool first_slab) {
  using C = tc::Cfg<BN>;
  constexpr uint64_t oBh = (uint64_t)((OFF + 2 * tc::A_BYTES) >> 4);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (g < ngran) {
      const uint64_t adv = (uint64_t)(g * 2);
      tc::umma_tf32_ts(tmem_d, tmem_a + 8 * g, d0 + oBh + adv, C::IDESC2, (g != 0 || !first_slab) ? 1u : 0u);
      tc::umma_tf32_ts(tmem_d, tmem_a + 32 + 8 * g, d0 + oBh + adv, C::IDESC, 1u);
    }
  }
}

template <int BN, bool RAW, bool ATM = false>
__global__ void __launch_bounds__(RAW ? tc::THREADS_RAW : tc::THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
               const __grid_constant__ CUtensorMap tm2Ah, const __grid_constant__ CUtensorMap tm2Al,
               const __grid_constant__ CUtensorMap tm2Bh, const __grid_constant__ CUtensorMap tm2Bl,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tm2C,
               const __grid_constant__ CUtensorMap tmCp, const __grid_constant__ CUtensorMap tm2Cp, TcParams p) {
  using C = tc::Cfg<BN>;
  constexpr int EPI_WARPS = RAW ? tc::EPI_WARPS_RAW : tc::EPI_WARPS_SPLIT;
  constexpr int ROLE_THREADS = 64 + 32 * EPI_WARPS;   // TMA warp, MMA warp, epilogue warps; the split warps follow
  // ATM: the A operand of the MMAs lives in tensor memory -- per ring stage 32 columns of hi (the raw words) and
  // 32 of lo behind the two accumulator stages -- written by the split warps with tcgen05.st.  The K = 8 TF32
  // instruction is bound by its shared-memory operand reads; A was half of them (and A_lo a shared-memory write).
  static_assert(!ATM || (RAW && C::TMEM_COLS + 64 * C::STAGES <= 512), "A-in-TMEM needs room behind the accumulators");
  constexpr int TMEM_ALLOC = ATM ? 512 : C::TMEM_COLS;
  constexpr uint32_t TMEM_A0 = C::TMEM_COLS;          // first column of the A stages
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // 128B swizzle needs 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  const uint32_t bar0 = base + C::STAGES * C::STAGE_BYTES;
  // barrier slots: full[STAGES] | empty[STAGES] | tmem_full[2] | tmem_empty[2] | tmem base address | split[STAGES]
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * C::STAGES + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 2 * C::STAGES + 4);
  auto split_bar = [&](int s) { return bar0 + 8u * (2 * C::STAGES + 5 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 16 : nullptr;
  if (dbg && threadIdx.x == 0) {
    dbg[0] = clock64();
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    dbg[7] = (long long)gt;   // wall-clock (ns) of this CTA's start
  }
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmAh); tc::tma_prefetch_desc(&tmAl);
    tc::tma_prefetch_desc(&tmBh); tc::tma_prefetch_desc(&tmBl);
    if (p.nprob > 1) {
      tc::tma_prefetch_desc(&tm2Ah); tc::tma_prefetch_desc(&tm2Al);
      tc::tma_prefetch_desc(&tm2Bh); tc::tma_prefetch_desc(&tm2Bl);
    }
    for (int s = 0; s < C::STAGES; ++s) { tc::mbar_init(full_bar(s), 1); tc::mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(tfull_bar(a), 1); tc::mbar_init(tempty_bar(a), EPI_WARPS); }
    if (RAW) for (int s = 0; s < C::STAGES; ++s) tc::mbar_init(split_bar(s), tc::SPLIT_WARPS / 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32((const void*)tmem_slot)), "r"(TMEM_ALLOC) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor
  // prefetch) may overlap the tail of the previous kernel in the stream; its results are only
  // touched below this point.  The next kernel in the chain may begin its own prologue at once.
  if (dbg && threadIdx.x == 0) dbg[1] = clock64();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (dbg && threadIdx.x == 0) dbg[2] = clock64();

  // tile index -> (graph b, problem, tile row, tile column); the two problems of a dual launch
  // alternate inside a graph so that neighbouring CTAs keep sharing operand rows in L2
  const int tiles_per_prob = p.tiles_m * p.tiles_n;
  const int tiles_per_batch = tiles_per_prob * p.nprob;
  const int total_tiles = tiles_per_batch * p.batch;
  const int num_kb = (p.K + tc::BK - 1) / tc::BK;

  if (warp == 0) {
    if (tc::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_batch, r0 = tile - b * tiles_per_batch;
        const int prob = r0 / tiles_per_prob, r = r0 - prob * tiles_per_prob;
        const int tm = r / p.tiles_n, tn = r - tm * p.tiles_n;
        const CUtensorMap* mAh = prob ? &tm2Ah : &tmAh;
        const CUtensorMap* mAl = prob ? &tm2Al : &tmAl;
        const CUtensorMap* mBh = prob ? &tm2Bh : &tmBh;
        const CUtensorMap* mBl = prob ? &tm2Bl : &tmBl;
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = base + stage * C::STAGE_BYTES;
          if (p.exp & 1) { tc::mbar_arrive(full_bar(stage)); if (++stage == C::STAGES) { stage = 0; phase ^= 1u; } continue; }
          tc::mbar_arrive_expect_tx(full_bar(stage), RAW ? C::RAW_TX_BYTES : C::STAGE_BYTES);
          tc::tma_load_3d(sa, mAh, full_bar(stage), kb * tc::BK, tm * tc::BM, b);
          if (!RAW) tc::tma_load_3d(sa + tc::A_BYTES, mAl, full_bar(stage), kb * tc::BK, tm * tc::BM, b);
          tc::tma_load_3d(sa + 2 * tc::A_BYTES, mBh, full_bar(stage), kb * tc::BK, tn * BN, b);
          if (!RAW) tc::tma_load_3d(sa + 2 * tc::A_BYTES + C::B_BYTES, mBl, full_bar(stage), kb * tc::BK, tn * BN, b);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (tc::elect_one()) {
      const uint64_t d0 = tc::umma_desc(base);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        tc::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * C::ACC_STRIDE;
        long long tw = 0, ti = 0, tcm = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          const long long c0 = dbg ? clock64() : 0;
          tc::mbar_wait(RAW ? split_bar(stage) : full_bar(stage), phase);
          tc::tc_fence_after();
          const long long c1 = dbg ? clock64() : 0;
          if (dbg && kb == 0 && tile == (int)blockIdx.x) dbg[3] = c1;
          const int krem = p.K - kb * tc::BK;
          const int ngran = krem >= tc::BK ? 4 : (krem + 7) >> 3;   // K granules of 8 that hold data
          if (ATM) {
            const uint32_t ta = tmem_base + TMEM_A0 + 64u * stage;
            switch (stage) {
              case 0: tc_issue_slab_ts<BN, 0>(d0, tmem_d, ta, ngran, kb == 0); break;
              case 1: tc_issue_slab_ts<BN, C::STAGE_BYTES>(d0, tmem_d, ta, ngran, kb == 0); break;
              case 2: tc_issue_slab_ts<BN, 2 * C::STAGE_BYTES>(d0, tmem_d, ta, ngran, kb == 0); break;
              default: tc_issue_slab_ts<BN, (C::STAGES - 1) * C::STAGE_BYTES>(d0, tmem_d, ta, ngran, kb == 0); break;
            }
          } else
          switch (stage) {
            case 0: tc_issue_slab<BN, 0>(d0, tmem_d, ngran, kb == 0); break;
            case 1: tc_issue_slab<BN, C::STAGE_BYTES>(d0, tmem_d, ngran, kb == 0); break;
            case 2: tc_issue_slab<BN, 2 * C::STAGE_BYTES>(d0, tmem_d, ngran, kb == 0); break;
            default: tc_issue_slab<BN, (C::STAGES - 1) * C::STAGE_BYTES>(d0, tmem_d, ngran, kb == 0); break;
          }
          const long long c2 = dbg ? clock64() : 0;
          tc::umma_commit(empty_bar(stage));
          if (dbg) { const long long c3 = clock64(); tw += c1 - c0; ti += c2 - c1; tcm += c3 - c2; }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        if (dbg && tile == (int)blockIdx.x) { dbg[8] = tw; dbg[9] = ti; dbg[10] = tcm; }
        tc::umma_commit(tfull_bar(acc));
        if (dbg && tile == (int)blockIdx.x) dbg[4] = clock64();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (RAW && warp >= 2 + EPI_WARPS) {
    // split warps: lo = x - trunc_tf32(x) next to the raw slab (which the tensor core reads as hi).
    // The 128B swizzle permutes 16-byte chunks identically in the hi and lo slots, so the pass is a
    // flat elementwise walk over the raw bytes.
    // Two groups of four warps take alternate slabs, so that the wait -> load -> store -> fence ->
    // arrive latency chain of one slab overlaps the next one's.
    constexpr int NT = 32 * tc::SPLIT_WARPS / 2;
    const int t = (threadIdx.x - ROLE_THREADS) % NT, grp = (threadIdx.x - ROLE_THREADS) / NT;
    constexpr int A_CH = tc::A_BYTES / 16, B_CH = C::B_BYTES / 16;   // 16-byte chunks per slab
    const long long my_tiles = total_tiles > (int)blockIdx.x ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const long long nslab = my_tiles * num_kb;
    long long tw = 0, tcv = 0;
    // Every group waits for EVERY slab in order, also the ones the other group splits: a parity
    // wait only distinguishes "this phase" from "the previous one", so a group that skipped a
    // completion of a stage barrier (three stages, two groups: the groups alternate on each stage)
    // would take the stale completion for the one it is waiting for as soon as the TMA loads of
    // two stages land out of order -- premature splits of a slab that has not arrived, then a hang.
    for (long long i = 0; i < nslab; ++i) {
      const int stage = (int)(i % C::STAGES);
      const uint32_t phase = (uint32_t)((i / C::STAGES) & 1);
      const long long c0 = dbg ? clock64() : 0;
      tc::mbar_wait(full_bar(stage), phase);
      if ((int)(i & 1) != grp) continue;
      const long long c1 = dbg ? clock64() : 0;
      uint8_t* sa = smem + stage * C::STAGE_BYTES;
      const float4* a_hi = reinterpret_cast<const float4*>(sa);
      float4* a_lo = reinterpret_cast<float4*>(sa + tc::A_BYTES);
      const float4* b_hi = reinterpret_cast<const float4*>(sa + 2 * tc::A_BYTES);
      float4* b_lo = reinterpret_cast<float4*>(sa + 2 * tc::A_BYTES + C::B_BYTES);
      // batches of eight 16-byte chunks per thread: enough loads in flight, 32 registers
      constexpr int BATCH = 8;
      if (ATM) {
        // A goes to tensor memory: every thread owns one tile row (TMEM lane 32 (warp % 4) + lane), reads its
        // 128 swizzled bytes of the raw slab and stores 32 columns of hi (the raw words: the tensor core
        // drops the low mantissa bits itself) and 32 of lo
        const int r = 32 * (warp & 3) + lane;
        const uint8_t* arow = sa + r * 128;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 x = *reinterpret_cast<const float4*>(arow + ((j ^ (r & 7)) << 4));
          hi[4 * j + 0] = __float_as_uint(x.x); lo[4 * j + 0] = __float_as_uint(tc::lo1(x.x));
          hi[4 * j + 1] = __float_as_uint(x.y); lo[4 * j + 1] = __float_as_uint(tc::lo1(x.y));
          hi[4 * j + 2] = __float_as_uint(x.z); lo[4 * j + 2] = __float_as_uint(tc::lo1(x.z));
          hi[4 * j + 3] = __float_as_uint(x.w); lo[4 * j + 3] = __float_as_uint(tc::lo1(x.w));
        }
        const uint32_t ta = tmem_base + TMEM_A0 + 64u * stage + ((uint32_t)(32 * (warp & 3)) << 16);
        tc::tmem_st32(ta, hi);
        tc::tmem_st32(ta + 32, lo);
        tc::tmem_st_wait();
        tc::tc_fence_before();
      } else {
#pragma unroll
      for (int j0 = 0; j0 < A_CH / NT; j0 += BATCH) {
        float4 x[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) x[j] = a_hi[t + (j0 + j) * NT];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) a_lo[t + (j0 + j) * NT] = tc::lo4(x[j]);
      }
      }
      constexpr int B_IT = (B_CH + NT - 1) / NT;
#pragma unroll
      for (int j0 = 0; j0 < B_IT; j0 += BATCH) {
        float4 x[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j)
          if (j0 + j < B_IT && (B_CH % NT == 0 || t + (j0 + j) * NT < B_CH)) x[j] = b_hi[t + (j0 + j) * NT];
#pragma unroll
        for (int j = 0; j < BATCH; ++j)
          if (j0 + j < B_IT && (B_CH % NT == 0 || t + (j0 + j) * NT < B_CH)) b_lo[t + (j0 + j) * NT] = tc::lo4(x[j]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tensor-core (async proxy) reads
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(split_bar(stage));
      if (dbg) { const long long c2 = clock64(); tw += c1 - c0; tcv += c2 - c1; }
    }
    if (dbg && t == 0 && grp == 0) { dbg[11] = tw; dbg[12] = 0; dbg[13] = tcv; }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    int acc = 0; uint32_t acc_phase = 0;
    int cbuf = 0;            // staging tile of the next TMA store (raw kernel)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_batch, r0 = tile - b * tiles_per_batch;
      const int prob = r0 / tiles_per_prob, r = r0 - prob * tiles_per_prob;
      const int tm = r / p.tiles_n, tn = r - tm * p.tiles_n;
      const TcEpi& ep = p.e[prob];
      tc::mbar_wait(tfull_bar(acc), acc_phase);
      tc::tc_fence_after();
      if (dbg && tile == (int)blockIdx.x && warp == 2 && lane == 0) dbg[5] = clock64();
      if (RAW && p.tma_c[prob]) {
        tc_epilogue_tile_tma<BN>(p, ep, prob ? &tm2C : &tmC, prob ? &tm2Cp : &tmCp, tmem_base + acc * C::ACC_STRIDE, tm, tn, b, q, lane,
                                 smem + C::STAGES * C::STAGE_BYTES + 1024, base + C::STAGES * C::STAGE_BYTES + 1024,
                                 cbuf, warp == 2 && lane == 0);
      } else {
        tc_epilogue_tile<BN, EPI_WARPS>(p, ep, tmem_base + acc * C::ACC_STRIDE, tm, tn, b, q, (warp - 2) >> 2, lane,
                             reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + 256) + (warp - 2) * 32 * EPI_LD,
                             tile == (int)blockIdx.x ? dbg : nullptr);
      }
      tc::tc_fence_before();
      __syncwarp();
      if (dbg && tile == (int)blockIdx.x && warp == 2 && lane == 0) dbg[6] = clock64();
      if (lane == 0) tc::mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (RAW && warp == 2 && lane == 0) tc::tma_store_wait_all();   // staging tiles and C are settled before the CTA leaves
  }

  tc::tc_fence_before();
  __syncthreads();
  if (dbg && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    dbg[0] = (long long)gt;   // wall-clock (ns) of this CTA's end (overwrites the cycle stamp)
  }
  if (warp == 2) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_ALLOC) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// Persistent CHAIN of dependent products (the Newton-Schulz iterations of the large-D path).
//
// One launch runs a whole list of stages; a stage is one product or two independent products of
// the same D x D x D shape (plus, optionally, the in-place antisymmetrisation W <- W - W^T of one
// buffer, done by the epilogue warps before their tiles).  Consecutive stages depend on each other
// through whole matrices, so they are separated by a grid-wide barrier: every CTA (one per SM, all
// co-resident) publishes the end of its tiles of stage s on a global counter, and the TMA producer
// of every CTA waits for that counter before its first load of stage s + 1.  Everything else of
// the pipeline -- the shared-memory slab ring, the split warps, the two TMEM accumulator stages --
// runs straight through the stage boundaries.
//
// Why: as separate launches every product of the chain is ONE tile per CTA, and the launch, the
// dependency wait, the first TMA round trip, the epilogue and the teardown of each of the ~50
// launches per layer are all exposed (at D = 1000 the MMA main loop is ~9 us of a ~26-32 us launch;
// at 32 x D = 200 ~2 us of ~20 us).  The chain pays launch and teardown once and a ~2 us barrier
// per stage.
//
// Operands are plain FP32 [batch][D][ld] buffers (raw mode: lo = x - trunc(x) formed in shared
// memory).  A buffer has ONE tensor map per box height (128 rows for the A operand and for the
// C stores -- the store box is the same 32 x 128 box -- and BN rows for the B operand), so a stage
// only names buffer indices and the whole description travels as a kernel parameter.
namespace tc {
constexpr int CH_MAX_BUF = 12, CH_MAX_STAGE = 40;
}
struct ChainProb {
  int a, b, c, e1;            // buffer indices: C = alpha A B^T + beta E1 + diag I   (e1 < 0: no addend)
  float alpha, beta, diag;
  const float* alpha_dev;     // optional per-graph factor of alpha
};
struct ChainStage {
  int nprob;                  // 1 or 2 independent products
  int antisym;                // >= 0: buffer to antisymmetrise in place (W <- W - W^T) before this stage's tiles
  ChainProb p[2];
};
struct ChainBuf {
  float* ptr;
  long long stride;           // floats between graphs
  int ld;
  int tma;                    // 1: rows are 16-byte multiples -> has tensor maps (operand / TMA-stored output)
};
struct ChainParams {
  CUtensorMap mapA[tc::CH_MAX_BUF];   // box 32 x 128 (A operand, C store)
  CUtensorMap mapB[tc::CH_MAX_BUF];   // box 32 x BN  (B operand)
  ChainBuf buf[tc::CH_MAX_BUF];
  ChainStage st[tc::CH_MAX_STAGE];
  int nstages, M, N, K, batch, tiles_m, tiles_n;
  unsigned* barrier;                  // zeroed before the launch
  long long* dbg;                     // developer timeline [stage][CTA][8] of %globaltimer stamps (null = off)
};

namespace tc {
__device__ __forceinline__ long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return (long long)t;
}
// wait until `target` CTAs-times-stages have arrived; faults (never hangs) if an arrival is lost
__device__ __forceinline__ void grid_wait(const unsigned* bar, unsigned target) {
  unsigned v = 0;
  long long t0 = 0;
  for (int spin = 0;; ++spin) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    if (v >= target) break;
    if ((spin & 255) == 255) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) asm volatile("trap;");
    }
  }
  __threadfence();                                             // L1 of this SM holds nothing older than the arrivals
  asm volatile("fence.proxy.async.global;" ::: "memory");      // ... and neither do the TMA (async proxy) reads that follow
}
__device__ __forceinline__ void grid_arrive(unsigned* bar) {
  asm volatile("fence.proxy.async.global;" ::: "memory");      // completed TMA stores -> generic proxy
  __threadfence();
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
}
}  // namespace tc

template <int BN>
__global__ void __launch_bounds__(tc::THREADS_RAW, 1) tc_chain_kernel(const __grid_constant__ ChainParams P) {
  using C = tc::Cfg<BN>;
  static_assert(BN % 32 == 0, "the chain stores C through the 32-column TMA box only");
  constexpr int EPI_WARPS = tc::EPI_WARPS_RAW;
  constexpr int ROLE_THREADS = 64 + 32 * EPI_WARPS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  const uint32_t bar0 = base + C::STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * C::STAGES + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 2 * C::STAGES + 4);
  auto split_bar = [&](int s) { return bar0 + 8u * (2 * C::STAGES + 5 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) { tc::mbar_init(full_bar(s), 1); tc::mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(tfull_bar(a), 1); tc::mbar_init(tempty_bar(a), EPI_WARPS); }
    for (int s = 0; s < C::STAGES; ++s) tc::mbar_init(split_bar(s), tc::SPLIT_WARPS / 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32((const void*)tmem_slot)), "r"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int tiles_per_prob = P.tiles_m * P.tiles_n;
  const int num_kb = (P.K + tc::BK - 1) / tc::BK;
  const int G = (int)gridDim.x;

  if (warp == 0) {
    if (tc::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int s = 0; s < P.nstages; ++s) {
        long long* dg = P.dbg ? P.dbg + ((size_t)s * G + blockIdx.x) * 8 : nullptr;
        if (dg) dg[0] = tc::gtime();
        if (s > 0) tc::grid_wait(P.barrier, (unsigned)s * (unsigned)G);   // every CTA is done with stage s - 1
        if (dg) dg[1] = tc::gtime();
        const ChainStage& S = P.st[s];
        const int tiles_per_batch = tiles_per_prob * S.nprob;
        const int total_tiles = tiles_per_batch * P.batch;
        for (int tile = blockIdx.x; tile < total_tiles; tile += G) {
          const int b = tile / tiles_per_batch, r0 = tile - b * tiles_per_batch;
          const int prob = r0 / tiles_per_prob, r = r0 - prob * tiles_per_prob;
          const int tm = r / P.tiles_n, tn = r - tm * P.tiles_n;
          const CUtensorMap* mA = &P.mapA[S.p[prob].a];
          const CUtensorMap* mB = &P.mapB[S.p[prob].b];
          for (int kb = 0; kb < num_kb; ++kb) {
            tc::mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = base + stage * C::STAGE_BYTES;
            tc::mbar_arrive_expect_tx(full_bar(stage), C::RAW_TX_BYTES);
            tc::tma_load_3d(sa, mA, full_bar(stage), kb * tc::BK, tm * tc::BM, b);
            tc::tma_load_3d(sa + 2 * tc::A_BYTES, mB, full_bar(stage), kb * tc::BK, tn * BN, b);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (tc::elect_one()) {
      const uint64_t d0 = tc::umma_desc(base);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int s = 0; s < P.nstages; ++s) {
        const int total_tiles = tiles_per_prob * P.st[s].nprob * P.batch;
        long long* dg = P.dbg ? P.dbg + ((size_t)s * G + blockIdx.x) * 8 : nullptr;
        for (int tile = blockIdx.x; tile < total_tiles; tile += G) {
          tc::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          tc::tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * C::ACC_STRIDE;
          for (int kb = 0; kb < num_kb; ++kb) {
            tc::mbar_wait(split_bar(stage), phase);
            tc::tc_fence_after();
            if (dg && kb == 0 && tile == (int)blockIdx.x) dg[2] = tc::gtime();
            const int krem = P.K - kb * tc::BK;
            const int ngran = krem >= tc::BK ? 4 : (krem + 7) >> 3;
            switch (stage) {
              case 0: tc_issue_slab<BN, 0>(d0, tmem_d, ngran, kb == 0); break;
              case 1: tc_issue_slab<BN, C::STAGE_BYTES>(d0, tmem_d, ngran, kb == 0); break;
              case 2: tc_issue_slab<BN, 2 * C::STAGE_BYTES>(d0, tmem_d, ngran, kb == 0); break;
              default: tc_issue_slab<BN, (C::STAGES - 1) * C::STAGE_BYTES>(d0, tmem_d, ngran, kb == 0); break;
            }
            tc::umma_commit(empty_bar(stage));
            if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
          }
          tc::umma_commit(tfull_bar(acc));
          if (dg) dg[3] = tc::gtime();
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 2 + EPI_WARPS) {
    // split warps: the slabs of all stages form one sequence through the ring
    constexpr int NT = 32 * tc::SPLIT_WARPS / 2;
    const int t = (threadIdx.x - ROLE_THREADS) % NT, grp = (threadIdx.x - ROLE_THREADS) / NT;
    constexpr int A_CH = tc::A_BYTES / 16, B_CH = C::B_BYTES / 16;
    long long nslab = 0;
    for (int s = 0; s < P.nstages; ++s) {
      const int total_tiles = tiles_per_prob * P.st[s].nprob * P.batch;
      const long long my_tiles = total_tiles > (int)blockIdx.x ? (total_tiles - 1 - (int)blockIdx.x) / G + 1 : 0;
      nslab += my_tiles * num_kb;
    }
    for (long long i = 0; i < nslab; ++i) {
      const int stage = (int)(i % C::STAGES);
      const uint32_t phase = (uint32_t)((i / C::STAGES) & 1);
      tc::mbar_wait(full_bar(stage), phase);     // every group waits for every slab in order (see tc_gemm_kernel)
      if ((int)(i & 1) != grp) continue;
      uint8_t* sa = smem + stage * C::STAGE_BYTES;
      const float4* a_hi = reinterpret_cast<const float4*>(sa);
      float4* a_lo = reinterpret_cast<float4*>(sa + tc::A_BYTES);
      const float4* b_hi = reinterpret_cast<const float4*>(sa + 2 * tc::A_BYTES);
      float4* b_lo = reinterpret_cast<float4*>(sa + 2 * tc::A_BYTES + C::B_BYTES);
      constexpr int BATCH = 8;
#pragma unroll
      for (int j0 = 0; j0 < A_CH / NT; j0 += BATCH) {
        float4 x[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) x[j] = a_hi[t + (j0 + j) * NT];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) a_lo[t + (j0 + j) * NT] = tc::lo4(x[j]);
      }
      constexpr int B_IT = (B_CH + NT - 1) / NT;
#pragma unroll
      for (int j0 = 0; j0 < B_IT; j0 += BATCH) {
        float4 x[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j)
          if (j0 + j < B_IT && (B_CH % NT == 0 || t + (j0 + j) * NT < B_CH)) x[j] = b_hi[t + (j0 + j) * NT];
#pragma unroll
        for (int j = 0; j < BATCH; ++j)
          if (j0 + j < B_IT && (B_CH % NT == 0 || t + (j0 + j) * NT < B_CH)) b_lo[t + (j0 + j) * NT] = tc::lo4(x[j]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(split_bar(stage));
    }
  } else {
    // epilogue warps (128 threads): per stage [optional antisymmetrisation], the tiles, then the arrival
    const int q = warp & 3;
    const bool leader = warp == 2 && lane == 0;
    int acc = 0; uint32_t acc_phase = 0;
    int cbuf = 0;
    uint8_t* cst = smem + C::STAGES * C::STAGE_BYTES + 1024;
    const uint32_t cst_s = base + C::STAGES * C::STAGE_BYTES + 1024;
    TcParams tp;
    tp.M = P.M; tp.N = P.N;
    for (int s = 0; s < P.nstages; ++s) {
      const ChainStage& S = P.st[s];
      if (S.antisym >= 0) {
        // W <- W - W^T in place, mirrored 32 x 32 tile pairs spread over the CTAs.  Its input was written by
        // stage s - 1 (other CTAs), its result is read by stage s + 1: both sides are covered by the barriers.
        if (leader) {
          tc::tma_store_wait_all();                       // the staging area doubles as the transpose buffer
          if (s > 0) tc::grid_wait(P.barrier, (unsigned)s * (unsigned)G);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const ChainBuf W = P.buf[S.antisym];
        float (*ta)[33] = reinterpret_cast<float (*)[33]>(cst);
        float (*tb)[33] = reinterpret_cast<float (*)[33]>(cst + 32 * 33 * 4);
        const int nt = (P.M + 31) / 32, ty = warp - 2, tx = lane;
        const int npair = nt * nt * P.batch;
        for (int pr = blockIdx.x; pr < npair; pr += G) {
          const int bb = pr / (nt * nt), rr = pr - bb * nt * nt;
          const int bxi = rr / nt, byi = rr - bxi * nt;
          if (bxi > byi) continue;                        // uniform over the CTA
          const int bx = bxi * 32, by = byi * 32;
          float* Wb = W.ptr + (size_t)bb * W.stride;
          for (int r = ty; r < 32; r += 4) {
            ta[r][tx] = (bx + r < P.M && by + tx < P.M) ? Wb[(size_t)(bx + r) * W.ld + by + tx] : 0.f;
            tb[r][tx] = (by + r < P.M && bx + tx < P.M) ? Wb[(size_t)(by + r) * W.ld + bx + tx] : 0.f;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          for (int r = ty; r < 32; r += 4) {
            if (bx + r < P.M && by + tx < P.M) Wb[(size_t)(bx + r) * W.ld + by + tx] = ta[r][tx] - tb[tx][r];
            if (bxi != byi && by + r < P.M && bx + tx < P.M) Wb[(size_t)(by + r) * W.ld + bx + tx] = tb[r][tx] - ta[tx][r];
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the area goes back to the TMA stores
      }
      const int tiles_per_batch = tiles_per_prob * S.nprob;
      const int total_tiles = tiles_per_batch * P.batch;
      for (int tile = blockIdx.x; tile < total_tiles; tile += G) {
        const int b = tile / tiles_per_batch, r0 = tile - b * tiles_per_batch;
        const int prob = r0 / tiles_per_prob, r = r0 - prob * tiles_per_prob;
        const int tm = r / P.tiles_n, tn = r - tm * P.tiles_n;
        const ChainProb& cp = S.p[prob];
        const ChainBuf cb = P.buf[cp.c];
        TcEpi ep;
        ep.alpha = cp.alpha; ep.beta = cp.beta; ep.diag = cp.diag; ep.alpha_dev = cp.alpha_dev;
        ep.E1_lo = nullptr; ep.C_lo = nullptr;
        if (cp.e1 >= 0) { const ChainBuf eb = P.buf[cp.e1]; ep.E1_hi = eb.ptr; ep.sE1 = eb.stride; ep.lde1 = eb.ld; }
        else { ep.E1_hi = nullptr; ep.sE1 = 0; ep.lde1 = 0; }
        ep.C_hi = cb.ptr; ep.sC = cb.stride; ep.ldc = cb.ld;
        tc::mbar_wait(tfull_bar(acc), acc_phase);
        tc::tc_fence_after();
        if (P.dbg && leader) P.dbg[((size_t)s * G + blockIdx.x) * 8 + 4] = tc::gtime();   // accumulator of the (last) tile visible
        if (cb.tma) {
          tc_epilogue_tile_tma<BN>(tp, ep, &P.mapA[cp.c], &P.mapA[cp.c], tmem_base + acc * C::ACC_STRIDE, tm, tn, b, q, lane,
                                   cst, cst_s, cbuf, leader);
        } else {
          tc_epilogue_tile<BN, EPI_WARPS>(tp, ep, tmem_base + acc * C::ACC_STRIDE, tm, tn, b, q, 0, lane, nullptr, nullptr);
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tempty_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      // end of this CTA's share of stage s: direct stores of all 128 threads and the TMA stores of the leader
      // are complete and visible before the arrival.  The counter is cumulative (target (s + 1) G), so an
      // arrival for stage s must never be counted before stage s - 1 is complete EVERYWHERE: a CTA without a
      // tile in stage s would otherwise run ahead and let the count reach the target while a slow CTA is
      // still inside stage s - 1.  (A CTA with tiles has passed that barrier already: its loads waited for it.)
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (leader) {
        long long* dg = P.dbg ? P.dbg + ((size_t)s * G + blockIdx.x) * 8 : nullptr;
        if (dg) dg[5] = tc::gtime();                      // epilogues issued
        tc::tma_store_wait_all();
        if (dg) dg[6] = tc::gtime();                      // stores complete
        if (s > 0) tc::grid_wait(P.barrier, (unsigned)s * (unsigned)G);
        tc::grid_arrive(P.barrier);
        if (dg) dg[7] = tc::gtime();                      // arrived
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// host side: tensor maps (cached: the operands live in a handful of fixed scratch matrices)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

struct MapKey {
  const void* ptr; int rows, cols, ld, batch, box_rows, box_cols; long long stride;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && batch == o.batch &&
           box_rows == o.box_rows && box_cols == o.box_cols && stride == o.stride;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&](size_t v) { h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); };
    mix(k.rows); mix(k.cols); mix(k.ld); mix(k.batch); mix(k.box_rows); mix(k.box_cols); mix((size_t)k.stride);
    return h;
  }
};
static std::mutex g_map_mu;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// tensor map of a batched row-major matrix [batch][rows][cols] (row stride ld, batch stride `stride`
// floats), box = 32 columns x box_rows rows x 1, 128-byte swizzle, zero fill outside the matrix
// (box_cols < 32: a dense, unswizzled box -- the trailing partial chunk of a C tile)
static int get_map(const float* ptr, int rows, int cols, int ld, int batch, long long stride, int box_rows,
                   CUtensorMap* out, int box_cols = tc::BK) {
  const MapKey key{ptr, rows, cols, ld, batch, box_rows, box_cols, stride};
  std::lock_guard<std::mutex> lk(g_map_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return 0; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return 1; }
  const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(batch > 1 ? stride : (long long)ld * rows) * 4};
  const cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == tc::BK ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for rows=%d cols=%d ld=%d batch=%d", (int)r, rows, cols, ld, batch);
    return 1;
  }
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps.emplace(key, m);
  *out = m;
  return 0;
}
void tc_forget_maps() {
  std::lock_guard<std::mutex> lk(g_map_mu);
  g_maps.clear();
}

// per-device caches (one process may drive several devices)
constexpr int MAX_DEV = 64;
static int g_num_sms_dev[MAX_DEV] = {0};
static int num_sms(int* out) {
  int dev = 0;
  UGLAD_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= MAX_DEV) { set_error("device ordinal %d out of range", dev); return 1; }
  if (g_num_sms_dev[dev] == 0)
    UGLAD_CUDA(cudaDeviceGetAttribute(&g_num_sms_dev[dev], cudaDevAttrMultiProcessorCount, dev));
  *out = g_num_sms_dev[dev];
  return 0;
}
static int g_tc_bn = 0;   // 0 = auto
static int g_tc_pdl = 1;  // programmatic dependent launch between the chained products
int tc_tune_bn(int bn) { g_tc_bn = bn; return 0; }
int tc_tune_pdl(int on) { g_tc_pdl = on; return 0; }
static int g_tc_raw = 1;   // plain FP32 operands, hi/lo split inside the kernel (0: pre-split pairs in HBM)
int tc_tune_raw(int on) { g_tc_raw = on ? 1 : 0; return 0; }
bool tc_raw_enabled() { return g_tc_raw != 0; }
static int g_tc_tma_store = 1;   // raw kernel: C through TMA stores (0: direct stores from the epilogue warps)
int tc_tune_tma_store(int on) { g_tc_tma_store = on ? 1 : 0; return 0; }
static int g_tc_exp = 0;
int tc_tune_exp(int v) { g_tc_exp = v; return 0; }
static int g_tc_atm = 1;   // BN = 64 raw kernel: A operand of the MMAs in tensor memory (0: in shared memory)
int tc_tune_atm(int on) { g_tc_atm = on ? 1 : 0; return 0; }
static int g_tc_dual = 1;  // pair independent same-shape products into one launch
int tc_tune_dual(int on) { g_tc_dual = on; return 0; }
static long long* g_tc_dbg = nullptr;   // developer timeline buffer (uglad_tc_debug_buffer)
void tc_set_debug(long long* buf) { g_tc_dbg = buf; }

static void fill_epi(TcEpi& e, const TcGemm& g) {
  e.alpha = g.alpha; e.beta = g.beta; e.diag = g.diag; e.alpha_dev = g.alpha_dev;
  e.E1_hi = g.E1_hi; e.E1_lo = g.E1_lo; e.sE1 = g.sE1; e.lde1 = g.lde1;
  e.C_hi = g.C_hi; e.C_lo = g.C_lo; e.sC = g.sC; e.ldc = g.ldc;
}

// g2 == nullptr: one product; otherwise two independent products of identical shape in one launch
template <int BN, bool RAW, bool ATM = false>
static int launch_tc(const TcGemm& g, const TcGemm* g2, int batch, cudaStream_t st) {
  using C = tc::Cfg<BN>;
  static bool attr_set[MAX_DEV] = {false};   // the function attribute is per device
  int dev = 0;
  UGLAD_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= MAX_DEV) { set_error("device ordinal %d out of range", dev); return 1; }
  if (!attr_set[dev]) {
    UGLAD_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, RAW, ATM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    RAW ? C::SMEM_RAW : C::SMEM));
    attr_set[dev] = true;
  }
  CUtensorMap m[8];
  const TcGemm* gs[2] = {&g, g2 ? g2 : &g};
  for (int i = 0; i < 2; ++i) {
    const TcGemm& q = *gs[i];
    if (get_map(q.A_hi, q.M, q.K, q.lda, batch, q.sA, tc::BM, &m[4 * i + 0])) return 1;
    if (get_map(q.B_hi, q.N, q.K, q.ldb, batch, q.sB, BN, &m[4 * i + 2])) return 1;
    if (RAW) {   // plain operands: the lo maps are never dereferenced
      m[4 * i + 1] = m[4 * i + 0];
      m[4 * i + 3] = m[4 * i + 2];
    } else {
      if (get_map(q.A_lo, q.M, q.K, q.lda, batch, q.sA, tc::BM, &m[4 * i + 1])) return 1;
      if (get_map(q.B_lo, q.N, q.K, q.ldb, batch, q.sB, BN, &m[4 * i + 3])) return 1;
    }
  }
  TcParams p;
  CUtensorMap mc[2], mcp[2];
  for (int i = 0; i < 2; ++i) {   // raw kernel: C leaves through TMA stores when its rows are 16-byte aligned
    const TcGemm& q = *gs[i];
    auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
    const bool ok = RAW && g_tc_tma_store && q.C_lo == nullptr && q.ldc % 4 == 0 && q.sC % 4 == 0 && al16(q.C_hi) &&
                    (q.E1_hi == nullptr || (q.E1_lo == nullptr && q.lde1 % 4 == 0 && q.sE1 % 4 == 0 && al16(q.E1_hi)));
    p.tma_c[i] = ok ? 1 : 0;
    if (ok) {
      if (get_map(q.C_hi, q.M, q.N, q.ldc, batch, q.sC, tc::BM, &mc[i])) return 1;
      mcp[i] = mc[i];
      if (BN % 32 && get_map(q.C_hi, q.M, q.N, q.ldc, batch, q.sC, tc::BM, &mcp[i], BN % 32)) return 1;
    } else {
      mc[i] = mcp[i] = m[0];   // never dereferenced
    }
  }
  p.dbg = g_tc_dbg;
  p.M = g.M; p.N = g.N; p.K = g.K; p.batch = batch;
  p.tiles_m = (g.M + tc::BM - 1) / tc::BM;
  p.tiles_n = (g.N + BN - 1) / BN;
  p.nprob = g2 ? 2 : 1;
  p.exp = g_tc_exp;
  fill_epi(p.e[0], g);
  fill_epi(p.e[1], *gs[1]);
  const long long total = (long long)p.tiles_m * p.tiles_n * batch * p.nprob;
  int g_num_sms = 0;
  if (num_sms(&g_num_sms)) return 1;
  const int grid = (int)(total < g_num_sms ? total : g_num_sms);
  profile_begin(st, 1, 2.0 * g.M * g.N * g.K * batch * p.nprob);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(RAW ? tc::THREADS_RAW : tc::THREADS);
  cfg.dynamicSmemBytes = RAW ? C::SMEM_RAW : C::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_tc_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  UGLAD_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<BN, RAW, ATM>, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], mc[0], mc[1],
                                mcp[0], mcp[1], p));
  profile_end(st);
  UGLAD_CHECK_LAUNCH("tc_gemm_kernel");
  return 0;
}

bool tc_gemm_supported(const TcGemm& g) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  // both operands pre-split (hi + lo) or both plain (lo == nullptr: split inside the kernel)
  return g.lda % 4 == 0 && g.ldb % 4 == 0 && g.sA % 4 == 0 && g.sB % 4 == 0 && al16(g.A_hi) && al16(g.A_lo) &&
         al16(g.B_hi) && al16(g.B_lo) && g.A_hi && g.B_hi && ((g.A_lo == nullptr) == (g.B_lo == nullptr)) &&
         g.M > 0 && g.N > 0 && g.K > 0;
}

static int launch_tc_any(const TcGemm& g, const TcGemm* g2, int batch, cudaStream_t st) {
  if (!tc_gemm_supported(g) || (g2 && !tc_gemm_supported(*g2))) {
    set_error("tc_gemm: operands must be 16-byte aligned with ld %% 4 == 0, both split or both plain");
    return 1;
  }
  if (g2 && (g2->M != g.M || g2->N != g.N || g2->K != g.K || (g2->A_lo == nullptr) != (g.A_lo == nullptr))) {
    set_error("tc_gemm: dual launch needs identical shapes and operand forms");
    return 1;
  }
  const bool raw = g.A_lo == nullptr;
  if (batch <= 0) return 0;
  int g_num_sms = 0;
  if (num_sms(&g_num_sms)) return 1;
  const int nprob = g2 ? 2 : 1;
  int bn = g_tc_bn;
  if (bn == 0) {
    // narrowest tile that covers N in one pass; for wide N pick the width that fills the SMs best
    if (g.N <= 64) bn = 64;
    else if (g.N <= 112) bn = 112;
    else if (g.N <= 128) bn = 128;
    else {
      const long long tm = (g.M + tc::BM - 1) / tc::BM;
      double best = 1e300;
      const int cand[3] = {128, 112, 64};
      for (int c : cand) {
        const long long tiles = tm * ((g.N + c - 1) / c) * batch * nprob;
        const long long waves = (tiles + g_num_sms - 1) / g_num_sms;
        const double cost = (double)waves * (c + 40);   // per-tile time ~ operand rows, fixed overhead
        if (cost < best) { best = cost; bn = c; }
      }
    }
  }
  switch (bn) {
    case 64: return raw ? (g_tc_atm ? launch_tc<64, true, true>(g, g2, batch, st) : launch_tc<64, true>(g, g2, batch, st))
                        : launch_tc<64, false>(g, g2, batch, st);
    case 112: return raw ? launch_tc<112, true>(g, g2, batch, st) : launch_tc<112, false>(g, g2, batch, st);
    case 128: return raw ? launch_tc<128, true>(g, g2, batch, st) : launch_tc<128, false>(g, g2, batch, st);
  }
  set_error("tc_gemm: unsupported tile width %d", bn);
  return 1;
}
int launch_tc_gemm(const TcGemm& g, int batch, cudaStream_t st) { return launch_tc_any(g, nullptr, batch, st); }
int launch_tc_gemm2(const TcGemm& g1, const TcGemm& g2, int batch, cudaStream_t st) {
  return g_tc_dual ? launch_tc_any(g1, &g2, batch, st)
                   : (launch_tc_any(g1, nullptr, batch, st) || launch_tc_any(g2, nullptr, batch, st));
}


// ---- host side of the persistent chain ----------------------------------------------------------
static int g_tc_chain = 0;      // 1: the Newton-Schulz iterations run as one persistent chain launch (measured slower than the
                                // PDL-chained launches while the main loop is shared-memory-bound: DESIGN.md)
static int g_tc_chain_bn = 0;   // 0 auto / 64 / 128
int tc_tune_chain(int on) { g_tc_chain = on ? 1 : 0; return 0; }
int tc_tune_chain_bn(int bn) { g_tc_chain_bn = bn; return 0; }
bool tc_chain_enabled() { return g_tc_chain != 0 && g_tc_raw != 0; }

template <int BN>
static int launch_chain_bn(const TcChain& c, cudaStream_t st) {
  using C = tc::Cfg<BN>;
  static bool attr_set[MAX_DEV] = {false};
  int dev = 0;
  UGLAD_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= MAX_DEV) { set_error("device ordinal %d out of range", dev); return 1; }
  if (!attr_set[dev]) {
    UGLAD_CUDA(cudaFuncSetAttribute(tc_chain_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_RAW));
    attr_set[dev] = true;
  }
  static ChainParams P;   // large (several KB): built in place under the lock below, copied by the launch
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  memset(&P, 0, sizeof(P));
  P.nstages = c.nstages; P.M = P.N = P.K = c.D; P.batch = c.batch;
  P.tiles_m = (c.D + tc::BM - 1) / tc::BM;
  P.tiles_n = (c.D + BN - 1) / BN;
  P.barrier = c.barrier;
  P.dbg = g_tc_dbg;
  for (int i = 0; i < c.nbuf; ++i) {
    const TcChainBuf& b = c.buf[i];
    const bool tma = b.ld % 4 == 0 && b.stride % 4 == 0 && (reinterpret_cast<uintptr_t>(b.ptr) & 15) == 0;
    P.buf[i].ptr = b.ptr; P.buf[i].ld = b.ld; P.buf[i].stride = b.stride; P.buf[i].tma = tma ? 1 : 0;
    if (tma) {
      if (get_map(b.ptr, c.D, c.D, b.ld, c.batch, b.stride, tc::BM, &P.mapA[i])) return 1;
      if (get_map(b.ptr, c.D, c.D, b.ld, c.batch, b.stride, BN, &P.mapB[i])) return 1;
    }
  }
  int max_tiles = 0;
  for (int s = 0; s < c.nstages; ++s) {
    const TcChainStage& S = c.st[s];
    P.st[s].nprob = S.nprob;
    P.st[s].antisym = S.antisym;
    for (int k = 0; k < S.nprob; ++k) {
      const TcChainProb& q = S.p[k];
      if (!P.buf[q.a].tma || !P.buf[q.b].tma) { set_error("tc_chain: operand buffers need 16-byte aligned rows"); return 1; }
      P.st[s].p[k] = ChainProb{q.a, q.b, q.c, q.e1, q.alpha, q.beta, q.diag, q.alpha_dev};
    }
    const int tiles = P.tiles_m * P.tiles_n * c.batch * S.nprob;
    if (tiles > max_tiles) max_tiles = tiles;
  }
  int nsm = 0;
  if (num_sms(&nsm)) return 1;
  const int grid = max_tiles < nsm ? max_tiles : nsm;
  UGLAD_CUDA(cudaMemsetAsync(c.barrier, 0, sizeof(unsigned), st));
  double flops = 0.0;
  for (int s = 0; s < c.nstages; ++s) flops += 2.0 * c.D * c.D * c.D * c.batch * c.st[s].nprob;
  profile_begin(st, 1, flops);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(tc::THREADS_RAW);
  cfg.dynamicSmemBytes = C::SMEM_RAW;
  cfg.stream = st;
  cfg.attrs = nullptr;
  cfg.numAttrs = 0;
  UGLAD_CUDA(cudaLaunchKernelEx(&cfg, tc_chain_kernel<BN>, P));
  profile_end(st);
  UGLAD_CHECK_LAUNCH("tc_chain_kernel");
  return 0;
}

int launch_tc_chain(const TcChain& c, cudaStream_t st) {
  if (c.nstages <= 0 || c.nstages > tc::CH_MAX_STAGE || c.nbuf > tc::CH_MAX_BUF || !c.barrier) {
    set_error("tc_chain: %d stages / %d buffers outside the kernel's limits", c.nstages, c.nbuf);
    return 1;
  }
  int nsm = 0;
  if (num_sms(&nsm)) return 1;
  int bn = g_tc_chain_bn;
  if (bn == 0) {
    // wide tiles when a single-product stage still fills most of the SMs, narrow ones otherwise
    const long long t128 = (long long)((c.D + 127) / 128) * ((c.D + 127) / 128) * c.batch;
    bn = (t128 * 10 >= (long long)nsm * 8) ? 128 : 64;
  }
  return bn == 128 ? launch_chain_bn<128>(c, st) : launch_chain_bn<64>(c, st);
}
}  // namespace uglad
