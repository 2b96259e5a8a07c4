/* uglad_b200 C-ABI: B200 (sm_100a) kernels for uGLAD's unrolled GLAD hot path.
 *
 * The reference (Harshs27/uGLAD) is pure Python/torch and has no FFI; each entry point
 * below replaces the Python function cited beside it (paths under /root/reference/uglad)
 * and is what a ctypes binding in that file would call.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to float32 unless the name ends in _host;
 *  - matrices are dense row-major, batches are contiguous: S[b][i][j];
 *  - eigenvector matrices are stored "vector-major": Vt[b][k][:] is eigenvector k;
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *  - every function returns 0 on success, non-zero on error; uglad_last_error() gives the
 *    message of the last failing call on the calling thread.  Nothing here falls back to
 *    the CPU.
 */
#ifndef UGLAD_B200_H
#define UGLAD_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define UGLAD_ABI_VERSION 1
#define UGLAD_NF 3      /* rho_l1 input features (theta_k1, S, theta_prev): glad.py:144 */
#define UGLAD_MAX_H 8   /* largest hidden width supported by the fused MLP kernels     */

typedef struct uglad_dims {
  int B;            /* graphs held by this process                                     */
  int D;            /* nodes per graph                                                  */
  int L;            /* unrolled layers (glad.py:78)                                     */
  int H;            /* hidden width of rho_l1 and lambda_f (glad_params.py:26)          */
  int init_diag;    /* glad.py:106-117: 0 -> (S + t I)^-1, 1 -> diag(1/(S_ii + t))      */
  int B_total;      /* graphs over all processes (mean of the Frobenius term, loss /B)  */
  int exact_sqrt;   /* 0: reproduce torch_sqrtm.py's 10 Newton-Schulz steps; 1: sqrt    */
  float lambda_init;/* glad.py:77                                                       */
} uglad_dims;

int uglad_abi_version(void);
const char* uglad_last_error(void);

/* number of floats in the packed parameter / gradient vector for hidden width H:
 * [theta_init_offset | rho_l1.0.weight(H x 3) .bias(H) | rho_l1.2.weight(H x H) .bias(H) |
 *  rho_l1.4.weight(1 x H) .bias(1) | lambda_f.0.weight(H x 2) .bias(H) |
 *  lambda_f.2.weight(1 x H) .bias(1)]  -- GladParams.state_dict() order, glad_params.py:10-31 */
size_t uglad_param_count(int H);

/* (1) covariance: replaces sklearn empirical_covariance in prepare_data.py:342-344.
 * X[B][M][D] samples -> S[B][D][D] = (X-mean)^T (X-mean) / M.  mean_out[B][D] is required
 * (it is also the kernel's scratch). */
int uglad_covariance(const float* X, int B, int M, int D, float* S, float* mean_out, void* stream);
/* the same covariance with the contraction on the tensor pipe (tcgen05 3xTF32 on the centred,
 * feature-major samples staged in scratch: uglad_covariance_scratch_floats floats); scratch == NULL
 * or uglad_tune("use_tc", 0) selects the FP32 kernel of uglad_covariance.                        */
size_t uglad_covariance_scratch_floats(int B, int M, int D);
int uglad_covariance_ws(const float* X, int B, int M, int D, float* S, float* mean_out, float* scratch, void* stream);

/* symmetric eigensolver used by every stage.  shift_mode 0: A is positive definite (high
 * relative accuracy, eigenvalues returned as column norms); 1: A is symmetric indefinite.
 * w[B][D], Vt[B][D][D], info[B][4] = {sweeps, shift, trace(A), sum(w)} (info may be NULL).
 * scratch: uglad_eigh_scratch_floats(B, D) floats (may be NULL when that returns 0).      */
size_t uglad_eigh_scratch_floats(int B, int D);
int uglad_eigh(const float* A, int B, int D, int shift_mode, float* w, float* Vt, float* info,
               float* scratch, void* stream);

/* same with a warm start: warm_Vt / warm_w = eigenvectors / eigenvalues of a NEARBY matrix batch of
 * the same shape (e.g. the previous batch of a stream of similar sample matrices).  They only
 * change the work done (fewer Jacobi sweeps), never the converged result; NULL = cold.         */
int uglad_eigh_warm(const float* A, int B, int D, int shift_mode, float* w, float* Vt, float* info,
                    float* scratch, const float* warm_Vt, const float* warm_w, void* stream);
int uglad_condition_covariance_warm(float* S, int B, int D, float offset, float* wS, float* VtS,
                                    float* info, float* scratch, const float* warm_Vt, const float* warm_w,
                                    void* stream);

/* prepare_data.py:345-355: eigen-decompose S; if min eig <= 1e-6 add (offset - min) to the
 * diagonal of S (in place) and to wS.  wS / VtS are reused by every later forward.        */
int uglad_condition_covariance(float* S, int B, int D, float offset, float* wS, float* VtS,
                               float* info, float* scratch, void* stream);
/* The reference decides on float64 eigenvalues; here the smallest eigenvalue is refined in double
 * before the test (D <= uglad_small_d_max(): Rayleigh quotient of its FP32 eigenvector).
 * D > uglad_small_d_max(): the large-D path keeps no eigendecomposition (wS / VtS / info are
 * ignored and may be NULL); an FP32 Cholesky of S - (1e-6 + band) I certifies the common
 * well-conditioned case, anything closer to the threshold is decided by a float64 Cholesky test and
 * a float64 bisection on the shift.  Host-synchronous.  scratch must hold
 * uglad_condition_scratch_floats(B, D) floats (for D <= small max that equals
 * uglad_eigh_scratch_floats).                                                             */
size_t uglad_condition_scratch_floats(int B, int D);
/* the same with the samples at hand (X[B][M][D] and the column means uglad_covariance wrote): the
 * smallest eigenvalue that the repair decision and the shift use is then that of the float64
 * covariance of the samples themselves -- what the reference computes -- instead of that of the
 * FP32 matrix S.  X / mean may be NULL (then identical to uglad_condition_covariance_warm).   */
int uglad_condition_covariance_x(float* S, const float* X, const float* mean, int B, int M, int D, float offset,
                                 float* wS, float* VtS, float* info, float* scratch, const float* warm_Vt,
                                 const float* warm_w, void* stream);
/* largest D served by the one-CTA eigensolver; above it the theta update runs the reference's
 * Newton-Schulz iteration as dense products (tcgen05 3xTF32) and logdet / inverses come from a
 * blocked Cholesky.  uglad_eigh itself is only available up to this size.                  */
int uglad_small_d_max(void);
/* 1 when a batch of B graphs of size D takes the eigensolver path, 0 for the Newton-Schulz chain: D <=
 * uglad_small_d_max(), and -- while a batch is small enough for every graph's warm solves to spread over a
 * resident cluster of 4 CTAs (csrc/eig_cluster.cu) -- D <= 200 with D % 4 == 0 (configs[3]: 32 x D = 200).
 * It concerns the layers (uglad_glad_*): conditioning, theta_0 and the glasso loss switch at
 * uglad_small_d_max() alone (wS / VtS are NULL above it).                                              */
int uglad_eig_path(int B, int D);

/* (2)+(3) the unrolled model.  The workspace holds everything the backward needs
 * (theta_k1, theta_pred, eigenvectors, eigenvalues per layer) plus scratch.              */
size_t uglad_workspace_floats(const uglad_dims* d);
/* float offset of a named workspace region: "theta" (final theta_pred, [B][D][D]),
 * "theta0", "lambda" ([L]), "normf" ([L] local sums of ||Z-X||_F^2), "sweeps" ([L+1]).
 * Returns (size_t)-1 for an unknown name.                                                 */
size_t uglad_workspace_offset(const uglad_dims* d, const char* name);

/* glad.py:106-135: theta_0 and nothing else. */
int uglad_glad_init_forward(const uglad_dims* d, const float* S, const float* params,
                            const float* wS, const float* VtS, float* ws, void* stream);
/* glad.py:136-150, one iteration k (0-based).  Reads normf[k-1] (which the caller has
 * all-reduced over processes when B_total > B) to form lambda_k, then runs the theta update,
 * the Z update and leaves the local sum of ||Z-X||_F^2 in normf[k].                        */
int uglad_glad_layer_forward(const uglad_dims* d, int k, const float* S, const float* params,
                             float* ws, const float* warm_ws, void* stream);
/* glad.py:74-150 in one call (single process: B_total == B).
 * warm_ws (may be NULL): the workspace of an earlier forward with the SAME dims -- normally the
 * previous epoch's.  Its per-layer eigenvectors seed the Jacobi solver (the matrices differ by
 * one optimiser step, so ~2 sweeps replace ~9).  It only changes the work done, never the
 * converged result; any stale or unrelated workspace is a valid (if useless) seed. */
int uglad_glad_forward(const uglad_dims* d, const float* S, const float* params,
                       const float* wS, const float* VtS, float* ws, const float* warm_ws,
                       void* stream);
/* The same on ONE shard of a graph-sharded batch (B_total > B), still in one call.  The graphs of a
 * multitask / consensus batch are coupled only by the batch mean of ||Z - X||_F^2 after every layer
 * (glad.py:147).  Every rank owns an exchange buffer of uglad_peer_slots_bytes(L) bytes that all ranks
 * of the box have mapped (uglad_peer_alloc / uglad_peer_open: CUDA IPC over NVLink peer access); the
 * lambda kernel of layer k + 1 stores its local sum into every rank's buffer and adds up what the
 * other ranks stored into its own -- in rank order, so every rank computes the same lambda bit for
 * bit.  `tag` must be the same on all ranks and different from the previous call's (a call counter).
 * Between two calls the ranks must have met (the gradient all-reduce of the epoch, or a barrier).    */
#define UGLAD_MAX_PEERS 8
typedef struct uglad_peers {
  int world;                      /* ranks sharing the batch (<= UGLAD_MAX_PEERS, one box)           */
  int rank;                       /* this process                                                     */
  unsigned tag;                   /* call counter, identical on all ranks                             */
  void* slots[UGLAD_MAX_PEERS];   /* device pointers (valid in THIS process) to every rank's buffer  */
  void* tag_dev;                  /* optional device uint32 (zero-initialised, private to this rank): when set
                                   * the counter lives there (layer 0 increments it) and `tag` is ignored, so
                                   * that a captured CUDA graph of the call can be replayed                  */
} uglad_peers;
size_t uglad_peer_slots_bytes(int L);
int uglad_glad_forward_sharded(const uglad_dims* d, const float* S, const float* params, const float* wS,
                               const float* VtS, float* ws, const float* warm_ws, const uglad_peers* peers,
                               void* stream);
/* exchange-buffer plumbing: cudaMalloc + zero + cudaIpcGetMemHandle (64 bytes, to be sent to the
 * other ranks) / cudaIpcOpenMemHandle of a peer's handle / release (opened = 1 for peer mappings).  */
int uglad_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64);
int uglad_peer_open(const unsigned char* handle64, void** ptr);
int uglad_peer_close(void* ptr, int opened);
/* backward of the above: grad_theta[B][D][D] -> grad_params[uglad_param_count(H)] (local
 * contribution; the caller all-reduces it over processes).                               */
int uglad_glad_backward(const uglad_dims* d, const float* S, const float* params,
                        const float* wS, const float* VtS, float* ws,
                        const float* grad_theta, float* grad_params, void* stream);

/* (4) main.py:289-315 loss_uGLAD: loss_out[0] = sum_b(-logdet theta_b + <S_b, theta_b>) / Bdiv
 * and grad_theta (may be NULL) = (-theta^-1 + S) / Bdiv.  S_batch is 1 (broadcast, the
 * consensus mode of main.py:620-622) or B.  scratch: uglad_loss_scratch_floats(B, D).     */
size_t uglad_loss_scratch_floats(int B, int D);
int uglad_glasso_loss(const float* theta, const float* S, int B, int D, int S_batch, float Bdiv,
                      float* loss_out, float* grad_theta, float* scratch, void* stream);
/* the same plus the optional structure prior of main.py:325-334 (struct_theta[B][D][D], may be NULL):
 * loss += sum log cosh(theta o mask) / Bdiv, mask = (1 - struct_theta) - I, and grad_theta +=
 * tanh(theta o mask) o mask / Bdiv.                                                          */
int uglad_glasso_loss_prior(const float* theta, const float* S, const float* struct_theta, int B, int D,
                            int S_batch, float Bdiv, float* loss_out, float* grad_theta, float* scratch,
                            void* stream);

/* instrumentation for bench.py: kernels launched by this library since it was loaded, and
 * CUDA-event timing of the dominant kernel (the Jacobi eigensolver): uglad_profile(enable, ..)
 * returns the time and launch count accumulated since the previous call, then switches the
 * event brackets on or off.                                                               */
unsigned long long uglad_launch_count(void);
int uglad_profile(int enable, double* total_ms, unsigned long long* launches);
/* read (without clearing) what has accumulated since the last uglad_profile call for one kernel
 * class: kind 0 = Jacobi eigensolver (work = algorithmic HBM bytes), kind 1 = tcgen05 3xTF32 GEMM
 * (work = algorithmic flops 2 M N K batch), kind 2 = the FP32 SIMT flops the eigensolver launches
 * counted themselves (2 D per column-pair dot product, 8 D per applied rotation; launches = rotations). */
int uglad_profile_read(int kind, double* total_ms, unsigned long long* launches, double* work);

/* developer knobs used by the tuning scripts and tests: "eig_lp" (lanes per column pair, 0 = auto),
 * "eig_keepg" (-1 auto / 0 / 1: keep A + sigma I in a second shared-memory buffer),
 * "small_d_max" (threshold between the eigensolver path and the large-D path, <= 232),
 * "use_tc" (1: tcgen05 3xTF32 products in the large-D path, 0: FP32 SIMT products),
 * "tc_bn" (tile width of the tcgen05 kernel: 0 auto / 64 / 112 / 128),
 * "tc_raw" (1, default: the large-D chain keeps plain FP32 matrices and the tcgen05 kernel forms
 * hi/lo in shared memory; 0: pre-split hi/lo pairs in HBM), "tc_dual" / "tc_pdl" (pair independent
 * products into one launch / programmatic dependent launch, both 1 by default).              */
int uglad_tune(const char* key, int value);

/* developer timeline of the tcgen05 GEMM: buf = device int64[148][16] (NULL switches it off); each
 * CTA stamps clock64 at {start, setup done, dependency wait done, first operands landed, last MMA
 * committed, accumulator visible to the epilogue, epilogue done, start (wall clock)} for its first
 * tile, followed by cycle sums of the MMA thread {wait, issue, commit}, of the split warps {wait,
 * -, convert} and of one epilogue warp {tcgen05.ld, math + stores}.                           */
int uglad_tc_debug_buffer(void* buf);
/* developer benchmark: `reps` back-to-back launches of the product (operands pre-split once when
 * "tc_raw" is 0; split_out: write the hi/lo pair into scratch instead of C).                 */
int uglad_tc_gemm_repeat(const float* A, const float* B, float* C, int M, int N, int K, int batch, int reps,
                         int split_out, float* scratch, void* stream);

/* building blocks exported for the parity tests */
/* the tcgen05 3xTF32 product of the large-D path on plain operands:
 * C[b] = alpha A[b] B[b]^T + beta E1[b] + diag I, A [batch][M][K], B [batch][N][K], C / E1 [batch][M][N]
 * (E1 may be NULL).  scratch: uglad_tc_gemm_scratch_floats floats (the hi/lo split operands; unused
 * when "tc_raw" is 1, K % 4 == 0 and the operands are 16-byte aligned: they are then read in place). */
size_t uglad_tc_gemm_scratch_floats(int M, int N, int K, int batch);
int uglad_tc_gemm(const float* A, const float* B, const float* E1, float* C, int M, int N, int K, int batch,
                  float alpha, float beta, float diag, float* scratch, void* stream);
int uglad_z_update(const float* X, const float* S, const float* theta_prev, const float* params,
                   int H, int B, int D, float* Z, float* normf_out, float* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif
