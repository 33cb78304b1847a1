/*
 * vspectra.h -- C-ABI of the B200 weight-spectrum analysis path.
 *
 * This is the drop-in boundary for the reference's per-matrix hot loop.  The
 * reference has no FFI of its own (pure Python over SciPy/LAPACK); each entry
 * point below names the reference call it replaces, file:line relative to the
 * reference repository root (vision_spectra/...):
 *
 *   vsp_analyze_batch / vsp_plan_execute
 *       replaces the serial loop   experiments/run_spectral_analysis.py:323-336
 *       and its twin               training/base.py:399-405
 *       i.e. per matrix: get_spectral_metrics (metrics/spectral.py:371-414 ->
 *       spectral_entropy :49, stable_rank :112, alpha_exponent :176,
 *       power_law_alpha_hill :276) plus the fifth SVD that stores the singular
 *       values (run_spectral_analysis.py:331-334).
 *   vsp_analyze_batch_host
 *       same, for callers that hold HOST matrices (what the reference has after
 *       `.detach().cpu().numpy()`, metrics/extraction.py:56,101,142,184,226).
 *   vsp_record
 *       carries the four floats of get_spectral_metrics' dict (spectral.py:409-414)
 *       and the integers its estimators derive (m, OLS window, Hill k;
 *       spectral.py:245-256, 346-353).
 *
 * Conventions:
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - device entry points are asynchronous on `stream` (a cudaStream_t passed as
 *     void*), never synchronise, keep no global state and are re-entrant across
 *     streams as long as each call has its own workspace and outputs.
 *   - return value: 0 or a negative VSP_E_* code for CALL-level errors only.
 *     DATA-level failures (NaN/Inf in a matrix, all-zero matrix, too few
 *     singular values) are per-matrix: vsp_record.status + NaN metrics, exactly
 *     like the reference's "return np.nan, never raise" (spectral.py:87-105).
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     returns VSP_E_CUDA.
 */
#ifndef VSPECTRA_H
#define VSPECTRA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSP_VERSION 100 /* 0.1.0 */

/* call-level error codes */
#define VSP_OK 0
#define VSP_E_ARG (-1)       /* null pointer, count < 0, rows/cols < 1, ld < cols */
#define VSP_E_UNSUPPORTED (-2) /* min(rows,cols) > VSP_MAX_N or unknown dtype */
#define VSP_E_WORKSPACE (-3) /* workspace_bytes too small */
#define VSP_E_CUDA (-4)      /* a CUDA runtime call failed (see vsp_last_cuda_error) */
#define VSP_E_ALLOC (-5)     /* host allocation failed */

#define VSP_MAX_N 4096 /* largest min(rows, cols) */

/* element type of the input matrices */
#define VSP_F32 0
#define VSP_F64 1

/* per-matrix status bits (vsp_record.status); 0 = all four metrics finite */
#define VSP_ST_NONFINITE 1 /* NaN/Inf in the input -> all metrics NaN (SciPy check_finite) */
#define VSP_ST_ZERO 2      /* no positive singular value -> all metrics NaN           */
#define VSP_ST_FEW_SV 4    /* fewer than 8 positive SVs -> alpha and Hill NaN         */
#define VSP_ST_ALPHA_NAN 8 /* OLS window rejected (fit_range) or slope not finite     */
#define VSP_ST_HILL_NAN 16 /* Hill mean-log <= 0 or not finite                        */
#define VSP_ST_REFINED 32  /* singular values re-solved from W itself (FP64 bidiagonalisation) */
#define VSP_ST_ILLCOND 64  /* Gram route saw lambda_min/lambda_max < 1e-9 (kappa > ~3e4)  */

/* Options; every field -1 selects the reference's default. */
typedef struct vsp_opts {
    int32_t fit_start; /* alpha_exponent(fit_range=(start,end))  spectral.py:259-262 */
    int32_t fit_end;
    int32_t hill_k;    /* power_law_alpha_hill(k=...)            spectral.py:351     */
    int32_t want_sv;   /* 0: skip the singular-value output; default (-1) = write    */
    int32_t refine;    /* 0: never run the ill-conditioned re-solve; default = auto  */
    int32_t dist_k;    /* > 0: also write the first dist_k entries of the four distribution arrays of
                          get_spectral_distribution (spectral.py:545-557) per matrix -- singular values,
                          eigenvalues s^2, normalized_sv s/s_0, cumulative_variance cumsum(s^2)/sum(s^2) -- i.e.
                          SpectralTracker's max_singular_values truncation (spectral.py:683-692) done on the
                          device: vsp_plan_execute_dist.  0 / -1: off */
    int32_t clauset;   /* 1: also run the Clauset-Shalizi-Newman x_min scan on the eigenvalue spectrum (BASELINE
                          north_star stage 3; the reference has no such scan -- SURVEY D1 -- so this output is an
                          extra with its own oracle, oracle/spectral_oracle.py: clauset_xmin_scan): for EVERY
                          candidate cutoff x_min = lambda_(k) the continuous MLE alpha = 1 + t / sum ln(lambda_i / x_min)
                          over the t eigenvalues >= x_min and the Kolmogorov-Smirnov distance of the fitted tail; the
                          cutoff with the smallest distance wins.  Output through vsp_plan_execute_dist.  0: off */
    int32_t reserved[1];
} vsp_opts;

/* Doubles per matrix of the auxiliary output of vsp_plan_execute_dist: 4 * dist_k (rows singular_values, eigenvalues,
 * normalized_sv, cumulative_variance) followed, when opts.clauset is set, by 8 doubles
 * { alpha, x_min, ks_distance, x_min index (0 = largest eigenvalue), tail count t, 0, 0, 0 } (NaN / -1 when fewer
 * than 8 positive eigenvalues). */
#define VSP_AUX_STRIDE(dist_k, clauset) (4 * ((dist_k) > 0 ? (dist_k) : 0) + ((clauset) ? 8 : 0))

/* One result record per matrix: 64 bytes, the unit that is gathered across GPUs. */
typedef struct vsp_record {
    int32_t item;    /* index of the matrix in the batch                          */
    int32_t status;  /* VSP_ST_* bits                                             */
    int32_t m;       /* positive finite singular values      spectral.py:243-245 */
    int32_t start;   /* OLS window [start,end), -1 if none    spectral.py:254-256 */
    int32_t end;
    int32_t k;       /* Hill tail count, -1 if none           spectral.py:352-353 */
    int32_t n;       /* min(rows, cols)                                           */
    int32_t iters;   /* bisection iterations (max over eigenvalues)               */
    double metrics[4]; /* spectral_entropy, stable_rank, alpha_exponent, pl_alpha_hill */
} vsp_record;

typedef struct vsp_plan vsp_plan; /* opaque: shape tables of one batch, resident on the device */

int vsp_version(void);
const char* vsp_error_string(int code);
const char* vsp_last_cuda_error(void); /* text of the last CUDA failure on this thread */

/* Upper bound of the device workspace any plan over these shapes needs (Gram + band /
 * tridiagonal scratch + digit planes + the FP64 pool of the ill-conditioned re-solve); it is
 * laid out by the same code as vsp_plan_create, so vsp_analyze_batch never asks for more.
 * Negative = VSP_E_* code. */
int64_t vsp_workspace_bytes(int32_t count, const int32_t* rows, const int32_t* cols);

/* Offsets (in doubles) of each matrix's singular values inside the packed SV
 * output: sv_offsets[i] .. sv_offsets[i] + min(rows[i], cols[i]).  sv_offsets
 * has count+1 entries; the last is the total. */
int vsp_sv_offsets(int32_t count, const int32_t* rows, const int32_t* cols, int64_t* sv_offsets);

/* Plan API: shapes are validated, bucketed by shape class and uploaded once; the
 * same plan can be executed any number of times (e.g. once per epoch on the live
 * parameters), ONE EXECUTION AT A TIME: an execution rewrites the plan's item pointers
 * and uses the plan's own fork/join events and side stream, so two streams that want to
 * run the same shapes concurrently need one plan each.  All host arrays are read during
 * the call only.  fp32 matrices whose contraction length max(rows, cols) exceeds 65 536
 * take the FP64 Gram kernel instead of the int8 tensor-core split (whose int32 level sums
 * are proven exact up to that length); results are the same to working precision. */
int vsp_plan_create(int32_t count, const int32_t* rows, const int32_t* cols,
                    const int64_t* ld /* row stride in elements, NULL = cols */,
                    int32_t dtype, const vsp_opts* opts /* NULL = defaults */,
                    vsp_plan** out_plan);
int64_t vsp_plan_workspace_bytes(const vsp_plan* plan);
int64_t vsp_plan_sv_count(const vsp_plan* plan);
void vsp_plan_destroy(vsp_plan* plan);

/* Launch the three stages for every matrix of the plan on `stream`.
 *   d_ptrs      HOST array [count] of DEVICE pointers to row-major matrices
 *   d_sv        DEVICE f64 [vsp_plan_sv_count], descending per matrix (may be NULL if want_sv == 0)
 *   d_records   DEVICE vsp_record [count]
 *   d_workspace DEVICE scratch, 256-byte aligned, >= vsp_plan_workspace_bytes
 * Inputs are borrowed and never written. */
int vsp_plan_execute(vsp_plan* plan, const void* const* d_ptrs, double* d_sv,
                     vsp_record* d_records, void* d_workspace, int64_t workspace_bytes,
                     void* stream);

/* vsp_plan_execute for plans created with opts.dist_k > 0 and / or opts.clauset: additionally fills
 *   d_dist      DEVICE f64 [count][VSP_AUX_STRIDE(dist_k, clauset)]: per matrix (batch order) the rows
 *               singular_values, eigenvalues, normalized_sv, cumulative_variance, truncated to dist_k entries (entries
 *               beyond min(rows, cols) and every entry of a matrix with non-finite input are NaN), then the
 *               Clauset block.
 * d_sv may be NULL when the plan has want_sv == 0: a tracker epoch then moves 4 dist_k values per matrix
 * instead of min(rows, cols). */
int vsp_plan_execute_dist(vsp_plan* plan, const void* const* d_ptrs, double* d_sv,
                          vsp_record* d_records, double* d_dist, void* d_workspace,
                          int64_t workspace_bytes, void* stream);

/* vsp_plan_execute with CUDA events around each stage.  Synchronises `stream` and
 * returns the device time per stage, summed over shape classes:
 * stage_ms[0] Gram, [1] tridiagonalisation, [2] bisection + metrics.  Measurement aid
 * for bench.py's per-kernel roofline; the results are identical to vsp_plan_execute. */
int vsp_plan_execute_profiled(vsp_plan* plan, const void* const* d_ptrs, double* d_sv,
                              vsp_record* d_records, void* d_workspace, int64_t workspace_bytes,
                              void* stream, float* stage_ms /* [3] */);

/* Validation aid: run stage 1 only and expand every Gram matrix (smaller side, f64) into a
 * dense n*n block of d_out (blocks in batch order, n_i*n_i doubles each).  Synchronises. */
int vsp_plan_debug_gram(vsp_plan* plan, const void* const* d_ptrs, double* d_out, void* d_workspace,
                        int64_t workspace_bytes, void* stream);

/* One-shot form of create + execute + destroy. */
int vsp_analyze_batch(const void* const* d_ptrs, const int32_t* rows, const int32_t* cols,
                      const int64_t* ld, int32_t dtype, int32_t count, const vsp_opts* opts,
                      double* d_sv, vsp_record* d_records, void* d_workspace,
                      int64_t workspace_bytes, void* stream);

/* Host-buffer form: matrices and results live in HOST memory.  The call copies the
 * matrices to the device straight from the caller's pointers (runs of host-contiguous
 * matrices as one cudaMemcpyAsync, strided views as one 2-D copy; pinned caller memory
 * makes these asynchronous DMA transfers, pageable memory is staged by the driver), runs
 * the three stages and copies records (and SVs) back; it returns after the results are
 * in h_sv / h_records.  `device` is the CUDA ordinal to use. */
int vsp_analyze_batch_host(const void* const* h_ptrs, const int32_t* rows, const int32_t* cols,
                           const int64_t* ld, int32_t dtype, int32_t count,
                           const vsp_opts* opts, double* h_sv, vsp_record* h_records,
                           int32_t device);

/* Batched FP64 GEMM on the FP64 tensor cores, the building block of the singular-vector consumers (SURVEY 8f rank 4:
 * metrics/tail_truncation.py:63-152, metrics/gradient_alignment.py:48-70 -- evaluated as Newton-Schulz matrix
 * functions of W, vision_spectra_b200/lowrank.py):
 *     C[b] = alpha * op(A[b]) * op(B[b]) + beta * C[b] + gamma * I,   op(X) = X or X^T (trans flag), row-major f64,
 * M x N results, contraction length K.  d_A / d_B / d_C are DEVICE arrays of `batch` DEVICE pointers (batch <= 65535);
 * C must not alias A or B.  Asynchronous on `stream`. */
int vsp_dgemm_batched(int32_t batch, int32_t M, int32_t N, int32_t K, double alpha,
                      const double* const* d_A, int64_t lda, int32_t transA,
                      const double* const* d_B, int64_t ldb, int32_t transB,
                      double beta, double gamma, double* const* d_C, int64_t ldc, void* stream);

/* Counters for bench.py's `gpu_launches`: kernels launched by this library in this
 * process since load (or since the last reset). */
int64_t vsp_kernel_launch_count(void);
void vsp_reset_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VSPECTRA_H */
