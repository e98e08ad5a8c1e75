/*
 * mbar_b200.h — C ABI of libmbar_b200.so: the MBAR solve hot path on NVIDIA B200 (sm_100a).
 *
 * Drop-in boundary.  pymbar has no native FFI; its "operator API" for this path is the set of
 * module-level functions in pymbar/mbar_solvers.py that pymbar.MBAR reaches by attribute lookup
 * (mbar.py:413, :437, :455, :910).  Each entry point below names the reference function it stands
 * in for (paths relative to /root/reference/pymbar/).  INTEGRATION.md shows the ctypes stub a
 * pymbar maintainer would add beside the numpy/JAX switch at mbar_solvers.py:25-87.
 *
 * Conventions
 *   - plain C types only; every pointer is caller-owned HOST memory unless the name says "_dev";
 *   - all functions return 0 (MBAR_B200_OK) or a negative mbar_b200_status; no exceptions, no
 *     callbacks; mbar_b200_last_error() gives the message of the last failure on this thread;
 *   - one context = one GPU = one contiguous slice [n0, n0+N_local) of the samples.  K-sized
 *     state (N_k, f_k) is global and replicated.  With a communicator attached
 *     (mbar_b200_comm_init) every reduction over samples is followed by one sum all-reduce of the
 *     packed partials, so every rank returns identical K-vectors;
 *   - u_kn is float64 [K, N] row-major on the host exactly as pymbar.MBAR holds it (mbar.py:243).
 *     In HBM it is re-tiled to [N/32][K][32] and shifted per sample (see DESIGN.md "HBM layout");
 *   - states with N_k == 0 are "unsampled": they never enter a denominator and only
 *     mbar_b200_self_consistent_update / mbar_b200_log_W_nk produce values for them, exactly as in
 *     the reference (mbar_solvers.py:1002-1012).
 */
#ifndef MBAR_B200_H
#define MBAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBAR_B200_ABI_VERSION 1

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct mbar_b200_ctx mbar_b200_ctx;

typedef enum mbar_b200_status {
    MBAR_B200_OK = 0,
    MBAR_B200_ERR_INVALID = -1,      /* bad argument (shape, NULL, K<1, ...)  -> ParameterError/ValueError */
    MBAR_B200_ERR_CUDA = -2,         /* CUDA runtime failure                                             */
    MBAR_B200_ERR_NO_DEVICE = -3,    /* no usable sm_100 GPU: there is NO CPU fallback                    */
    MBAR_B200_ERR_NOT_READY = -4,    /* u_kn not uploaded / synthesised yet                               */
    MBAR_B200_ERR_NAN = -5,          /* NaN found in u_kn at upload                                       */
    MBAR_B200_ERR_RANGE = -6,        /* f_k + log N_k outside the supported +-1e6 range                   */
    MBAR_B200_ERR_COMM = -7,         /* NCCL failure or libnccl not loadable                              */
    MBAR_B200_ERR_SINGULAR = -8,     /* Newton system not positive definite (disconnected states)         */
    MBAR_B200_ERR_NOMEM = -9
} mbar_b200_status;

/* Which kernel family a pass uses (mbar_b200_set_pass_kernel). AUTO = fused when it applies. */
typedef enum mbar_b200_kernel {
    MBAR_B200_KERNEL_AUTO = 0,
    MBAR_B200_KERNEL_FUSED = 1,      /* TMA-pipelined persistent kernel (clusters of CTAs above K = 256), K <= 2048 */
    MBAR_B200_KERNEL_GENERIC = 2     /* any K, log-domain for unsampled states                            */
} mbar_b200_kernel;

/* Result of a native solve (reference: the `results` dict of adaptive(), mbar_solvers.py:662-665). */
typedef struct mbar_b200_solve_result {
    int32_t success;          /* results["success"]                                                   */
    int32_t iterations;       /* loop iterations executed                                             */
    int32_t nr_iterations;    /* Newton-Raphson steps taken (nr_iter, mbar_solvers.py:620)            */
    int32_t sci_iterations;   /* self-consistent steps taken (sci_iter, mbar_solvers.py:610)          */
    int32_t passes;           /* full streaming passes over u_kn                                      */
    int32_t hessian_passes;   /* of which with the K x K Hessian                                      */
    double max_delta;         /* last relative change (mbar_solvers.py:631)                           */
    double gnorm;             /* ||gradient||_2 at the returned f_k (mbar_solvers.py:938-940)          */
    double device_ms;         /* CUDA-event time of the whole solve on this rank                      */
} mbar_b200_solve_result;

/* Synthetic-input family of SURVEY.md 8(d): harmonic oscillators, samples in block order. */
typedef struct mbar_b200_synth {
    uint64_t seed;            /* Philox-4x32-10 key; counter = GLOBAL sample index                     */
    int64_t n_offset;         /* global index of this context's first sample                           */
    int64_t N_global;         /* total samples over all ranks (defines the state of origin of n)       */
    const double* O_k;        /* [K] oscillator centres                                                */
    const double* k_k;        /* [K] spring constants (beta = 1)                                       */
} mbar_b200_synth;

/* ---- library ------------------------------------------------------------------------------- */
int mbar_b200_abi_version(void);
const char* mbar_b200_last_error(void);
int mbar_b200_device_count(int* count);
/* Pinned host memory for zero-staging uploads/downloads (cudaHostAlloc / cudaFreeHost). */
int mbar_b200_host_alloc(void** ptr, uint64_t bytes);
int mbar_b200_host_free(void* ptr);
/* NUMA node the GPU hangs off (sysfs), -1 when the host exposes none.  Pinned staging buffers are allocated with
 * that node preferred so uploads / downloads do not cross the inter-socket link. */
int mbar_b200_gpu_numa_node(int device, int* node);

/* 64-bit content hash of a host matrix of `rows` rows of `row_bytes` bytes, row stride `stride_bytes` (threaded,
 * memory-bandwidth bound, independent of the thread count).  A binding that keeps u_kn resident between the
 * reference's pure-function calls keys its cache on this, never on a sample of the contents. */
int mbar_b200_host_hash(const void* base, int64_t rows, int64_t row_bytes, int64_t stride_bytes, uint64_t* hash_out);

/* Release the buffers parked by destroyed contexts (at most one u_kn-sized device buffer and one staging set
 * per device are kept for the next context; MBAR_B200_NO_POOL=1 in the environment disables the parking). */
int mbar_b200_trim(void);

/* ---- context ------------------------------------------------------------------------------- */
/* N_k: [K] global sample counts as float64 (validate_inputs casts to float, mbar_solvers.py:198). */
int mbar_b200_create(mbar_b200_ctx** ctx, int device, int32_t K, int64_t N_local, const double* N_k);
int mbar_b200_destroy(mbar_b200_ctx* ctx);
int mbar_b200_get_shape(const mbar_b200_ctx* ctx, int32_t* K, int64_t* N_local);
int mbar_b200_set_pass_kernel(mbar_b200_ctx* ctx, int kernel /* mbar_b200_kernel */);
/* Counters since creation: kernel launches, streaming passes, bytes moved H2D / D2H. */
int mbar_b200_get_counters(const mbar_b200_ctx* ctx, int64_t* launches, int64_t* passes,
                           int64_t* h2d_bytes, int64_t* d2h_bytes);
/* CUDA-event duration (ms) of the most recent pass kernel on the context's stream. */
int mbar_b200_last_pass_ms(mbar_b200_ctx* ctx, double* ms);

/* ---- data in / out -------------------------------------------------------------------------- */
/* Replaces the host copies at mbar.py:243 and mbar_solvers.py:1003: u_host is [K, N_local]
 * row-major with row stride `ld` (elements).  Pinned memory is DMA'd directly; pageable memory is
 * staged through internal pinned buffers.  NaN anywhere -> MBAR_B200_ERR_NAN. */
int mbar_b200_upload_u_kn(mbar_b200_ctx* ctx, const double* u_host, int64_t ld);
/* Same, from a row-major DEVICE buffer on ctx's device (e.g. a torch tensor's data_ptr). */
int mbar_b200_upload_u_kn_dev(mbar_b200_ctx* ctx, const double* u_dev, int64_t ld);
/* A new context holding the samples of `base` plus n_extra UNSAMPLED states whose energies are u_extra_host
 * [n_extra, N_local] (row stride ld).  The resident tiles are copied device-to-device; only the new rows cross
 * PCIe.  This is how expectations / perturbed free energies (the columns mbar.py:886-940 appends to Log_W_nk)
 * reuse the resident u_kn.  `base` stays valid and independent. */
int mbar_b200_create_augmented(mbar_b200_ctx* base, int32_t n_extra, const double* u_extra_host, int64_t ld,
                               mbar_b200_ctx** ctx_out);
/* Fill the context on device from the synthetic family (no host traffic). */
int mbar_b200_synthesize(mbar_b200_ctx* ctx, const mbar_b200_synth* spec);
/* Per-sample multiplicities w_n >= 0 ([N_local] host doubles; NULL restores w_n = 1).  Every sum over
 * samples becomes a weighted sum (S_k = sum_n w_n W_nk, sum_n w_n L_n, W^T diag(w) W) while the denominators
 * L_n are untouched: a bootstrap replicate of mbar.py:417-449 (`u_kn[:, rints]`, an 8*K*N gather per
 * replicate) is the same data with w_n = number of times sample n was drawn — no copy, same kernels. */
int mbar_b200_set_sample_weights(mbar_b200_ctx* ctx, const double* w_host);
/* Read back columns [n0, n0+n) of the ORIGINAL (unshifted) u_kn as [K, n] row-major, stride ld. */
int mbar_b200_download_u_kn(mbar_b200_ctx* ctx, int64_t n0, int64_t n, double* u_host, int64_t ld);

/* ---- the streaming pass and the reference primitives built on it ---------------------------- */
/* One read of u_kn at f_k.  Outputs (any may be NULL):
 *   S[K]    S_k = sum_n W_nk  (0 for unsampled states)
 *   sumL    sum_n L_n,  L_n = log sum_k N_k exp(f_k - u_kn)   (mbar_solvers.py:238)
 *   G[K*K]  G = W^T W row-major (0 rows/cols for unsampled states); costs the Hessian pass.
 * With a communicator the outputs are the all-reduced global sums. */
int mbar_b200_pass(mbar_b200_ctx* ctx, const double* f_k, double* S, double* sumL, double* G);

/* The same pass at M (1 or 2) candidate vectors f[m][K] in one call, one host synchronisation: what adaptive()
 * needs to compare its self-consistent and Newton-Raphson candidates (mbar_solvers.py:590-607 evaluates
 * mbar_gradient at both).  S is [M][K], sumL is [M]; either may be NULL. */
int mbar_b200_pass_multi(mbar_b200_ctx* ctx, int32_t M, const double* f, double* S, double* sumL);

/* self_consistent_update(u_kn, N_k, f_k)  — mbar_solvers.py:206-257, Eq. C3, ALL states. */
int mbar_b200_self_consistent_update(mbar_b200_ctx* ctx, const double* f_k, double* f_out);
/* mbar_gradient — mbar_solvers.py:260-292, Eq. C6.  Unsampled states get 0 (-N_k * ...). */
int mbar_b200_gradient(mbar_b200_ctx* ctx, const double* f_k, double* g_out);
/* mbar_objective_and_gradient — mbar_solvers.py:341-392 (g_out may be NULL = mbar_objective). */
int mbar_b200_objective_and_gradient(mbar_b200_ctx* ctx, const double* f_k, double* obj_out,
                                     double* g_out);
/* mbar_hessian — mbar_solvers.py:395-436, Eq. C9, [K, K] row-major. */
int mbar_b200_hessian(mbar_b200_ctx* ctx, const double* f_k, double* H_out);
/* mbar_log_W_nk — mbar_solvers.py:439-473: [N_local, K] row-major (note: transposed w.r.t. u_kn),
 * row stride ld_out elements; exponentiate != 0 gives mbar_W_nk (mbar_solvers.py:476-507). */
int mbar_b200_log_W_nk(mbar_b200_ctx* ctx, const double* f_k, double* logW_host, int64_t ld_out,
                       int exponentiate);
/* Rows [n0, n0 + n) of the same matrix (n0 a multiple of 32): lets a caller page through Log_W_nk, or fetch the
 * part an estimator needs, instead of materialising N x K doubles on the host (mbar.py:455 keeps all of it). */
int mbar_b200_log_W_nk_rows(mbar_b200_ctx* ctx, const double* f_k, int64_t n0, int64_t n, double* logW_host,
                            int64_t ld_out, int exponentiate);
/* Sums and second moments of the weights of ALL K states (sampled or not) at f_k:
 *   S[K] = sum_n W_nk,  G[K*K] = W^T W.
 * Everything MBAR's asymptotic covariance (mbar.py:1837-1858, "svd-ew"), compute_overlap (mbar.py:605-606) and
 * compute_effective_sample_number (mbar.py:546-549) need, without materialising the N x K weight matrix. */
int mbar_b200_weight_moments(mbar_b200_ctx* ctx, const double* f_k, double* S_out, double* G_out);
/* Per-sample log denominators L_n [N_local] (the logsumexp at mbar_solvers.py:238). */
int mbar_b200_log_denominator(mbar_b200_ctx* ctx, const double* f_k, double* L_host);

/* ---- native solver loops (no Python between iterations) ------------------------------------- */
/* Plain self-consistent iteration f <- f - log S(f), gauge f[first sampled] = 0 each step, until
 * max |delta f / f| < tol (the convergence rule of mbar_solvers.py:627-640) or maxiter. */
int mbar_b200_solve_sci(mbar_b200_ctx* ctx, double* f_inout, double tol, int32_t maxiter,
                        mbar_b200_solve_result* result);
/* adaptive() — mbar_solvers.py:510-667: Newton vs self-consistent step by gradient norm.
 * f_inout covers all K states; unsampled states are carried through untouched. */
int mbar_b200_solve_adaptive(mbar_b200_ctx* ctx, double* f_inout, double tol, int32_t maxiter,
                             int32_t min_sc_iter, double gamma, mbar_b200_solve_result* result);
/* How the two solvers above iterate.  mode 0 (default): device resident — every quantity of an iteration
 * (gradient, candidates, K x K Cholesky of the Newton system, step choice mbar_solvers.py:607, convergence test
 * :627-640) stays on the GPU, `batch` iterations are enqueued between two polls of a small state struct and
 * kernels of iterations past convergence exit immediately; jax_core_adaptive (mbar_solvers.py:670-694) is the
 * reference's counterpart of that single fused step.  mode 1: host-stepped (one round trip per pass), which is
 * also what mode 0 falls back to when the fast kernels cannot represent an iterate.  batch < 1 keeps the
 * current batch size. */
int mbar_b200_set_loop_mode(mbar_b200_ctx* ctx, int32_t mode, int32_t batch);
/* Host synchronisations spent polling the device-resident loops since creation, current mode and batch. */
int mbar_b200_get_loop_stats(const mbar_b200_ctx* ctx, int64_t* polls, int32_t* mode, int32_t* batch);
/* The adaptive iteration is captured into a CUDA graph after the first batch a context runs and relaunched once
 * per iteration: how often it was captured / launched (MBAR_B200_NO_GRAPH=1 disables the capture). */
int mbar_b200_get_graph_stats(const mbar_b200_ctx* ctx, int64_t* captures, int64_t* launches);
/* Run exactly `iters` self-consistent passes back to back with no host round trip (bench). */
int mbar_b200_sci_iterate(mbar_b200_ctx* ctx, double* f_inout, int32_t iters);

/* CUDA-event timing of the last mbar_b200_sci_iterate on the context's stream: whole loop (pass +
 * all-reduce + K-vector epilogue per iteration) and the sum of the pass-kernel launch durations. */
int mbar_b200_last_loop_ms(mbar_b200_ctx* ctx, double* total_ms, double* kernel_ms_sum, int32_t* iters);

/* ---- introspection for the benchmark ---------------------------------------------------------- */
/* Human-readable description of the pass-kernel / Hessian-kernel variants launched last (buffers of `len`). */
int mbar_b200_last_kernels(const mbar_b200_ctx* ctx, char* pass_kernel, char* hessian_kernel, int32_t len);
/* CUDA-event durations of the last Hessian evaluation: weight materialisation and DMMA kernel + reduction. */
int mbar_b200_last_hessian_ms(mbar_b200_ctx* ctx, double* weights_ms, double* hessian_ms);
/* fp64 ceilings of this GPU measured in place (register-only DMMA.8x8x4 and DFMA loops), in TFLOP/s: the
 * roofline denominator of mbar_b200_hessian (MEASURED_PEAKS.json has no fp64 figure). */
int mbar_b200_measure_fp64_peak(int device, double* dmma_tflops, double* dfma_tflops);

/* ---- one-shot, host-buffer entry (what a binding without residency would call) --------------- */
/* self_consistent_update on host buffers: upload u_kn, one pass, f_out — copies inside the call. */
int mbar_b200_self_consistent_update_host(int device, int32_t K, int64_t N, const double* u_host,
                                          int64_t ld, const double* N_k, const double* f_k,
                                          double* f_out);

/* ---- multi-GPU: samples sharded over ranks, one all-reduce per pass -------------------------- */
#define MBAR_B200_UNIQUE_ID_BYTES 128
int mbar_b200_comm_unique_id(void* id_out /* [128] */);
/* Collective over all ranks; call it after this rank's u_kn is uploaded / synthesised. */
int mbar_b200_comm_init(mbar_b200_ctx* ctx, int32_t nranks, int32_t rank, const void* unique_id);
int mbar_b200_comm_destroy(mbar_b200_ctx* ctx);

/* Peer-memory exchange for the device-resident iteration: every rank exports a 64-byte cudaIpc handle
 * of its inbox, the caller distributes them (any transport), every rank attaches all of them.  After
 * that mbar_b200_sci_iterate runs ONE kernel per iteration: the pass kernel's last CTA stores its
 * K+2 partial sums into every peer's inbox over NVLink, waits for the peers' flags, sums in rank order
 * and applies the K-vector update — no NCCL call and no second launch on the iteration's critical path. */
#define MBAR_B200_IPC_HANDLE_BYTES 64
int mbar_b200_peer_export(mbar_b200_ctx* ctx, void* handle_out /* [64] */);
/* (needs mbar_b200_comm_init first: every host-stepped path and the K x K Hessian reduce through the communicator;
 * at most 16 ranks) */
int mbar_b200_peer_attach(mbar_b200_ctx* ctx, int32_t nranks, int32_t rank, const void* handles /* [nranks][64] */);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* MBAR_B200_H */
