// Shared internals of libmbar_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <functional>
#include <vector>

#include "../../include/mbar_b200.h"
#include "exp_tables.h"

namespace mbar {

constexpr int TILE_N = 32;              // samples per tile (one warp lane each)
constexpr double U_CLAMP = 1.0e6;       // shifted energies are clamped to [.., 1e6] at upload
constexpr double C_RANGE = 1.0e6;       // |f_k + log N_k| must stay below this
constexpr double FUSED_SPREAD = 1200.0; // fused kernel needs max(c) - min(c) below this
constexpr int MAX_GRID = 148 * 4;
// Unsampled states ride through the fused kernel as "sampled with weight e^-80": their share of any
// denominator is below 2^-53 as long as their sum of weights stays below 1e12 (checked by the host).
constexpr double LOG_EPS_UNSAMPLED = -80.0;

void set_error(const char* fmt, ...);
#define MBAR_CUDA(call)                                                                     \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            mbar::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,            \
                            cudaGetErrorString(e__));                                       \
            return MBAR_B200_ERR_CUDA;                                                      \
        }                                                                                   \
    } while (0)
#define MBAR_REQUIRE(cond, status, ...)                                                     \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            mbar::set_error(__VA_ARGS__);                                                   \
            return (status);                                                                \
        }                                                                                   \
    } while (0)
#define MBAR_TRY(call)                                                                      \
    do {                                                                                    \
        int s__ = (call);                                                                   \
        if (s__ != MBAR_B200_OK) return s__;                                                \
    } while (0)

// Packed per-pass output on device / pinned host:  [0..K) S_k, [K] sumL, [K+1] flag (as double),
// [K+2 .. K+2+K) log-domain S for unsampled states (generic kernel), then K*K G when requested.
struct PassLayout {
    int K;
    __host__ __device__ int S() const { return 0; }
    __host__ __device__ int sumL() const { return K; }
    __host__ __device__ int flag() const { return K + 1; }
    __host__ __device__ int logS() const { return K + 2; }
    __host__ __device__ int G() const { return 2 * K + 2; }
    __host__ __device__ int size(bool withG) const { return 2 * K + 2 + (withG ? K * K : 0); }
};

// Peer-memory exchange of the per-pass partial sums (fused into the pass kernel's last CTA).
constexpr int MAX_PEERS = 16;
struct PeerCfg {
    int nranks = 1, rank = 0;
    double* inbox[MAX_PEERS] = {nullptr};               // inbox[q]: rank q's [2][nranks][K+2] buffer, mapped here
    unsigned long long* flags[MAX_PEERS] = {nullptr};   // flags[q]: rank q's [2][nranks] sequence numbers
    // Exchange sequence number, kept on the DEVICE and advanced by the last CTA of every launch that
    // actually runs: launches that exit early (solver already converged) must not consume a number, or two
    // executed launches could reuse an inbox parity without an exchange in between.
    unsigned long long* seq = nullptr;
};

// State of a device-resident solver loop (device memory + pinned host mirror).  Every kernel of an iteration
// starts with `if (loop->done) return;`, so the host can enqueue a batch of iterations without knowing when
// the solver converges and polls this struct once per batch: no per-iteration host round trip.
struct LoopState {
    int done;            // 1: stop (converged, failed or maxiter reached)
    int status;          // 0 ok | 1 range/underflow problem: redo on the robust host-stepped path | 2 comm
    int success;         // convergence criterion met (mbar_solvers.py:627-640)
    int iterations, nr_iterations, sci_iterations;
    int maxiter, min_sc_iter;
    int haveNr;          // this iteration has a valid Newton candidate
    int cholFail;        // last Cholesky attempt met a non-positive pivot
    double tol, gamma;
    double max_delta, max_diff, gnorm;
    double gn_sci, gn_nr;
};

}  // namespace mbar

struct mbar_b200_ctx {
    int device = 0;
    int K = 0;
    int64_t N = 0;        // local samples
    int64_t nTiles = 0;   // ceil(N / 32)
    bool ready = false;   // u uploaded
    int kernelChoice = MBAR_B200_KERNEL_AUTO;
    int smCount = 148;

    std::vector<double> h_Nk;       // [K]
    std::vector<double> h_logNk;    // [K], -inf for unsampled
    std::vector<double> h_logNkEff; // [K], LOG_EPS_UNSAMPLED for unsampled
    double* d_NkEff = nullptr;      // [K], exp(LOG_EPS_UNSAMPLED) for unsampled
    bool unsampledExtreme = false;  // an unsampled state lies > 1e5 kT below every sampled one somewhere
    std::vector<int> active;        // indices of sampled states
    int firstActive = 0;
    double N_total_states = 0;      // sum_k N_k (global N)

    double* d_u = nullptr;          // [nTiles][K][32] shifted, clamped
    size_t uBytes = 0;              // bytes this context allocated for d_u (0: came from the pool)
    double* d_xshift = nullptr;     // [nTiles*32] per-sample shift x_n = min over sampled k of u_kn
    double* d_wgt = nullptr;        // [nTiles*32] per-sample multiplicities w_n (bootstrap), or NULL = all 1
    double* d_sqrtw = nullptr;      // [nTiles*32] sqrt(w_n) for the second-moment kernel
    double sumW = 0.0;              // sum_n w_n over valid local samples (= N when unweighted)
    double sumXw = 0.0;             // sum_n w_n x_n
    double sumX = 0.0;              // sum_n x_n over valid local samples
    double* d_c = nullptr;          // [2][K] c_k = f_k + log N_k - mid (fused) | f_k (row K..2K)
    double* d_Nk = nullptr;         // [K]
    unsigned long long* d_rowmask = nullptr;  // [ceil(K/64)] bit per sampled state
    unsigned long long* d_zeromask = nullptr; // same size, all zero (log-domain for every row)
    unsigned long long* d_onesmask = nullptr; // same size, all one (second moments of every state)
    double* d_partial = nullptr;    // [MAX_GRID][K+2] per-CTA partials
    double* d_out = nullptr;        // PassLayout packed result (with G)
    double* h_out = nullptr;        // pinned mirror of d_out
    double* d_L = nullptr;          // [nTiles*32] per-sample L_n (lazy)
    double* d_W = nullptr;          // per-CTA partial blocks of the Hessian kernels (gpartBytes)
    unsigned int* d_ticket = nullptr;
    int* d_flag = nullptr;          // [4] error/diagnostic flags
    double* d_f = nullptr;          // [4][K] device-resident f vectors for native loops
    double* h_f = nullptr;          // pinned [4][K]
    double* d_scratch = nullptr;    // misc K*K scratch
    cudaStream_t stream = nullptr;
    cudaStream_t copyStream = nullptr;
    cudaEvent_t evA = nullptr, evB = nullptr;
    cudaEvent_t evCopy[2] = {nullptr, nullptr};
    double* stage_pinned[2] = {nullptr, nullptr};
    double* stage_dev[2] = {nullptr, nullptr};
    int64_t stageCols = 0;

    // peer-memory exchange (cudaIpc): this rank's inbox + flags and the peers' mappings
    double* d_inbox = nullptr;
    mbar::PeerCfg peer;
    bool peerReady = false;
    std::vector<void*> peerMapped;

    // communicator (NCCL, dlopen'd)
    void* comm = nullptr;
    int nranks = 1, rank = 0;

    // device-resident solver loops
    mbar::LoopState* d_loop = nullptr;   // device
    mbar::LoopState* h_loop = nullptr;   // pinned mirror
    double* d_av = nullptr;              // [8][K] adaptive work vectors (f_sci, f_nr, g, c_sci, c_nr, c_hess, x, -)
    double* d_outM = nullptr;            // [2][2K+2] pass outputs of the two candidates
    double* d_A = nullptr;               // [K*K] Newton matrix / Cholesky factor
    int* d_active = nullptr;             // [K] indices of the sampled states
    unsigned long long* d_seq = nullptr; // peer-exchange sequence number (device-side, see PeerCfg::seq)
    int loopMode = 0;                    // 0 device-resident, 1 host-stepped (round-1 behaviour)
    int loopBatch = 4;                   // iterations enqueued between two polls of LoopState
    int64_t loopPolls = 0;               // host synchronisations spent polling LoopState
    // the adaptive iteration as a CUDA graph (captured once, relaunched per iteration; see loops.cu)
    bool capturing = false;              // stream capture in progress: no timing events, no allocations
    bool graphWarm = false;              // one uncaptured iteration has sized every buffer / kernel attribute
    cudaGraphExec_t loopGraph = nullptr;
    std::vector<uint64_t> loopGraphKey;
    int64_t graphLaunches = 0, graphCaptures = 0;
    double* d_Wt = nullptr;              // [nTiles][K][32] materialised N_k W_nk (swizzled) for the Hessian
    bool wtAllocFailed = false;
    size_t gpartBytes = 0;               // size of d_W (per-CTA partial blocks of the Hessian kernels)
    char lastKernel[200] = "";           // description of the pass-kernel variant launched last
    char lastHessKernel[200] = "";       // ... and of the Hessian kernel path
    double lastHessMs = 0.0, lastWeightsMs = 0.0;
    cudaEvent_t evH0 = nullptr, evH1 = nullptr, evH2 = nullptr;

    // counters
    int64_t launches = 0, passes = 0, h2dBytes = 0, d2hBytes = 0;
    double lastPassMs = 0.0;
    double lastLoopMs = 0.0, lastLoopKernelMs = 0.0;
    int lastLoopIters = 0;
    bool timePasses = true;
};

namespace mbar {

struct FusedParams {
    const double* u;
    const double* c;                       // [K] f_k + log N_k - mid (sampled rows)
    const double* c2;                      // second candidate (M == 2): [K] f2_k + log N_k - mid2
    double* out2;                          // ... and its packed result
    double mid2;
    int M;                                 // candidates evaluated per launch (1 or 2)
    const unsigned long long* rowmask;
    const double* Nk;
    double* partial;                       // [grid][K + 2]
    double* out;
    unsigned int* ticket;
    double* Lout;                          // [nTiles*32] shifted-frame L'_n, or NULL
    double* Wout;                          // [nTiles][K][32] N_k W_nk (swizzled rows) for the Hessian kernels, or NULL
    const double* wgt;                     // [nTiles*32] sample multiplicities or NULL
    double sumW;                           // sum of the multiplicities of this shard (N if unweighted)
    double* f;                             // [K] device f_k (epilogue) or NULL
    double* cnext;                         // [K] where the epilogue writes c for the next launch
    PeerCfg peer;
    LoopState* loop;                       // device-resident loop state (early exit + convergence) or NULL
    int epi, first;
    int64_t N, nTiles, nStages;
    double mid;
    int K, Wk, Wn, Rw, TPW, NS, CW, batch, debugSkip, mode, CL, Kh, allStates;
    uint32_t tileBytes, stageBytes;
};

// What a host-driven pass should leave behind (api.cu: run_pass).
struct PassWant {
    bool L = false;          // keep per-sample L'_n on the device
    bool unsampled = false;  // need log-domain sums for N_k == 0 states
    bool G = false;          // K x K second moments
    bool Gall = false;       // ... including the unsampled states' rows and columns
};

// NVTX ranges around pass / exchange / Hessian / upload / solver loops.  nvtx3 is header
// only: without an attached tool every call is a no-op through a NULL function table.
struct NvtxRange {
    explicit NvtxRange(const char* name);
    ~NvtxRange();
};

// Persistent host worker threads for the memcpy-bound staging steps (packing pageable uploads, draining log W):
// spawning threads per 64 MB chunk cost as much as a quarter of the chunk's copy time.  run(n, fn) executes
// fn(0..n-1) on the pool plus the calling thread and returns when all are done.
void host_parallel(int nTasks, const std::function<void(int)>& fn);
int host_parallel_width();

// Prefer the GPU's NUMA node for host allocations made while this object lives (ctx.cu).
int gpu_numa_node(int device);
struct NumaPrefer {
    bool active = false;
    explicit NumaPrefer(int device);
    ~NumaPrefer();
};

// ---- host-side helpers implemented across the .cu files ----
int check_range(mbar_b200_ctx* c, const double* f);
int run_pass(mbar_b200_ctx* c, const double* f, PassWant want);   // pass + all-reduce + D2H into ctx->h_out
double global_sumx(mbar_b200_ctx* c, int* rc);
int solve_sci_stepped(mbar_b200_ctx* c, double* f, double tol, int32_t maxiter, mbar_b200_solve_result* res);
int solve_adaptive_stepped(mbar_b200_ctx* c, double* f, double tol, int32_t maxiter, int32_t min_sc_iter,
                           double gamma, mbar_b200_solve_result* res);
int solve_sci_device(mbar_b200_ctx* c, double* f, double tol, int32_t maxiter, mbar_b200_solve_result* res);
int solve_adaptive_device(mbar_b200_ctx* c, double* f, double tol, int32_t maxiter, int32_t min_sc_iter,
                          double gamma, mbar_b200_solve_result* res);
// stream-ordered rendezvous of all ranks (one tiny all-reduce) before the first in-kernel peer exchange of a loop
int comm_rendezvous(mbar_b200_ctx* ctx);
int retile_chunk(mbar_b200_ctx* ctx, const double* d_rowmajor, int64_t ldCols, int64_t tile0,
                 int64_t nTilesChunk, int64_t validCols, cudaStream_t s);
int launch_pass_generic(mbar_b200_ctx* ctx, const double* h_f, bool wantL, bool logAll);
int launch_pass_fused(mbar_b200_ctx* ctx, const double* h_f, bool wantL, bool allStates, bool* usedOut,
                      bool wantW = false, bool* wroteW = nullptr);
// d_cdst / h_stage: where c = f + log N - mid is staged (default: ctx->d_c / ctx->h_f); midForce: reuse the
// centring of a previous prepare (candidates evaluated against the same exp(c) range), NaN = derive from f
int fused_prepare(mbar_b200_ctx* ctx, const double* h_f, bool wantL, bool allStates, FusedParams* out, bool* ok,
                  double* d_cdst = nullptr, double* h_stage = nullptr, bool wantW = false, int M = 1,
                  double midQuantum = 0.0);
int fused_enqueue(mbar_b200_ctx* ctx, const FusedParams& p);
bool fused_applicable(const mbar_b200_ctx* ctx, const double* h_f, bool allStates, double* midOut, double* spreadOut);
// weightsReady: the fused pass at this f already wrote N_k W_nk into ctx->d_Wt (FusedParams::Wout)
int launch_hessian(mbar_b200_ctx* ctx, const double* h_f, bool allRows, bool weightsReady = false);
// same with c_k = f_k + log N_k already on the device (device-resident loops); loop may be NULL
int launch_hessian_dev(mbar_b200_ctx* ctx, const double* d_ch, bool allRows, LoopState* loop,
                       bool weightsReady = false);
// the 8*K*N weight buffer of the K > 64 Hessian path; false when it cannot be allocated (in-place fallback)
bool ensure_weight_buffer(mbar_b200_ctx* ctx);
int launch_logw(mbar_b200_ctx* ctx, const double* h_f, double* logW_host, int64_t ld, int expo, int64_t n0,
                int64_t n);
int launch_synth(mbar_b200_ctx* ctx, const mbar_b200_synth* spec);
int launch_untile(mbar_b200_ctx* ctx, int64_t n0, int64_t n, double* d_dst, int64_t ld);
int comm_allreduce(mbar_b200_ctx* ctx, double* d_buf, int count, int op /*0 sum, 2 max*/);
int reduce_sumx(mbar_b200_ctx* ctx);
int set_weights(mbar_b200_ctx* ctx, const double* w_host);

// ---- device helpers ----
#ifdef __CUDACC__
static __device__ const double MBAR_EXP_TABLE[MBAR_EXP_NT] = {MBAR_EXP_TABLE_VALUES};
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// Same wait for the single producer lane: back off between polls so the spin does not steal issue
// slots from the consumer warps that share its scheduler.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(0x989680u)
            : "memory");
        if (done) break;
        __nanosleep(200);
    }
}
// 1-D bulk async copy global -> shared (TMA engine; SASS UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
// order this thread's generic-proxy shared-memory accesses before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

constexpr double EXP_MAGIC = 6755399441055744.0;  // 1.5 * 2^52

// exp(a) with the binary exponent kept apart:  exp(a) = v * 2^q,  v in [1, 2).
// a in [-5e7, 5e7]; `tab` = 32-entry table of 2^(j/32) in shared memory (conflict-free: 256 B).
// 9 fp64 pipe ops (FMA t, ADD nf, FMA r, 4 FMA Horner, MUL, FMA) + integer work on the ALU pipe.
__device__ __forceinline__ void exp_split(double a, const double* __restrict__ tab, double& v, int& q) {
    const double t = fma(a, MBAR_EXP_SCALE, EXP_MAGIC);
    const int n = __double2loint(t);
    const double nf = t - EXP_MAGIC;
    const double r = fma(nf, -MBAR_EXP_LN2N, a);
    double p = fma(MBAR_EXP_C5, r, MBAR_EXP_C4);
    p = fma(p, r, MBAR_EXP_C3);
    p = fma(p, r, MBAR_EXP_C2);
    p = fma(p, r, MBAR_EXP_C1);
    p = p * r;
    const double T = tab[n & (MBAR_EXP_NT - 1)];
    v = fma(T, p, T);
    q = n >> 5;
}
// v * 2^q with q clamped to the normal range from below (result >= ~2^-1021, never denormal) and
// assumed < 1023 from above.
__device__ __forceinline__ double scale2(double v, int q) {
    q = max(q, -1021);
    const int hi = __double2hiint(v) + (q << 20);
    return __hiloint2double(hi, __double2loint(v));
}
__device__ __forceinline__ double exp_fast(double a, const double* __restrict__ tab) {
    double v;
    int q;
    exp_split(a, tab, v, q);
    return scale2(v, q);
}
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
#endif

}  // namespace mbar
