// K x K second-moment matrix of the weights, Ghat = (N W)^T (N W), on the fp64 tensor pipe.
//
// Reference being replaced: mbar_hessian (mbar_solvers.py:395-411): W = exp(f - u^T - L) [N,K];
// H = -( (W^T W) * N N^T - diag(N_k sum_n W_nk) ).  With  w_kn = N_k W_nk = exp(c_k - u'_kn - L'_n):
//   H_ij = delta_ij N_i S_i - Ghat_ij,   Ghat_ij = sum_n w_in w_jn.
// This is the only compute-bound piece of the path (2 K^2 N flop vs 8 K N bytes); there is no fp64
// tcgen05 MMA, so it uses the legacy warp-level DMMA (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4).
//
// Decomposition: the lower block-triangle of Ghat in 128 x 128 blocks; a CTA owns one block pair
// (bi >= bj) and one contiguous chunk of tiles, recomputes the two 128 x 32 weight panels per tile
// from u' and the stored L'_n (one exp per entry), stages them in shared memory (stride 36: the
// 8-row x 4-sample DMMA fragments are bank-conflict free) and accumulates 64 x 32 per warp in
// registers.  Per-chunk partial blocks are reduced by a second kernel in chunk order (deterministic).
#include <cmath>

#include "internal.cuh"

namespace mbar {

constexpr int HB = 128;       // block edge
constexpr int HSTRIDE = 36;   // smem row stride in doubles (32 samples + 4 pad)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 1)
hessian_kernel(const double* __restrict__ u, const double* __restrict__ Lp,
               const double* __restrict__ c, const unsigned long long* __restrict__ rowmask, int K,
               int64_t N, int64_t nTiles, int nChunks, double* __restrict__ Gpart) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tab = reinterpret_cast<double*>(smem_raw);           // [32]
    double* panels = tab + 32;                                   // [2 buf][2 panel][HB][HSTRIDE]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 32) tab[threadIdx.x] = MBAR_EXP_TABLE[threadIdx.x];

    // block pair from blockIdx.x: pairs enumerated (0,0),(1,0),(1,1),(2,0),...
    int bi = 0, rem = blockIdx.x;
    while (rem > bi) { rem -= bi + 1; ++bi; }
    const int bj = rem;
    const bool diag = (bi == bj);
    const int chunk = blockIdx.y;
    const int64_t t0 = nTiles * chunk / nChunks, t1 = nTiles * (chunk + 1) / nChunks;

    // phase-1 mapping: lane = sample, warp fills rows warp*16 .. +15 of each panel
    // phase-2 mapping: warp tile 64 (i) x 32 (j): wm = warp / 4 (0..1), wn = warp % 4 (0..3)
    const int wm = warp >> 2, wn = warp & 3;
    double acc[8][4][2];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    double cI[16], cJ[16];
    uint32_t actI = 0, actJ = 0;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int ki = bi * HB + warp * 16 + r, kj = bj * HB + warp * 16 + r;
        const bool ai = ki < K && ((rowmask[ki >> 6] >> (ki & 63)) & 1ull);
        const bool aj = kj < K && ((rowmask[kj >> 6] >> (kj & 63)) & 1ull);
        cI[r] = ai ? c[ki] : 0.0;
        cJ[r] = aj ? c[kj] : 0.0;
        actI |= (uint32_t)ai << r;
        actJ |= (uint32_t)aj << r;
    }
    __syncthreads();

    int buf = 0;
    for (int64_t tile = t0; tile < t1; ++tile, buf ^= 1) {
        double* Pi = panels + (size_t)buf * 2 * HB * HSTRIDE;
        double* Pj = diag ? Pi : Pi + HB * HSTRIDE;
        const bool valid = tile * TILE_N + lane < N;
        const double L = Lp[tile * TILE_N + lane];
        const double* tp = u + tile * (int64_t)K * TILE_N + lane;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int kl = warp * 16 + r;
            double wv = 0.0;
            if (valid && ((actI >> r) & 1u))
                wv = exp_fast(fmax(cI[r] - tp[(int64_t)(bi * HB + kl) * TILE_N] - L, -800.0), tab);
            Pi[kl * HSTRIDE + lane] = wv;
        }
        if (!diag) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int kl = warp * 16 + r;
                double wv = 0.0;
                if (valid && ((actJ >> r) & 1u))
                    wv = exp_fast(fmax(cJ[r] - tp[(int64_t)(bj * HB + kl) * TILE_N] - L, -800.0), tab);
                Pj[kl * HSTRIDE + lane] = wv;
            }
        }
        __syncthreads();
        const double* Ai = Pi + (wm * 64 + (lane >> 2)) * HSTRIDE + (lane & 3);
        const double* Bj = Pj + (wn * 32 + (lane >> 2)) * HSTRIDE + (lane & 3);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            double a[8], b[4];
#pragma unroll
            for (int mt = 0; mt < 8; ++mt) a[mt] = Ai[mt * 8 * HSTRIDE + ks * 4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) b[nt] = Bj[nt * 8 * HSTRIDE + ks * 4];
#pragma unroll
            for (int mt = 0; mt < 8; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
        }
        // no second barrier: the next tile writes the other buffer
    }
    // write this CTA's 128 x 128 partial block: Gpart[chunk][pair][128][128]
    const int nPairs = gridDim.x;
    double* out = Gpart + ((size_t)chunk * nPairs + blockIdx.x) * HB * HB;
#pragma unroll
    for (int mt = 0; mt < 8; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int row = wm * 64 + mt * 8 + (lane >> 2);
            const int col = wn * 32 + nt * 8 + (lane & 3) * 2;
            out[row * HB + col] = acc[mt][nt][0];
            out[row * HB + col + 1] = acc[mt][nt][1];
        }
}

// Sum partial blocks over chunks (in order) and scatter to the full symmetric K x K matrix.
__global__ void __launch_bounds__(256)
hessian_reduce_kernel(const double* __restrict__ Gpart, int K, int nPairs, int nChunks,
                      double* __restrict__ G) {
    const int pair = blockIdx.x;
    int bi = 0, rem = pair;
    while (rem > bi) { rem -= bi + 1; ++bi; }
    const int bj = rem;
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < HB * HB; e += gridDim.y * blockDim.x) {
        const int r = e / HB, cidx = e % HB;
        const int i = bi * HB + r, j = bj * HB + cidx;
        if (i >= K || j >= K) continue;
        double s = 0.0;
        for (int ch = 0; ch < nChunks; ++ch) s += Gpart[((size_t)ch * nPairs + pair) * HB * HB + e];
        G[(size_t)i * K + j] = s;
        if (bi != bj) G[(size_t)j * K + i] = s;
    }
}

// Requires ctx->d_L (shifted-frame L'_n) from the preceding pass at the same f.
int launch_hessian(mbar_b200_ctx* ctx, const double* h_f) {
    const int K = ctx->K;
    MBAR_REQUIRE(ctx->d_L, MBAR_B200_ERR_NOT_READY, "hessian: per-sample L not available");
    for (int k = 0; k < K; ++k)
        ctx->h_f[2 * K + k] = std::isinf(ctx->h_logNk[k]) ? 0.0 : h_f[k] + ctx->h_logNk[k];
    MBAR_CUDA(cudaMemcpyAsync(ctx->d_c + 2 * K, ctx->h_f + 2 * K, (size_t)K * sizeof(double),
                              cudaMemcpyHostToDevice, ctx->stream));
    const int nB = (K + HB - 1) / HB;
    const int nPairs = nB * (nB + 1) / 2;
    int nChunks = ctx->smCount / nPairs;
    if (nChunks < 1) nChunks = 1;
    if ((int64_t)nChunks > ctx->nTiles) nChunks = (int)ctx->nTiles;
    const size_t partBytes = (size_t)nChunks * nPairs * HB * HB * sizeof(double);
    if (!ctx->d_W) MBAR_CUDA(cudaMalloc((void**)&ctx->d_W, partBytes));
    const size_t smem = 256 + (size_t)2 * 2 * HB * HSTRIDE * sizeof(double);
    MBAR_CUDA(cudaFuncSetAttribute(hessian_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const PassLayout lay{K};
    hessian_kernel<<<dim3(nPairs, nChunks), 256, smem, ctx->stream>>>(
        ctx->d_u, ctx->d_L, ctx->d_c + 2 * K, ctx->d_rowmask, K, ctx->N, ctx->nTiles, nChunks, ctx->d_W);
    MBAR_CUDA(cudaGetLastError());
    hessian_reduce_kernel<<<dim3(nPairs, 16), 256, 0, ctx->stream>>>(ctx->d_W, K, nPairs, nChunks,
                                                                    ctx->d_out + lay.G());
    MBAR_CUDA(cudaGetLastError());
    ctx->launches += 2;
    ctx->passes++;
    return MBAR_B200_OK;
}

}  // namespace mbar
