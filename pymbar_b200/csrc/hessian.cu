// K x K second-moment matrix of the weights, Ghat = (N W)^T (N W), on the fp64 tensor pipe.
//
// Reference being replaced: mbar_hessian (mbar_solvers.py:395-411): W = exp(f - u^T - L) [N,K];
// H = -( (W^T W) * N N^T - diag(N_k sum_n W_nk) ).  With  w_kn = N_k W_nk = exp(c_k - u'_kn - L'_n):
//   H_ij = delta_ij N_i S_i - Ghat_ij,   Ghat_ij = sum_n w_in w_jn.
// This is the only compute-bound piece of the path (2 K^2 N flop vs 8 K N bytes).  There is no fp64
// tcgen05 MMA, so it uses the warp-level DMMA (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4), which on B200
// runs on its own tensor sub-pipe at 64 MAC/clk/SM (measured: 16 cycles per instruction per scheduler)
// — the same rate as the vector fp64 pipe, but concurrently with it.
//
// Decomposition: lower block-triangle of Ghat in 128 x 128 blocks; a CTA owns one block pair
// (bi >= bj) and one contiguous chunk of tiles.  Per tile:
//   * thread 0 streams the two 128 x 32 energy panels (32 KB contiguous each in the tile-major
//     layout) into a 3-deep shared-memory ring with cp.async.bulk + mbarriers;
//   * phase 1: every warp turns 16 rows of each panel into weights IN PLACE (one exp per entry,
//     stored with an XOR swizzle so that the 8-row x 4-sample DMMA fragments are conflict free);
//   * one __syncthreads; phase 2: 8 warps x (64 x 32) register tiles, 32 DMMA per 4-sample step.
// Per-chunk partial blocks are reduced by a second kernel in chunk order (deterministic).
#include <cmath>

#include "internal.cuh"

namespace mbar {

constexpr int HB = 128;                    // block edge
constexpr int HNS = 3;                     // ring depth
constexpr uint32_t HPANEL = HB * TILE_N * 8;   // 32 KB

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Converts row `row` of a 128 x 32 energy panel into weights in place (swizzled store).
__device__ __forceinline__ void convert_row(double* P, int row, int lane, double ck, bool act, bool valid,
                                            double L, const double* tab, double sw) {
    const double v = P[row * TILE_N + lane];
    __syncwarp();   // every lane has read the row before any lane overwrites a permuted slot of it
    double wv = 0.0;
    if (valid && act) wv = sw * exp_fast(fmin(fmax(ck - v - L, -800.0), 700.0), tab);
    P[row * TILE_N + (lane ^ ((row & 7) << 2))] = wv;
}

// Work split: CTAs [pairStart[p], pairStart[p+1]) own block pair p and share its tiles evenly.  Diagonal
// pairs only multiply the 10 lower-triangular 32 x 32 sub-blocks (3 instead of 4 per scheduler), so
// they get 3/4 of the CTAs an off-diagonal pair gets.
struct HessSplit {
    int nPairs;
    int pairStart[140];
};

__global__ void __launch_bounds__(512, 1)
hessian_kernel(const double* __restrict__ u, const double* __restrict__ Lp,
               const double* __restrict__ c, const unsigned long long* __restrict__ rowmask, int K,
               int64_t N, int64_t nTiles, const HessSplit split, double* __restrict__ Gpart,
               const double* __restrict__ sqrtw) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tab = reinterpret_cast<double*>(smem_raw);                      // [32]
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(tab + 32);             // [HNS]
    uint64_t* bar_empty = bar_full + 4;                                     // [HNS]
    unsigned char* ring = smem_raw + 512;                                   // [HNS][2][HPANEL]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;             // 16 warps

    int pair = 0;
    while (pair + 1 < split.nPairs && (int)blockIdx.x >= split.pairStart[pair + 1]) ++pair;
    const int chunk = blockIdx.x - split.pairStart[pair];
    const int nChunks = split.pairStart[pair + 1] - split.pairStart[pair];
    int bi = 0, rem = pair;                       // pairs enumerated (0,0),(1,0),(1,1),(2,0),...
    while (rem > bi) { rem -= bi + 1; ++bi; }
    const int bj = rem;
    const bool diag = (bi == bj);
    const int64_t t0 = nTiles * chunk / nChunks, t1 = nTiles * (chunk + 1) / nChunks;
    const int rowsI = min(HB, K - bi * HB), rowsJ = min(HB, K - bj * HB);

    if (threadIdx.x < 32) tab[threadIdx.x] = MBAR_EXP_TABLE[threadIdx.x];
    if (threadIdx.x == 0) {
        for (int i = 0; i < HNS; ++i) {
            mbar_init(smem_u32(&bar_full[i]), 1);
            mbar_init(smem_u32(&bar_empty[i]), 16);
        }
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int it) {
        const int64_t tile = t0 + it;
        if (tile >= t1) return;
        const int slot = it % HNS;
        if (it >= HNS) mbar_wait(smem_u32(&bar_empty[slot]), ((it / HNS) - 1) & 1);
        const uint32_t fb = smem_u32(&bar_full[slot]);
        const uint32_t bytesI = (uint32_t)rowsI * TILE_N * 8, bytesJ = (uint32_t)rowsJ * TILE_N * 8;
        mbar_arrive_expect_tx(fb, bytesI + (diag ? 0u : bytesJ));
        const double* base = u + tile * (int64_t)K * TILE_N;
        const uint32_t dst = smem_u32(ring + (size_t)slot * 2 * HPANEL);
        bulk_g2s(dst, base + (int64_t)bi * HB * TILE_N, bytesI, fb);
        if (!diag) bulk_g2s(dst + HPANEL, base + (int64_t)bj * HB * TILE_N, bytesJ, fb);
    };
    if (threadIdx.x == 0) {
        issue(0);
        issue(1);
    }

    // conversion (vector fp64 pipe): warp owns rows warp*8 .. +7 of each panel, lane = sample.
    // MMA (tensor pipe): warp tile 32 (i) x 32 (j).  Off-diagonal pair: wm = warp / 4, wn = warp % 4.
    // Diagonal pair: only the 10 sub-blocks with wm >= wn are needed (Ghat is symmetric); warps 0..9 take
    // them, which loads the four schedulers 3/3/2/2 instead of 4/4/4/4, warps 10..15 only convert.
    int wm = warp >> 2, wn = warp & 3;
    bool mmaWarp = true;
    if (diag) {
        mmaWarp = warp < 10;
        int t = warp < 10 ? warp : 0;
        wm = 0;
        while (t > wm) { t -= wm + 1; ++wm; }     // triangular enumeration (0,0),(1,0),(1,1),(2,0),...
        wn = t;
    }
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    double cI[8], cJ[8];
    uint32_t actI = 0, actJ = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int ki = bi * HB + warp * 8 + r, kj = bj * HB + warp * 8 + r;
        const bool ai = ki < K && ((rowmask[ki >> 6] >> (ki & 63)) & 1ull);
        const bool aj = kj < K && ((rowmask[kj >> 6] >> (kj & 63)) & 1ull);
        cI[r] = ai ? c[ki] : 0.0;
        cJ[r] = aj ? c[kj] : 0.0;
        actI |= (uint32_t)ai << r;
        actJ |= (uint32_t)aj << r;
    }
    const int fragCol = lane & 3, fragRow = lane >> 2;   // DMMA fragment coordinates of this lane
    const int nIter = (int)(t1 - t0);

    // prologue: weights of the first tile
    if (nIter > 0) {
        const double L = Lp[t0 * TILE_N + lane];
        const double sw0 = sqrtw ? sqrtw[t0 * TILE_N + lane] : 1.0;   // sqrt of the bootstrap multiplicity
        const bool valid = t0 * TILE_N + lane < N;
        mbar_wait(smem_u32(&bar_full[0]), 0);
        double* Pi = reinterpret_cast<double*>(ring);
        double* Pj = Pi + HB * TILE_N;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            convert_row(Pi, warp * 8 + r, lane, cI[r], (actI >> r) & 1u, valid, L, tab, sw0);
            if (!diag) convert_row(Pj, warp * 8 + r, lane, cJ[r], (actJ >> r) & 1u, valid, L, tab, sw0);
        }
    }
    __syncthreads();

    for (int it = 0; it < nIter; ++it) {
        const int slot = it % HNS;
        const int64_t tile = t0 + it;
        if (threadIdx.x == 0) issue(it + 2);          // slot (it+2)%3 was released at the end of it-1
        // the NEXT tile's energies are converted while this tile is multiplied (separate pipes)
        const bool haveNext = it + 1 < nIter;
        const int nslot = (it + 1) % HNS;
        double Ln = 0.0, swn = 1.0;
        bool validN = false;
        if (haveNext) {
            Ln = Lp[(tile + 1) * TILE_N + lane];
            if (sqrtw) swn = sqrtw[(tile + 1) * TILE_N + lane];
            validN = (tile + 1) * TILE_N + lane < N;
            mbar_wait(smem_u32(&bar_full[nslot]), ((it + 1) / HNS) & 1);
        }
        double* Ni = reinterpret_cast<double*>(ring + (size_t)nslot * 2 * HPANEL);
        double* Nj = Ni + HB * TILE_N;
        const double* Pi = reinterpret_cast<const double*>(ring + (size_t)slot * 2 * HPANEL);
        const double* Pj = diag ? Pi : Pi + HB * TILE_N;
        const double* Ai = Pi + (wm * 32 + fragRow) * TILE_N;
        const double* Bj = Pj + (wn * 32 + fragRow) * TILE_N;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            if (haveNext) {
                convert_row(Ni, warp * 8 + ks, lane, cI[ks], (actI >> ks) & 1u, validN, Ln, tab, swn);
                if (!diag) convert_row(Nj, warp * 8 + ks, lane, cJ[ks], (actJ >> ks) & 1u, validN, Ln, tab, swn);
            }
            if (!mmaWarp) continue;
            const int col = ((ks ^ fragRow) << 2) + fragCol;      // (4 ks + fragCol) ^ (fragRow << 2)
            double a[4], b[4];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) a[mt] = Ai[mt * 8 * TILE_N + col];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) b[nt] = Bj[nt * 8 * TILE_N + col];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
        }
        fence_proxy_async_smem();   // slot `slot` was written in place; its next writer is the TMA engine
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_empty[slot]));
        __syncthreads();            // next tile's weights are complete for every warp
    }
    // write this CTA's 128 x 128 partial block: Gpart[chunk][pair][128][128]
    double* out = Gpart + (size_t)blockIdx.x * HB * HB;
    if (mmaWarp)
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int row = wm * 32 + mt * 8 + (lane >> 2);
            const int col = wn * 32 + nt * 8 + (lane & 3) * 2;
            out[row * HB + col] = acc[mt][nt][0];
            out[row * HB + col + 1] = acc[mt][nt][1];
        }
}

// Sum the partial blocks of each pair over its CTAs (in order: deterministic) and scatter to the full
// symmetric K x K matrix.  Diagonal pairs only hold sub-blocks with row-block >= column-block.
__global__ void __launch_bounds__(256)
hessian_reduce_kernel(const double* __restrict__ Gpart, int K, const HessSplit split, double* __restrict__ G) {
    const int pair = blockIdx.x;
    int bi = 0, rem = pair;
    while (rem > bi) { rem -= bi + 1; ++bi; }
    const int bj = rem;
    const int c0 = split.pairStart[pair], c1 = split.pairStart[pair + 1];
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < HB * HB; e += gridDim.y * blockDim.x) {
        const int r = e / HB, cidx = e % HB;
        const int i = bi * HB + r, j = bj * HB + cidx;
        if (i >= K || j >= K) continue;
        if (bi == bj && (r >> 5) < (cidx >> 5)) continue;      // filled by its mirror image below
        double s = 0.0;
        for (int ch = c0; ch < c1; ++ch) s += Gpart[(size_t)ch * HB * HB + e];
        G[(size_t)i * K + j] = s;
        if (bi != bj || (r >> 5) > (cidx >> 5)) G[(size_t)j * K + i] = s;
    }
}

// Requires ctx->d_L (shifted-frame L'_n) from the preceding pass at the same f.
int launch_hessian(mbar_b200_ctx* ctx, const double* h_f, bool allRows) {
    const int K = ctx->K;
    MBAR_REQUIRE(ctx->d_L, MBAR_B200_ERR_NOT_READY, "hessian: per-sample L not available");
    // sampled rows carry N_k W_nk (c = f + log N); with allRows the unsampled rows carry W_nk (c = f)
    for (int k = 0; k < K; ++k)
        ctx->h_f[2 * K + k] = std::isinf(ctx->h_logNk[k]) ? (allRows ? h_f[k] : 0.0) : h_f[k] + ctx->h_logNk[k];
    MBAR_CUDA(cudaMemcpyAsync(ctx->d_c + 2 * K, ctx->h_f + 2 * K, (size_t)K * sizeof(double),
                              cudaMemcpyHostToDevice, ctx->stream));
    const int nB = (K + HB - 1) / HB;
    const int nPairs = nB * (nB + 1) / 2;
    MBAR_REQUIRE(nPairs < 140, MBAR_B200_ERR_INVALID, "K=%d too large for the Hessian kernel (K <= 2048)", K);
    // CTAs per pair proportional to its cost per tile (4 for off-diagonal, 3 for diagonal pairs)
    HessSplit split{};
    split.nPairs = nPairs;
    {
        double wsum = 0.0;
        for (int bi = 0, p = 0; bi < nB; ++bi)
            for (int bj = 0; bj <= bi; ++bj, ++p) wsum += (bi == bj) ? 3.0 : 4.0;
        int total = ctx->smCount > nPairs ? ctx->smCount : nPairs;
        int used = 0;
        for (int bi = 0, p = 0; bi < nB; ++bi)
            for (int bj = 0; bj <= bi; ++bj, ++p) {
                int n = (int)(total * ((bi == bj) ? 3.0 : 4.0) / wsum);
                if (n < 1) n = 1;
                if ((int64_t)n > ctx->nTiles) n = (int)ctx->nTiles;
                split.pairStart[p] = used;
                used += n;
            }
        split.pairStart[nPairs] = used;
    }
    const int nCtas = split.pairStart[nPairs];
    const size_t partBytes = (size_t)nCtas * HB * HB * sizeof(double);
    if (!ctx->d_W) {
        // sized for the larger of the two row selections (identical split for both)
        MBAR_CUDA(cudaMalloc((void**)&ctx->d_W, partBytes));
        MBAR_CUDA(cudaMemset(ctx->d_W, 0, partBytes));
    }
    const size_t smem = 512 + (size_t)HNS * 2 * HPANEL;
    static bool attr[16] = {false};
    if (!attr[ctx->device & 15]) {
        MBAR_CUDA(cudaFuncSetAttribute(hessian_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr[ctx->device & 15] = true;
    }
    const PassLayout lay{K};
    hessian_kernel<<<nCtas, 512, smem, ctx->stream>>>(
        ctx->d_u, ctx->d_L, ctx->d_c + 2 * K, allRows ? ctx->d_onesmask : ctx->d_rowmask, K, ctx->N, ctx->nTiles,
        split, ctx->d_W, ctx->d_sqrtw);
    MBAR_CUDA(cudaGetLastError());
    hessian_reduce_kernel<<<dim3(nPairs, 16), 256, 0, ctx->stream>>>(ctx->d_W, K, split, ctx->d_out + lay.G());
    MBAR_CUDA(cudaGetLastError());
    ctx->launches += 2;
    ctx->passes++;
    return MBAR_B200_OK;
}

}  // namespace mbar
