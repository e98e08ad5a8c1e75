// mbar_log_W_nk / mbar_W_nk (mbar_solvers.py:439-507): logW[n, k] = f_k - u_kn - L_n, [N, K] row-major.
// Reads the tile-major u' once (+ the stored L'_n) and writes the transposed [N, K] layout through a
// 32 x 33 shared-memory transposition buffer per warp, so both the read and the write are coalesced.
#include <cmath>

#include "internal.cuh"

namespace mbar {

__global__ void __launch_bounds__(128)
logw_kernel(const double* __restrict__ u, const double* __restrict__ Lp, const double* __restrict__ f,
            int K, int64_t N, int64_t tile0, double* __restrict__ out, int expo) {
    __shared__ double T[4][32 * 33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile = tile0 + blockIdx.x;
    const double* tp = u + tile * (int64_t)K * TILE_N + lane;
    const double L = Lp[tile * TILE_N + lane];
    double* t = T[warp];
    for (int k0 = warp * 32; k0 < K; k0 += 4 * 32) {
        const int kmax = min(32, K - k0);
        for (int kk = 0; kk < kmax; ++kk) {
            double v = f[k0 + kk] - tp[(int64_t)(k0 + kk) * TILE_N] - L;
            if (expo) v = exp(v);
            t[kk * 33 + lane] = v;
        }
        __syncwarp();
        if (lane < kmax) {
            for (int n = 0; n < 32; ++n) {
                const int64_t row = (int64_t)blockIdx.x * TILE_N + n;       // row inside this chunk
                if (tile * TILE_N + n < N) out[row * K + k0 + lane] = t[lane * 33 + n];
            }
        }
        __syncwarp();
    }
}

// Requires ctx->d_L from a pass at the same f.  Streams [N, K] to the host in chunks.
int launch_logw(mbar_b200_ctx* ctx, const double* h_f, double* logW_host, int64_t ld, int expo) {
    const int K = ctx->K;
    MBAR_REQUIRE(ctx->d_L, MBAR_B200_ERR_NOT_READY, "log_W: per-sample L not available");
    for (int k = 0; k < K; ++k) ctx->h_f[3 * K + k] = h_f[k];
    MBAR_CUDA(cudaMemcpyAsync(ctx->d_c + 3 * K, ctx->h_f + 3 * K, (size_t)K * sizeof(double),
                              cudaMemcpyHostToDevice, ctx->stream));
    int64_t tilesPerChunk = (64ll << 20) / ((int64_t)K * TILE_N * 8);
    if (tilesPerChunk < 1) tilesPerChunk = 1;
    if (tilesPerChunk > ctx->nTiles) tilesPerChunk = ctx->nTiles;
    double* d_out[2] = {nullptr, nullptr};
    const size_t chunkBytes = (size_t)tilesPerChunk * TILE_N * K * sizeof(double);
    for (int i = 0; i < 2; ++i) MBAR_CUDA(cudaMalloc((void**)&d_out[i], chunkBytes));
    int rc = MBAR_B200_OK;
    cudaEvent_t done[2];
    for (int i = 0; i < 2; ++i) cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
    int buf = 0;
    for (int64_t t0 = 0; t0 < ctx->nTiles; t0 += tilesPerChunk, buf ^= 1) {
        const int64_t nt = (ctx->nTiles - t0 < tilesPerChunk) ? (ctx->nTiles - t0) : tilesPerChunk;
        const int64_t row0 = t0 * TILE_N;
        const int64_t rows = ((row0 + nt * TILE_N > ctx->N) ? ctx->N : row0 + nt * TILE_N) - row0;
        // kernel on `stream` must wait until the previous D2H out of this buffer finished
        cudaStreamWaitEvent(ctx->stream, done[buf], 0);
        logw_kernel<<<(unsigned)nt, 128, 0, ctx->stream>>>(ctx->d_u, ctx->d_L, ctx->d_c + 3 * K, K, ctx->N,
                                                          t0, d_out[buf], expo);
        ctx->launches++;
        cudaEvent_t ready;
        cudaEventCreateWithFlags(&ready, cudaEventDisableTiming);
        cudaEventRecord(ready, ctx->stream);
        cudaStreamWaitEvent(ctx->copyStream, ready, 0);
        cudaEventDestroy(ready);
        cudaError_t e = cudaMemcpy2DAsync(logW_host + row0 * ld, (size_t)ld * sizeof(double), d_out[buf],
                                          (size_t)K * sizeof(double), (size_t)K * sizeof(double),
                                          (size_t)rows, cudaMemcpyDeviceToHost, ctx->copyStream);
        cudaEventRecord(done[buf], ctx->copyStream);
        ctx->d2hBytes += rows * K * 8;
        if (e != cudaSuccess) {
            set_error("log_W download failed: %s", cudaGetErrorString(e));
            rc = MBAR_B200_ERR_CUDA;
            break;
        }
    }
    cudaError_t e = cudaStreamSynchronize(ctx->copyStream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess && rc == MBAR_B200_OK) {
        set_error("log_W failed: %s", cudaGetErrorString(e));
        rc = MBAR_B200_ERR_CUDA;
    }
    for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(done[i]);
        cudaFree(d_out[i]);
    }
    return rc;
}

}  // namespace mbar
