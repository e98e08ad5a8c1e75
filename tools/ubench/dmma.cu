// DMMA (mma.sync.m8n8k4.f64) issue rate on B200 vs warps/SM and independent accumulators.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP, bool MIX>
__global__ void k(double* out, int iters, double a, double b) {
    double c[ILP][2];
    double x = threadIdx.x * 1e-3;
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            dmma(c[i][0], c[i][1], a, b);
            if (MIX) { x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); }
        }
    }
    long long t1 = clock64();
    double s = x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
template <int ILP, bool MIX>
void run(int warpsPerSM, double* d) {
    int iters = 2048;
    k<ILP, MIX><<<148, warpsPerSM * 32>>>(d, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<ILP, MIX><<<148, warpsPerSM * 32>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    double inst = (double)iters * ILP * warpsPerSM;
    printf("warps/SM %2d ILP %2d mix=%d: %.3f DMMA/clk/SM (%.1f cyc per DMMA per SMSP), %.1f TFLOP/s dmma%s\n",
           warpsPerSM, ILP, (int)MIX, inst / cyc, 4.0 * cyc / inst, 148.0 * inst * 512 / (ms * 1e-3) / 1e12,
           MIX ? " + 4 DFMA per DMMA on the vector pipe" : "");
}
int main() {
    double* d; cudaMalloc(&d, 148 * 1024 * 8);
    run<1, false>(4, d); run<4, false>(4, d); run<8, false>(4, d); run<16, false>(4, d);
    run<8, false>(8, d); run<16, false>(8, d); run<8, false>(16, d); run<16, false>(16, d);
    run<8, true>(8, d); run<8, true>(16, d);
    return 0;
}
