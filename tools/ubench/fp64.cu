// FP64 pipe microbenchmark on B200: DFMA issue rate and dependent latency vs warps/SM and ILP.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
template <int ILP>
void run(int warpsPerSM, double* d) {
    int iters = 4096;
    k<ILP><<<148, warpsPerSM * 32>>>(d, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<ILP><<<148, warpsPerSM * 32>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    double inst = (double)iters * ILP * warpsPerSM;  // warp-instructions per SM
    printf("warps/SM %2d ILP %2d: %.0f cycles, %.3f DFMA warp-inst/clk/SM, %.2f cyc per dependent step, %.1f TFLOP/s\n",
           warpsPerSM, ILP, cyc, inst / cyc, cyc / iters, 148.0 * inst * 64 / (ms * 1e-3) / 1e12);
}
int main() {
    double* d; cudaMalloc(&d, 148 * 1024 * 8);
    run<1>(1, d); run<1>(4, d); run<2>(4, d); run<4>(4, d); run<8>(4, d);
    run<1>(8, d); run<4>(8, d); run<8>(8, d); run<16>(8, d);
    run<4>(16, d); run<8>(16, d); run<8>(32, d); run<4>(32, d);
    return 0;
}
