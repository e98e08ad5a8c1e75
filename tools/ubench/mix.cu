// Does a DFMA occupy the warp scheduler for 2 cycles, or only the fp64 pipe?  Mix DFMA with
// independent integer / fp32 / shuffle / LDS instructions and watch the DFMA rate.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int NX>
__global__ void k(double* out, int iters, double a, double b, int ia) {
    __shared__ double sm[1024];
    sm[threadIdx.x & 1023] = threadIdx.x;
    __syncthreads();
    double x[8];
    int y[8];
    float z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3 + i; y[i] = threadIdx.x + i; z[i] = i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = fma(x[i], a, b);
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                if (MODE == 1) y[i] = y[i] * ia + it;                       // IMAD
                if (MODE == 2) z[i] = fmaf(z[i], 1.0001f, 0.5f);            // FFMA
                if (MODE == 3) y[i] = __shfl_sync(0xffffffffu, y[i], y[i]); // SHFL.IDX
                if (MODE == 4) y[i] = (y[i] >> 3) ^ ia;                     // SHF + LOP3 (ALU)
                if (MODE == 5) y[i] += (int)sm[(y[i] + j) & 1023];          // LDS (+conv)
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
template <int MODE, int NX>
void run(const char* name, int warps, double* d) {
    int iters = 2048;
    k<MODE, NX><<<148, warps * 32>>>(d, iters, 1.0000001, 1e-9, 3);
    cudaDeviceSynchronize();
    k<MODE, NX><<<148, warps * 32>>>(d, iters, 1.0000001, 1e-9, 3);
    cudaDeviceSynchronize();
    double cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    double dfma = (double)iters * 8 * warps / 4;   // per SMSP
    printf("%-28s warps/SM %2d: %.2f cycles per DFMA per SMSP (extra insts per DFMA: %d)\n", name, warps,
           cyc / dfma, NX);
}
int main() {
    double* d; cudaMalloc(&d, 148 * 1024 * 8);
    for (int w : {8, 16}) {
        if (w == 8) {
            run<0, 0>("DFMA only", 8, d); run<1, 1>("+1 IMAD", 8, d); run<1, 2>("+2 IMAD", 8, d);
            run<2, 1>("+1 FFMA", 8, d); run<2, 2>("+2 FFMA", 8, d); run<3, 1>("+1 SHFL", 8, d);
            run<4, 1>("+1 SHF+LOP3", 8, d); run<4, 2>("+2 SHF+LOP3", 8, d); run<5, 1>("+1 LDS", 8, d);
        } else {
            run<0, 0>("DFMA only", 16, d); run<1, 1>("+1 IMAD", 16, d); run<1, 2>("+2 IMAD", 16, d);
            run<4, 2>("+2 SHF+LOP3", 16, d);
        }
    }
    return 0;
}
