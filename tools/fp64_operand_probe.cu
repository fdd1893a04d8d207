// Does the B200 FP64 pipe care about operand bit patterns?  8 independent DFMA chains per thread,
// x <- fma(x, a, b); `a` either has a full 53-bit mantissa or only its high word populated (the form a
// MUFU.RSQ64H seed has).  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_operand_probe fp64_operand_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) chains(double *sink, int iters, double a, double b, int mode)
{
    double x[8];
    for (int k = 0; k < 8; k++) x[k] = 1.0 + threadIdx.x * 1e-3 + k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (mode == 0) x[k] = fma(x[k], a, b);          // multiplicand: loop-invariant `a`
                else if (mode == 1) x[k] = x[k] * a;            // DMUL
                else x[k] = fma(a, a, x[k]);                    // both multiplicands `a`
            }
    }
    double s = 0;
    for (int k = 0; k < 8; k++) s += x[k];
    if (s == 123.456) sink[0] = s;
}

int main()
{
    double *sink;
    cudaMalloc(&sink, 8);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, iters = 4096;
    union { double d; unsigned long long u; } full, half;
    full.d = 0.99999912345678901;
    half.u = full.u & 0xFFFFFFFF00000000ull;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; mode++)
        for (int rep = 0; rep < 2; rep++)
            for (int which = 0; which < 2; which++) {
                const double a = which ? half.d : full.d;
                chains<<<blocks, 256>>>(sink, iters, a, 1e-9, mode);
                cudaEventRecord(e0);
                chains<<<blocks, 256>>>(sink, iters, a, 1e-9, mode);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                const double ops = 64.0 * iters * 256.0 * blocks;
                printf("mode %d (%s) a = %s mantissa: %.3f ms, %.2f T FP64 instr/s (x2 = TFLOP/s for FMA)\n", mode,
                       mode == 0 ? "x = fma(x, a, b)" : mode == 1 ? "x = x * a" : "x = fma(a, a, x)", which ? "high-word-only" : "full 53-bit",
                       ms, ops / (ms * 1e-3) / 1e12);
            }
    return 0;
}
