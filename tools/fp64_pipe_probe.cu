// fp64_pipe_probe.cu -- what the FP64 pipe of a B200 SM actually sustains for different
// instruction/operand patterns (roofline denominator sanity check).  Standalone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe_probe tools/fp64_pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(256) probe(double *sink, int iters, double a, double b)
{
    double x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = 0.999999 + 1e-9 * i + a * 1e-12; z[i] = 1e-9 * (i + 1) + b; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (KIND == 0) x[i] = fma(x[i], a, b);                 // 1 register + 2 shared operands
                if (KIND == 1) x[i] = fma(x[i], y[i], z[i]);           // 3 distinct registers
                if (KIND == 2) x[i] = x[i] * y[i];                     // DMUL
                if (KIND == 3) x[i] = x[i] + z[i];                     // DADD
                if (KIND == 4) { if (k & 1) x[i] = fma(x[i], y[i], z[i]); else if (k & 2) x[i] = x[i] * y[i]; else x[i] = x[i] + z[i]; }
                if (KIND == 5) { x[i] = fma(x[i], y[i], z[(i + 1) & 7]); y[i] = y[i] * y[(i + 3) & 7]; }  // 2 streams, cross operands
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i] + y[i];
    if (s == 123.456) sink[0] = s;
}

template <int KIND>
double run(int blocks, int iters, double ops_per_iter)
{
    double *sink; cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<KIND><<<blocks, 256>>>(sink, 64, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        probe<KIND><<<blocks, 256>>>(sink, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFree(sink);
    return ops_per_iter * iters * 256.0 * blocks / (best * 1e-3);   // DP instructions (lane-ops) per second
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const double nominal = 64.0 * p.multiProcessorCount * p.clockRate * 1e3;  // lane-ops/s at max clock
    printf("%s: %d SMs, max clock %.0f MHz, nominal 64 DP lanes/clk/SM = %.3e lane-ops/s\n", p.name, p.multiProcessorCount, p.clockRate / 1e3, nominal);
    for (int occ : {4, 8}) {
        const int blocks = p.multiProcessorCount * occ;
        printf("resident CTAs/SM = %d (256 threads each)\n", occ);
        double r;
        r = run<0>(blocks, 2048, 64);  printf("  DFMA x,a,b (shared operands)   : %.3e /s = %.1f%% of nominal\n", r, 100 * r / nominal);
        r = run<1>(blocks, 2048, 64);  printf("  DFMA 3 distinct registers      : %.3e /s = %.1f%%\n", r, 100 * r / nominal);
        r = run<2>(blocks, 2048, 64);  printf("  DMUL                           : %.3e /s = %.1f%%\n", r, 100 * r / nominal);
        r = run<3>(blocks, 2048, 64);  printf("  DADD                           : %.3e /s = %.1f%%\n", r, 100 * r / nominal);
        r = run<4>(blocks, 2048, 64);  printf("  mix DFMA/DMUL/DADD             : %.3e /s = %.1f%%\n", r, 100 * r / nominal);
        r = run<5>(blocks, 2048, 128); printf("  2 streams, cross operands      : %.3e /s = %.1f%%\n", r, 100 * r / nominal);
    }
    return 0;
}
