// rk4_pipe_probe.cu -- how much of the FP64 pipe the RK4 step's own instruction stream can
// sustain when nothing else runs (no ray setup, no sky lookup, no tile scheduling).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I blackstar_b200/csrc -o tools/rk4_pipe_probe tools/rk4_pipe_probe.cu
#include "trace_core.cuh"

#include <cstdio>
#include <cuda_runtime.h>

using namespace bsb;

template <int MINB>
__global__ void __launch_bounds__(256, MINB) probe(const __grid_constant__ FrameParams P, double *sink, int steps)
{
    double ua = 20.0 + 1e-3 * threadIdx.x, va = 0.5 + 1e-3 * blockIdx.x, du = -0.9, dv = 0.1, qa = ua * ua + va * va;
    double ub, vb, qb;
    double yh = 0.0;
    const double k4 = P.k4375;
    for (int i = 0; i < steps; i += 2) {
        rk4_step(P, ua, va, qa, du, dv, ub, vb, qb, yh, k4);
        rk4_step(P, ub, vb, qb, du, dv, ua, va, qa, yh, k4);
        if (qa < 4.0) { ua = 20.0; va = 0.5; du = -0.9; dv = 0.1; qa = ua * ua + va * va; }  // keep radii sane
    }
    if (ua + va + du + dv == 123.456) sink[0] = qa;
}

template <int MINB>
void run(const FrameParams &P, int n_sms, double nominal)
{
    double *sink; cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = n_sms * MINB, steps = 20000;
    probe<MINB><<<blocks, 256>>>(P, sink, 200);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        probe<MINB><<<blocks, 256>>>(P, sink, steps);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double rate = 64.0 * steps * 256.0 * blocks / (best * 1e-3);
    printf("  %d CTAs/SM (%d warps/SMSP): %.3f ms, %.3e DP lane-ops/s = %.1f%% of nominal, %.1f G steps/s\n", MINB, MINB * 2, best,
           rate, 100 * rate / nominal, steps * 256.0 * blocks / (best * 1e-3) / 1e9);
    cudaFree(sink);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const double nominal = 64.0 * p.multiProcessorCount * p.clockRate * 1e3;
    FrameParams P = {};
    const double h = 0.3;
    P.h = h; P.hh = h / 2; P.h6 = h / 6; P.hh2 = (h / 2) * (h / 2); P.hhh = h * (h / 2); P.hsq6 = h * h / 6; P.k4375 = 4.375;
    printf("rk4_step stream alone (64 DP + 4 MUFU per step):\n");
    run<2>(P, p.multiProcessorCount, nominal);
    run<3>(P, p.multiProcessorCount, nominal);
    run<4>(P, p.multiProcessorCount, nominal);
    run<6>(P, p.multiProcessorCount, nominal);
    run<8>(P, p.multiProcessorCount, nominal);
    return 0;
}
