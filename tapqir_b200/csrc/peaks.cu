// Micro-benchmarks of the two non-tensor pipes that bound the likelihood kernel:
// FP32 FMA and MUFU (ex2 / lg2).  MEASURED_PEAKS.json only has HBM and bf16 tensor peaks, so
// bench.py measures these on the box and divides by them.  Register-resident, no memory traffic.
#include "common.cuh"

namespace tq {

__global__ void __launch_bounds__(256) peak_fma_kernel(int iters, float* out) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 1e-4f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678f) out[0] = a0;
}

__global__ void __launch_bounds__(256) peak_mufu_kernel(int iters, float* out) {
    float a0 = 1.0f + threadIdx.x * 1e-3f, a1 = a0 + .1f, a2 = a0 + .2f, a3 = a0 + .3f;
    float a4 = a0 + .4f, a5 = a0 + .5f, a6 = a0 + .6f, a7 = a0 + .7f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            // ex2 then lg2: values stay in range, every op is one MUFU issue
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a5));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a7));
            asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a1));
            asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a3));
            asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a5));
            asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a7));
        }
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678f) out[0] = a0;
}

}  // namespace tq

// Launches `blocks` x 256 threads; returns the number of FMAs issued (per launch) in *ops.
extern "C" int tq_peak_fma(int blocks, int iters, void* scratch, double* ops, void* stream) {
    TQ_CHECK_ARG(blocks > 0 && iters > 0 && scratch && ops, "bad argument");
    tq::peak_fma_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, (float*)scratch);
    TQ_LAUNCH_CHECK("peak_fma_kernel launch");
    *ops = (double)blocks * 256.0 * iters * 64.0;
    return TQ_OK;
}

// Returns the number of MUFU ops issued (per launch) in *ops.
extern "C" int tq_peak_mufu(int blocks, int iters, void* scratch, double* ops, void* stream) {
    TQ_CHECK_ARG(blocks > 0 && iters > 0 && scratch && ops, "bad argument");
    tq::peak_mufu_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, (float*)scratch);
    TQ_LAUNCH_CHECK("peak_mufu_kernel launch");
    *ops = (double)blocks * 256.0 * iters * 64.0;
    return TQ_OK;
}
