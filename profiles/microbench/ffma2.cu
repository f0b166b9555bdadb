// Micro-benchmark (not part of the library): issue throughput of scalar FFMA vs packed fma.rn.f32x2 on sm_100a,
// alone and mixed with MUFU ex2 (the likelihood kernel's instruction mix).  Register-resident.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a ffma2.cu -o ffma2 && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d) : "l"(a), "l"(b));
}

template <int MODE>
__global__ void __launch_bounds__(256) k(int iters, float* out) {
    float a[16];
    unsigned long long p[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = pack(a[2 * i], a[2 * i + 1]);
    const float m = 0.999f, c = 1e-4f;
    const unsigned long long pm = pack(m, m), pc = pack(c, c);
    float e0 = 0.5f + threadIdx.x * 1e-4f, e1 = e0 + .1f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (MODE == 0 || MODE == 2) {   // 16 scalar FMAs
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
            }
            if (MODE == 1 || MODE == 3) {   // 8 packed FMAs = 16 FMAs
#pragma unroll
                for (int i = 0; i < 8; ++i) fma2(p[i], pm, pc);
            }
            if (MODE >= 2) {                // + 2 MUFU per 16 FMAs (ratio of the likelihood kernel ~ 1:10)
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e0));
                asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(e0));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e1));
                asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(e1));
            }
        }
    }
    float s = e0 + e1;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)(p[i] & 0xffffffffu)) + __uint_as_float((unsigned)(p[i] >> 32));
    if (s == 12345.678f) out[0] = s;
}

template <int MODE> void run(const char* name, int sms) {
    float* out; cudaMalloc(&out, 64);
    const int iters = 8192, blocks = sms * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int t = 0; t < 5; ++t) {
        cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (t && ms < best) best = ms;
    }
    const double fmas = (double)blocks * 256 * iters * 64;
    printf("%-28s %.3f ms  %.2f T FMA/s (%.1f TFLOP/s)\n", name, best, fmas / best * 1e-9, 2 * fmas / best * 1e-9);
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("%s, %d SMs, %d MHz\n", pr.name, pr.multiProcessorCount, pr.clockRate / 1000);
    run<0>("scalar FFMA", pr.multiProcessorCount);
    run<1>("packed fma.f32x2", pr.multiProcessorCount);
    run<2>("scalar FFMA + MUFU 1:4", pr.multiProcessorCount);
    run<3>("packed fma.f32x2 + MUFU 1:4", pr.multiProcessorCount);
    return 0;
}
