// Micro-benchmark: issue rate of packed / scalar FP32 FMA forms on sm_100a (cycles per warp-instruction per SM sub-partition).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o morl_baselines_b200/csrc/_build/ffma2_rate scripts/ubench/ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float seed) {
    float x = seed + threadIdx.x, y = seed * 0.5f;
    u64 a[8];
    float s[8];
    for (int i = 0; i < 8; ++i) { a[i] = pk2(x + i, x - i); s[i] = x + i; }
    const u64 w = pk2(y, y + 1.f), w2 = pk2(y * 0.25f, y * 0.125f);
    const float ws = y;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fma2(a[i], w, w2);                       // FFMA2, three packed register operands (one reused)
            if (MODE == 1) a[i] = fma2(w, pk2(s[i], s[i]), a[i]);          // FFMA2 with a scalar-broadcast operand (the wp kernel's form)
            if (MODE == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(ws), "f"(y));   // scalar FFMA, 3 registers
            if (MODE == 3) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, %1;" : "+f"(s[i]) : "f"(y));    // scalar FFMA, immediate multiplier
            if (MODE == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(ws), "f"(y));      // FMNMX3
        }
    }
    long long t1 = clock64();
    float acc = 0.f;
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i])); acc += lo + hi + s[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 4096;
    const char* names[5] = {"FFMA2 3 packed regs", "FFMA2 scalar-broadcast operand", "FFMA 3 regs", "FFMA immediate", "FMNMX3"};
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int m = 0; m < 5; ++m) {
            for (int rep = 0; rep < 2; ++rep) {
                if (m == 0) k<0><<<148, warps * 32>>>(out, cyc, iters, 1.0f);
                if (m == 1) k<1><<<148, warps * 32>>>(out, cyc, iters, 1.0f);
                if (m == 2) k<2><<<148, warps * 32>>>(out, cyc, iters, 1.0f);
                if (m == 3) k<3><<<148, warps * 32>>>(out, cyc, iters, 1.0f);
                if (m == 4) k<4><<<148, warps * 32>>>(out, cyc, iters, 1.0f);
            }
            cudaDeviceSynchronize();
            long long h[148];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
            // warp-instructions per sub-partition = (warps / 4) * iters * 8
            printf("%-34s warps/SM %2d: %.2f cycles per warp-instruction per SMSP\n", names[m], warps, avg / ((warps / 4.0) * iters * 8));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
