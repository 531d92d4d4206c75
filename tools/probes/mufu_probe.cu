// MUFU.EX2 throughput on B200 vs warps per SM, alone and inside the softmax inner loop (FFMA2 + EX2 + FADD2 + F2FP).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/mufu_probe tools/probes/mufu_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up2(u64 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int MODE>
__global__ void k(float *out, int iters, long long *cyc) {
    float v[16];
    for (int i = 0; i < 16; ++i) v[i] = -0.001f * (threadIdx.x + i);
    u64 rs = pk2(0.f, 0.f);
    uint32_t acc = 0;
    const u64 c2 = pk2(1.4426950408889634f, 1.4426950408889634f), nb = pk2(-0.5f, -0.5f);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = ex2f(v[i]);
        } else {
            float e[16];
#pragma unroll
            for (int i = 0; i < 16; i += 2) up2(fma2(pk2(v[i], v[i + 1]), c2, nb), e[i], e[i + 1]);
#pragma unroll
            for (int i = 0; i < 16; ++i) e[i] = ex2f(e[i]);
#pragma unroll
            for (int i = 0; i < 16; i += 2) { rs = add2(rs, pk2(e[i], e[i + 1])); acc ^= pack_bf16(e[i], e[i + 1]); }
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = e[i] - 1.0f;      // keep the chain alive without another MUFU
        }
    }
    long long t1 = clock64();
    float a, b; up2(rs, a, b);
    float s = a + b + __uint_as_float(acc & 0x3f800000u);
    for (int i = 0; i < 16; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4 * sizeof(float)); cudaMalloc(&cyc, 8);
    const int iters = 4000;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps : {4, 8, 16, 32}) {
            long long c = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, cyc); else k<1><<<148, warps * 32>>>(out, iters, cyc);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            double exps = (double)iters * 16 * warps * 32;
            printf("mode %d (%s) warps/SM %2d: %.2f ex2 per clk per SM\n", mode, mode ? "fma2+ex2+add2+f2fp" : "ex2 only", warps, exps / c);
        }
    return 0;
}
