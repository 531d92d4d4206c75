// Rates of the packed half-precision MUFU forms on B200: ex2.approx.ftz.f16x2, tanh.approx.f16x2 / bf16x2 against the f32 forms,
// and max.f32 with three inputs.  Results are per clock and SM, counted in ELEMENTS (a packed instruction produces two).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/mufu2_probe tools/probes/mufu2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE> __device__ __forceinline__ uint32_t op(uint32_t x) {
    uint32_t y;
    if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
    if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
    if (MODE == 2) asm volatile("tanh.approx.f32 %0, %1;" : "=r"(y) : "r"(x));
    if (MODE == 3) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
    if (MODE == 4) asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
    if (MODE == 5) asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
template <int MODE>
__global__ void k(uint32_t *out, int iters, long long *cyc) {
    uint32_t v[16];
    for (int i = 0; i < 16; ++i) v[i] = 0x3c003800u + threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = op<MODE>(v[i]);
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 16; ++i) s ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// the GELU inner step as the stack would run it: f32 pair -> polynomial (packed fp32) -> cvt f16x2 -> tanh.f16x2 -> back to f32 -> fma2 -> bf16x2
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up2(u64 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <int HALF>
__global__ void gelu_k(uint32_t *out, int iters, long long *cyc) {
    u64 x[8];
    for (int i = 0; i < 8; ++i) x[i] = pk2(0.01f * (threadIdx.x + i), -0.02f * (threadIdx.x + i));
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float a, b;
            up2(mul2(x[i], x[i]), a, b);
            const u64 x2 = pk2(fminf(a, 64.f), fminf(b, 64.f));
            u64 q = fma2(pk2(-0.00035f, -0.00035f), x2, pk2(0.037f, 0.037f));
            q = fma2(q, x2, pk2(0.7975f, 0.7975f));
            float u0, u1, t0f, t1f;
            up2(mul2(x[i], q), u0, u1);
            if (HALF) {
                uint32_t h, th;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(u1), "f"(u0));
                asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(th) : "r"(h));
                asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %2; cvt.f32.f16 %0, lo; cvt.f32.f16 %1, hi;}" : "=f"(t0f), "=f"(t1f) : "r"(th));
            } else {
                asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t0f) : "f"(u0));
                asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t1f) : "f"(u1));
            }
            const u64 hx = mul2(x[i], pk2(0.5f, 0.5f));
            const u64 r = fma2(hx, pk2(t0f, t1f), hx);
            float r0, r1;
            up2(r, r0, r1);
            uint32_t pb;
            asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pb) : "f"(r1), "f"(r0));
            acc ^= pb;
            x[i] = fma2(r, pk2(0.999f, 0.999f), pk2(0.001f, 0.001f));
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int M, typename F> void run(const char *name, F launch, int elems_per_thread_iter) {
    (void)M;
}
int main() {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4 * sizeof(uint32_t)); cudaMalloc(&cyc, 8);
    const int iters = 4000;
    const char *names[6] = {"ex2.f32", "ex2.f16x2", "tanh.f32", "tanh.f16x2", "tanh.bf16x2", "ex2.bf16x2"};
    for (int mode = 0; mode < 6; ++mode)
        for (int warps : {8, 16}) {
            long long c = 0;
            for (int rep = 0; rep < 2; ++rep) {
                switch (mode) {
                    case 0: k<0><<<148, warps * 32>>>(out, iters, cyc); break;
                    case 1: k<1><<<148, warps * 32>>>(out, iters, cyc); break;
                    case 2: k<2><<<148, warps * 32>>>(out, iters, cyc); break;
                    case 3: k<3><<<148, warps * 32>>>(out, iters, cyc); break;
                    case 4: k<4><<<148, warps * 32>>>(out, iters, cyc); break;
                    case 5: k<5><<<148, warps * 32>>>(out, iters, cyc); break;
                }
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            const int per = (mode == 0 || mode == 2) ? 1 : 2;
            printf("%-12s warps/SM %2d: %.2f instr per clk per SM = %.2f elements per clk per SM\n", names[mode], warps,
                   (double)iters * 16 * warps * 32 / c, (double)iters * 16 * warps * 32 * per / c);
        }
    for (int half = 0; half < 2; ++half)
        for (int warps : {8, 16}) {
            long long c = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (half) gelu_k<1><<<148, warps * 32>>>(out, iters, cyc); else gelu_k<0><<<148, warps * 32>>>(out, iters, cyc);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("gelu step (%s) warps/SM %2d: %.2f elements per clk per SM\n", half ? "tanh.f16x2" : "tanh.f32", warps,
                   (double)iters * 16 * warps * 32 / c);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
