// Probe: which 4-D TMA box shapes (SWIZZLE_NONE) complete, and with how many transaction bytes.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
typedef CUresult (*EncFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                          const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                          CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3, uint32_t bytes, float *out, int n, int *status) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    uint32_t dst = ((uint32_t)__cvta_generic_to_shared(sm) + 127u) & ~127u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes));
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                     "l"((uint64_t)&tm), "r"(b), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    __syncthreads();
    int ok = 0;
    for (long i = 0; i < 20000000 && !ok; ++i) {
        uint32_t p;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(p) : "r"(b) : "memory");
        ok = p;
    }
    if (threadIdx.x == 0) *status = ok;
    if (ok) {
        const float *s = (const float *)(sm + (dst - (uint32_t)__cvta_generic_to_shared(sm)));
        for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = s[i];
    }
}
int main() {
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncFn enc = (EncFn)p;
    const int W = 104, H = 72, B = 2;
    std::vector<float> h((size_t)B * 3 * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o; int *st;
    cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&o, 1 << 20); cudaMalloc(&st, 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    int boxes[][3] = {{96, 26, 3}, {100, 26, 3}};
    for (auto &bx : boxes) {
        CUtensorMap tm;
        cuuint64_t dims[4] = {W, H, 3, B}, strides[3] = {W * 4ull, (cuuint64_t)H * W * 4, 3ull * H * W * 4};
        cuuint32_t box[4] = {(cuuint32_t)bx[0], (cuuint32_t)bx[1], (cuuint32_t)bx[2], 1}, es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        uint32_t bytes = bx[0] * bx[1] * bx[2] * 4;
        int hs = -1;
        if (r == CUDA_SUCCESS) {
            const int starts[] = {0, 4, -4, 8, -8, 100, 2};
            for (int start : starts) {
                k<<<1, 128, 100 * 1024>>>(tm, start, -1, 0, 1, bytes, o, bx[0] * bx[1] * bx[2], st);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(&hs, st, 4, cudaMemcpyDeviceToHost);
                float v[3]; cudaMemcpy(v, o, 12, cudaMemcpyDeviceToHost);
                printf("box %3d x %2d x %d start %2d: enc ok, sync=%s, completed=%d first=%.0f %.0f %.0f\n", bx[0], bx[1], bx[2], start, cudaGetErrorString(e), hs, v[0], v[1], v[2]);
            }
        } else printf("box %3d x %2d x %d: encode failed %d\n", bx[0], bx[1], bx[2], (int)r);
    }
    return 0;
}
