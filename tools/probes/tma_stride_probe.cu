// Probe: rank-5 TMA with elementStrides {1,1,1,8,1} and SWIZZLE_128B: load a (c 64, dx 1, tx 8, y 4 rows at stride 8, b 1)
// box from an NHWC tensor, add 1, store it back through the same kind of map into a second tensor; check on the host.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
typedef CUresult (*EncFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                          const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                          CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, int dx, int tx0, int y0, int b, int *status) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    uint32_t bb = (uint32_t)__cvta_generic_to_shared(&bar);
    uint32_t dst = ((uint32_t)__cvta_generic_to_shared(sm) + 1023u) & ~1023u;
    uint8_t *p = sm + (dst - (uint32_t)__cvta_generic_to_shared(sm));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bb));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb), "r"(4096));
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
                     "l"((uint64_t)&tin), "r"(bb), "r"(0), "r"(dx), "r"(tx0), "r"(y0), "r"(b) : "memory");
    }
    __syncthreads();
    int ok = 0;
    for (long i = 0; i < 20000000 && !ok; ++i) {
        uint32_t q;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(q) : "r"(bb) : "memory");
        ok = q;
    }
    if (threadIdx.x == 0) *status = ok;
    if (!ok) return;
    // thread = row (32 rows), add 1 to every element (swizzle-agnostic: touch all 128 bytes of the row)
    __nv_bfloat16 *row = reinterpret_cast<__nv_bfloat16 *>(p + threadIdx.x * 128);
    for (int c = 0; c < 64; ++c) row[c] = __float2bfloat16(__bfloat162float(row[c]) + 1.0f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"((uint64_t)&tout), "r"(dst),
                     "r"(0), "r"(dx), "r"(tx0), "r"(y0), "r"(b) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
int main() {
    void *pf = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &pf, cudaEnableDefault, &q);
    EncFn enc = (EncFn)pf;
    const int B = 2, H = 45, W = 80, Wt = W / 8;     // H not a multiple of 8 rows x 8: clipping in y
    const size_t n = (size_t)B * H * W * 64;
    std::vector<__nv_bfloat16> h(n), o(n);
    for (size_t i = 0; i < n; ++i) h[i] = __float2bfloat16((float)(i % 251));
    __nv_bfloat16 *din, *dout; int *st;
    cudaMalloc(&din, n * 2); cudaMalloc(&dout, n * 2); cudaMalloc(&st, 4);
    cudaMemcpy(din, h.data(), n * 2, cudaMemcpyHostToDevice); cudaMemset(dout, 0, n * 2);
    CUtensorMap tin, tout;
    cuuint64_t dims[5] = {64, 8, (cuuint64_t)Wt, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[4] = {128, 1024, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
    cuuint32_t box[5] = {64, 1, 8, 32, 1}, es[5] = {1, 1, 1, 8, 1};
    CUresult r1 = enc(&tin, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, din, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&tout, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dout, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode in=%d out=%d\n", (int)r1, (int)r2);
    if (r1 || r2) return 0;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024);
    const int cases[][4] = {{3, 0, 5, 0}, {7, 2, 32 + 2, 1}};     // dx, tx0, y0, b ; second case: y0 + 24 = 58 > 44 -> clipped rows
    for (auto &c : cases) {
        cudaMemset(dout, 0, n * 2);
        k<<<1, 32, 8 * 1024>>>(tin, tout, c[0], c[1], c[2], c[3], st);
        cudaError_t e = cudaDeviceSynchronize();
        int hs = -1; cudaMemcpy(&hs, st, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(o.data(), dout, n * 2, cudaMemcpyDeviceToHost);
        long bad = 0, touched = 0, expect = 0;
        for (int b = 0; b < B; ++b) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int ch = 0; ch < 64; ++ch) {
            size_t i = (((size_t)b * H + y) * W + x) * 64 + ch;
            bool in = b == c[3] && (x % 8) == c[0] && (x / 8) >= c[1] && (x / 8) < c[1] + 8 && y >= c[2] && (y - c[2]) % 8 == 0 && (y - c[2]) / 8 < 4;
            float want = in ? __bfloat162float(h[i]) + 1.0f : 0.f;
            float got = __bfloat162float(o[i]);
            if (in) ++expect;
            if (got != 0.f) ++touched;
            if (got != want) ++bad;
        }
        printf("case dx=%d tx0=%d y0=%d b=%d: sync=%s completed=%d expect=%ld touched=%ld mismatches=%ld\n", c[0], c[1], c[2], c[3], cudaGetErrorString(e), hs, expect, touched, bad);
    }
    return 0;
}
