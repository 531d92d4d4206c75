// patch_embed (Conv2d k8 s8 == GEMM over 8x8x64 patches, W:208,251-254; F:215,268-270; R:93,135-137) with TWO CTAs per SM.
//
// Why a second kernel: the GEMM is memory bound (it reads the 236 MB feature map of 8 frames once; 30 GFLOP), and the general
// kernel (gemm_tcgen05.cu: one persistent CTA per SM, 225 tiles of 128 tokens on 148 SMs) runs a full wave and then a wave
// that keeps only 77 SMs busy: 0.067 ms, 56 % of the copy bandwidth.  Here a CTA owns ONE tile and is small enough (3 stages
// of 32 KB, 128 TMEM columns, 192 threads) for two to share an SM, so all 225 tiles are resident at once and stream
// concurrently: every tile gets 1/225 of the HBM bandwidth and they finish together.
//   warp 0      TMA producer: per k-block (pixel (ky, kx) of every patch of the tile) a rank-5 box of the NHWC map
//               (128 tokens x 64 channels) and the matching 128 x 64 block of the filter
//   warp 1      TMEM allocation, MMA issuer (4 x 128x128x16 per k-block)
//   warps 2-5   epilogue: + bias (+ pos_embed), fp32 token written at its (window-ordered) row
#include <cuda.h>

#include "ptx.cuh"
#include "tc_api.cuh"

namespace tu {

namespace {

constexpr int E_BM = 128, E_BK = 64, E_THREADS = 192;
constexpr int E_A = E_BM * E_BK * 2;                 // 16 KB of activations per stage
// dim 128: 3 stages of 32 KB, 128 TMEM columns; dim 192 (FastTransformer): 2 stages of 40 KB, 256 TMEM columns -- two CTAs per SM both
template <int DIM> struct EmbedCfg {
    static constexpr int NST = DIM == 128 ? 3 : 2;
    static constexpr int W = DIM * E_BK * 2, STAGE = E_A + W;
    static constexpr int TMEM_COLS = DIM == 128 ? 128 : 256;
    static constexpr int SMEM = NST * STAGE + 256 + 1024;
};
constexpr int E_MAX_NST = 3;

struct EmbedParams {
    int B, Ht, Wt, nWy, nWx, window, tiles_tx, nk;
    int prefetch;       // L2 prefetch of the next ky row of the tile's patches (all 8 kx: 1 KB contiguous per patch) one ky ahead
    const float *bias, *pos;
    float *tok;
};

struct EmbedBarriers {
    uint64_t full[E_MAX_NST], empty[E_MAX_NST], acc_full;
    uint32_t tmem_base;
};

// ---- cluster helpers (MC variant: the two CTAs of a cluster load half of every filter stage each and multicast it to both)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap *m, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask, uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred pl;\n\tsetp.ne.b32 pl, %2, 0;\n\t"
        "@pl tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(bar),
        "h"(mask), "r"(leader)
        : "memory");
}

__device__ __forceinline__ void prefetch_5d(const CUtensorMap *m, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}

// MC: launched as clusters of two CTAs (two tiles); tmap_w then has a box of E_DIM / 2 rows
template <int E_DIM, bool MC>
__global__ void __launch_bounds__(E_THREADS, 2)
embed_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                const __grid_constant__ CUtensorMap tmap_pf, const EmbedParams p) {
    pdl_trigger();
    constexpr int E_NST = EmbedCfg<E_DIM>::NST, E_STAGE = EmbedCfg<E_DIM>::STAGE, E_COLS = EmbedCfg<E_DIM>::TMEM_COLS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *smem_al = smem_raw + (smem0 - ptx::smem_u32(smem_raw));
    EmbedBarriers *bars = reinterpret_cast<EmbedBarriers *>(smem_al + E_NST * E_STAGE);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < E_NST; ++i) {
            ptx::mbar_init(ptx::smem_u32(&bars->full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->empty[i]), MC ? 2 : 1);      // MC: a slot is free when BOTH CTAs' MMAs have read it
        }
        ptx::mbar_init(ptx::smem_u32(&bars->acc_full), 1);
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_w);
    }
    if (warp == 1) {
        ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), E_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const uint32_t crank = MC ? cluster_ctarank() : 0u;
    if (MC) cluster_sync();         // the peer's barriers are initialised before anything of ours can signal them
    pdl_wait();
    const int tm = blockIdx.x;      // MC: the grid is rounded up to whole clusters; a CTA past the last tile loads zeros and stores nothing

    if (warp == 0 && lane == 0) {
        const int tx0 = (tm % p.tiles_tx) * 16, r0 = (tm / p.tiles_tx) * 8;
        int stage = 0;
        uint32_t phase = 0;
        if (p.prefetch) prefetch_5d(&tmap_pf, 0, 0, tx0, 0, r0);
        for (int s = 0; s < p.nk; ++s) {
            // the 8 kx stages of one ky read 8 adjacent 128-byte pieces of every patch row: fetch the next ky's 1 KB rows into L2 in one go
            if (p.prefetch && (s & 7) == 0 && (s >> 3) + 1 < 8) prefetch_5d(&tmap_pf, 0, 0, tx0, (s >> 3) + 1, r0);
            ptx::mbar_wait(ptx::smem_u32(&bars->empty[stage]), phase ^ 1);
            const uint32_t dst = smem0 + stage * E_STAGE, fb = ptx::smem_u32(&bars->full[stage]);
            ptx::mbar_expect_tx(fb, E_STAGE);
            ptx::tma_load_5d(dst, &tmap_a, fb, 0, s & 7, tx0, s >> 3, r0);
            if (MC) tma_load_2d_mc(dst + E_A + crank * (E_DIM / 2) * 128, &tmap_w, fb, s * E_BK, crank * (E_DIM / 2), (uint16_t)3);
            else ptx::tma_load_2d(dst + E_A, &tmap_w, fb, s * E_BK, 0);
            if (++stage == E_NST) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        const uint32_t leader = ptx::elect_one();
        const uint32_t idesc = ptx::make_idesc_bf16(E_BM, E_DIM);
        int stage = 0;
        uint32_t phase = 0;
        for (int s = 0; s < p.nk; ++s) {
            ptx::mbar_wait(ptx::smem_u32(&bars->full[stage]), phase);
            ptx::tc_fence_after();
            const uint32_t a_lo = ptx::sdesc_lo(smem0 + stage * E_STAGE), w_lo = ptx::sdesc_lo(smem0 + stage * E_STAGE + E_A);
            ptx::umma_bf16_lo_rt(tmem_base, a_lo, w_lo, idesc, s > 0 ? 1u : 0u, leader);
#pragma unroll
            for (int k4 = 1; k4 < 4; ++k4) ptx::umma_bf16_lo<1>(tmem_base, a_lo + k4 * 2, w_lo + k4 * 2, idesc, leader);
            if (MC) umma_commit_mc(ptx::smem_u32(&bars->empty[stage]), (uint16_t)3, leader);
            else ptx::umma_commit_pred(ptx::smem_u32(&bars->empty[stage]), leader);
            if (++stage == E_NST) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit_pred(ptx::smem_u32(&bars->acc_full), leader);
    } else if (warp >= 2) {
        const int q = warp & 3, i = q * 32 + lane;          // TMEM lane quadrant of this warp = warp id mod 4
        const int tx = (tm % p.tiles_tx) * 16 + (i & 15);
        const int bty = (tm / p.tiles_tx) * 8 + (i >> 4);
        const int b = bty / p.Ht, ty = bty - b * p.Ht;
        const bool valid = tx < p.Wt && b < p.B;
        const long row = p.window ? (((long)b * p.nWy + (ty >> 3)) * p.nWx + (tx >> 3)) * 64 + (ty & 7) * 8 + (tx & 7)
                                  : ((long)b * p.Ht + ty) * p.Wt + tx;
        ptx::mbar_wait(ptx::smem_u32(&bars->acc_full), 0);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int n = 0; n < E_DIM; n += 32) {
            uint32_t v[32];
            ptx::tmem_ld_x32(tmem_base + ((uint32_t)(q * 32) << 16) + n, v);
            ptx::tmem_ld_wait();
            if (!valid) continue;
            float *o = p.tok + row * E_DIM + n;
            const float *pe = p.pos ? p.pos + ((long)ty * p.Wt + tx) * E_DIM + n : nullptr;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                float4 r;
                r.x = __uint_as_float(v[c + 0]) + __ldg(p.bias + n + c + 0);
                r.y = __uint_as_float(v[c + 1]) + __ldg(p.bias + n + c + 1);
                r.z = __uint_as_float(v[c + 2]) + __ldg(p.bias + n + c + 2);
                r.w = __uint_as_float(v[c + 3]) + __ldg(p.bias + n + c + 3);
                if (pe) {
                    const float4 pv = *reinterpret_cast<const float4 *>(pe + c);
                    r.x += pv.x; r.y += pv.y; r.z += pv.z; r.w += pv.w;
                }
                *reinterpret_cast<float4 *>(o + c) = r;
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (MC) cluster_sync();         // the peer may still multicast into this CTA's shared memory / arrive on its barriers
    if (warp == 1) ptx::tmem_dealloc(tmem_base, E_COLS);
}

// debug key "embed_pair" (bit mask): 0 general GEMM kernel; 1 one tile per CTA, two CTAs per SM (default); 2 the same with the filter
// stages multicast inside clusters of two CTAs; +4 also at dim 192; +8 L2 prefetch of whole patch rows one ky ahead.
// Measured (profiles/r05e_ab.log, r05f_ab.log): dim 128: 69.1 -> 62.1 us (1), 63.5 us (2), 75.3 us (1 + 8); dim 192 (two 40 KB stages per
// CTA): 134 -> 140 (1 + 4) / 145 (2 + 4) / 171 us (1 + 4 + 8): only the plain dim-128 variant is on
thread_local int g_embed_pair = 1;

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl_cluster2(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_use_pdl ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace

void tc_set_embed_pair(int on) { g_embed_pair = on; }

// same contract as tc_patch_embed (gemm_tcgen05.cu); dim 128 or 192, TU_TC_UNSUPPORTED otherwise
int tc_patch_embed_pair(const bf16 *feat, const bf16 *W, const float *bias, const float *pos, float *tok, int B, int H, int Wd, int Ht,
                        int Wt, int dim, int window, cudaStream_t st) {
    if (!(g_embed_pair & 3) || (dim != 128 && !(dim == 192 && (g_embed_pair & 4))) || !tc_encode_fn() || (reinterpret_cast<uintptr_t>(feat) & 127) || (reinterpret_cast<uintptr_t>(W) & 127) ||
        (reinterpret_cast<uintptr_t>(tok) & 15) || (pos && (reinterpret_cast<uintptr_t>(pos) & 15)) || H < 8 * Ht || Wd < 8 * Wt)
        return TU_TC_UNSUPPORTED;
    CUtensorMap ta, tw, tpf;
    {
        // rank-5 view of NHWC(64): (c, kx, tx, ky, b*ty); rows of different frames are H*Wd*128 bytes apart = Ht*8 rows only if H == 8*Ht
        if (H != 8 * Ht) return TU_TC_UNSUPPORTED;
        cuuint64_t dims[5] = {64, 8, (cuuint64_t)Wt, 8, (cuuint64_t)B * Ht};
        cuuint64_t strides[4] = {128, 1024, (cuuint64_t)Wd * 128, (cuuint64_t)Wd * 128 * 8};
        cuuint32_t box[5] = {64, 1, 16, 1, 8}, es[5] = {1, 1, 1, 1, 1};
        CUresult r = tc_encode_fn()(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void *)feat, dims, strides, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint32_t pbox[5] = {64, 8, 16, 1, 8};
        if (r == CUDA_SUCCESS)
            r = tc_encode_fn()(&tpf, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void *)feat, dims, strides, pbox, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t wd[2] = {4096, (cuuint64_t)dim}, ws[1] = {4096 * 2};
        const bool mc = (g_embed_pair & 2) != 0;
        cuuint32_t wb[2] = {E_BK, (cuuint32_t)(mc ? dim / 2 : dim)}, we[2] = {1, 1};
        if (r == CUDA_SUCCESS)
            r = tc_encode_fn()(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)W, wd, ws, wb, we, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("tu: cuTensorMapEncodeTiled(patch embed, two CTAs per SM) failed with code " + std::to_string((int)r));
            return TU_ERR_CUDA;
        }
    }
    static PerDeviceFlag attr_set;
    if (!attr_set.is_set()) {
        cudaError_t e = cudaFuncSetAttribute(embed_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EmbedCfg<128>::SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(embed_tc_kernel<192, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EmbedCfg<192>::SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(embed_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EmbedCfg<128>::SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(embed_tc_kernel<192, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EmbedCfg<192>::SMEM);
        if (e != cudaSuccess) return cuda_fail(e, "embed_tc smem attribute");
        attr_set.set();
    }
    EmbedParams p;
    p.B = B; p.Ht = Ht; p.Wt = Wt; p.nWy = (Ht + 7) / 8; p.nWx = (Wt + 7) / 8; p.window = window;
    p.tiles_tx = ceil_div(Wt, 16);
    p.nk = 4096 / E_BK;
    p.bias = bias; p.pos = pos; p.tok = tok;
    p.prefetch = (g_embed_pair & 8) ? 1 : 0;
    const int tiles_m = p.tiles_tx * ceil_div(B * Ht, 8);
    const bool mc = (g_embed_pair & 2) != 0;
    const int grid = mc ? (tiles_m + 1) & ~1 : tiles_m;
    if (mc && dim == 128) launch_pdl_cluster2(embed_tc_kernel<128, true>, dim3(grid), dim3(E_THREADS), EmbedCfg<128>::SMEM, st, ta, tw, tpf, p);
    else if (mc) launch_pdl_cluster2(embed_tc_kernel<192, true>, dim3(grid), dim3(E_THREADS), EmbedCfg<192>::SMEM, st, ta, tw, tpf, p);
    else if (dim == 128) launch_pdl(embed_tc_kernel<128, false>, dim3(grid), dim3(E_THREADS), EmbedCfg<128>::SMEM, st, ta, tw, tpf, p);
    else launch_pdl(embed_tc_kernel<192, false>, dim3(grid), dim3(E_THREADS), EmbedCfg<192>::SMEM, st, ta, tw, tpf, p);
    TU_CHECK_LAUNCH("embed_tc");
    return TU_OK;
}

}  // namespace tu
