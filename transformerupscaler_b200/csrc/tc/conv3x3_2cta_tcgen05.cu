// 3x3 / pad-1 convolution 64 -> 64 channels on a CTA PAIR (tcgen05 cta_group::2): the variant of conv3x3_tcgen05.cu
// used for the plain convolutions (conv2, downsample, decoder_conv1: one 64-channel output chunk, no PixelShuffle).
//
// Why a pair: with N = 64 output channels a 128x64x16 MMA reads 4 KB of activations + 2 KB of filter taps from shared
// memory every 32 cycles, and profiles/ (sm__throughput 88 %, tensor pipe 58 %) show the single-CTA kernel bound by
// exactly that traffic.  With cta_group::2 one instruction drives both SMs of a TPC: M = 256 = the two CTAs' own
// 128-pixel row segments, while the 64 x 64 tap is SPLIT between them (each CTA keeps and reads only 32 of the 64
// output-channel rows), so the filter traffic per SM halves and the filter bank shrinks to 36 KB.
//
// Protocol (rank 0 = leader, the only CTA that issues MMAs; all barrier objects exist at the same offset in both CTAs):
//   * each CTA's producer streams the input rows of ITS tile into ITS ring (TMA, local `full` barrier);
//   * the peer's MMA warp forwards every local `full` to the leader's `peer_full` (remote mbarrier arrive), so the leader
//     issues the MMAs of a step when both CTAs' rows have landed;
//   * tcgen05.commit ... multicast::cluster releases `empty[slot]` / publishes `acc_full[set]` in BOTH CTAs;
//   * the epilogue warps of both CTAs arrive on the leader's `acc_empty[set]` (remote arrive for the peer).
// Tiles are taken in pairs (2k, 2k+1); an odd tail pairs the last tile with a dummy one whose loads fall outside the
// tensor (zero-filled) and whose stores are skipped.
//
// MEASURED (profiles/r02n_conv_2cta_ncu_summary.txt, 8 x 720p conv2): correct (all parity tests pass with it enabled) but
// SLOWER than the single-CTA kernel: 0.526 ms vs 0.457 ms.  Shared-memory pressure does drop (sm__throughput 88 % -> 67 %),
// but the tensor pipe is busy only 49 % of the time: a 32-cycle N = 64 MMA is too short to hide the pair's operand
// exchange (each SM fetches the other half of the tap from its peer for every instruction), so the issue cadence settles at
// ~65 cycles per MMA.  cta_group::2 pays off with N >= 128 tiles; this model's channel count is 64.  The kernel is kept
// behind tu_debug_set("conv_2cta", 1) as the A/B evidence and is not used by default.
#include <cuda.h>
#include <string.h>

#include "ptx.cuh"
#include "tc_api.cuh"

namespace tu {

namespace {

constexpr int TILE_R = 4, TILE_M = 128, BOXW = 136;
constexpr int UNIT_BYTES = BOXW * 128;    // 17408
constexpr int RING_UNITS = 9;             // deeper than one tile's 6 steps: refilling a slot costs a TMA round trip plus the pair hand-off
constexpr int W_HALF_TAP = 32 * 128;      // this CTA's 32 output-channel rows of one tap
constexpr int W_BYTES = 9 * W_HALF_TAP;   // 36864
constexpr int STG_BYTES = 4 * 2 * 4096;
constexpr int SMEM_BYTES = W_BYTES + RING_UNITS * UNIT_BYTES + STG_BYTES + 512 + 1024;
constexpr int NUM_THREADS = 256;

struct ConvParams2 {
    int B, H, W, Ho, Wo, relu;
    int tiles_x, tiles_y, total_tiles;
    const float *bias;
};

struct Barriers2 {
    uint64_t full[RING_UNITS], peer_full[RING_UNITS], empty[RING_UNITS];
    uint64_t acc_full[2], acc_empty[2];
    uint64_t w_full, peer_w_full;
    uint32_t tmem_base;
};

// ---- cluster / pair helpers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same object in CTA `rank`
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, both CTAs] (+)= A[smem of each CTA: its 128 rows] * B[smem: 32 of the 64 rows in each CTA]^T
template <int ACC>
__device__ __forceinline__ void umma2_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred pl, pa;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 pl, %4, 0;\n\t"
        "setp.ne.b32 pa, %6, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "@pl tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, pa;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(leader), "r"(ptx::SDESC_HI_SW128), "n"(ACC)
        : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once all MMAs issued so far have retired
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred pl;\n\t.reg .b16 m;\n\tsetp.ne.b32 pl, %1, 0;\n\tmov.b16 m, 3;\n\t"
        "@pl tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(bar),
        "r"(leader)
        : "memory");
}

template <int S>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv3x3_tc2_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_w,
                   const __grid_constant__ CUtensorMap tmap_out, const ConvParams2 p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_sm = smem0, ring_sm = smem0 + W_BYTES, stg_sm = ring_sm + RING_UNITS * UNIT_BYTES;
    uint8_t *smem_al = smem_raw + (smem0 - ptx::smem_u32(smem_raw));
    Barriers2 *bars = reinterpret_cast<Barriers2 *>(smem_al + W_BYTES + RING_UNITS * UNIT_BYTES + STG_BYTES);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int npair_tiles = (p.total_tiles + 1) >> 1;
    constexpr int units_per_step = S;
    constexpr int nslots = RING_UNITS / units_per_step;     // 9 or 4
    constexpr int nsteps = S == 1 ? TILE_R + 2 : 2 * TILE_R + 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < RING_UNITS; ++i) {
            ptx::mbar_init(ptx::smem_u32(&bars->full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->peer_full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->empty[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(ptx::smem_u32(&bars->acc_full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->acc_empty[i]), 8);      // 4 epilogue warps of each CTA (leader's copy is used)
        }
        ptx::mbar_init(ptx::smem_u32(&bars->w_full), 1);
        ptx::mbar_init(ptx::smem_u32(&bars->peer_w_full), 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc2(ptx::smem_u32(&bars->tmem_base), 512);
        tmem_relinquish2();
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_act);
        ptx::prefetch_tmap(&tmap_w);
        ptx::prefetch_tmap(&tmap_out);
    }
    ptx::tc_fence_before();
    __syncthreads();
    cluster_sync_all();           // both CTAs' barriers are initialised before anyone signals across the pair
    ptx::tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    // tile of this CTA in pair-iteration tp; a dummy tile (odd tail) has b == B: loads zero-fill, stores are skipped
    auto tile_coords = [&](int tp, int &tx, int &ty, int &b) {
        int t = 2 * tp + (int)rank;
        if (t >= p.total_tiles) { tx = 0; ty = 0; b = p.B; return; }
        tx = t % p.tiles_x;
        t /= p.tiles_x;
        ty = t % p.tiles_y;
        b = t / p.tiles_y;
    };

    if (warp == 0 && lane == 0) {
        // ================================ TMA producer (each CTA: its own tile, its half of the filter bank) ================================
        ptx::mbar_expect_tx(ptx::smem_u32(&bars->w_full), W_BYTES);
        for (int tap = 0; tap < 9; ++tap)
            ptx::tma_load_2d(w_sm + tap * W_HALF_TAP, &tmap_w, ptx::smem_u32(&bars->w_full), 0, tap * 64 + (int)rank * 32);
        int slot = 0;
        uint32_t phase = 0;
        for (int tp = pair; tp < npair_tiles; tp += npairs) {
            int tx, ty, b;
            tile_coords(tp, tx, ty, b);
            const int x0 = tx * TILE_M, y0 = ty * TILE_R;
            for (int j = 0; j < nsteps; ++j) {
                ptx::mbar_wait(ptx::smem_u32(&bars->empty[slot]), phase ^ 1);
                const uint32_t dst = ring_sm + slot * units_per_step * UNIT_BYTES;
                const uint32_t fb = ptx::smem_u32(&bars->full[slot]);
                ptx::mbar_expect_tx(fb, units_per_step * UNIT_BYTES);
                if (S == 1) {
                    ptx::tma_load_4d(dst, &tmap_act, fb, 0, x0 - 1, y0 - 1 + j, b);
                } else {
                    const int iy = 2 * y0 - 1 + j;
                    ptx::tma_load_4d(dst, &tmap_act, fb, 0, x0, iy, b);
                    ptx::tma_load_4d(dst + UNIT_BYTES, &tmap_act, fb, 64, x0 - 1, iy, b);
                }
                if (++slot == nslots) { slot = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && rank != 0) {
        // ================================ peer: forward "my rows have landed" to the leader ================================
        if (lane == 0) {
            ptx::mbar_wait(ptx::smem_u32(&bars->w_full), 0);
            mbar_arrive_remote(map_to_rank(ptx::smem_u32(&bars->peer_w_full), 0));
            int slot = 0;
            uint32_t phase = 0;
            for (int tp = pair; tp < npair_tiles; tp += npairs) {
                for (int j = 0; j < nsteps; ++j) {
                    ptx::mbar_wait(ptx::smem_u32(&bars->full[slot]), phase);
                    mbar_arrive_remote(map_to_rank(ptx::smem_u32(&bars->peer_full[slot]), 0));
                    if (++slot == nslots) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ leader: MMA issuer for the pair ================================
        const uint32_t leader = ptx::elect_one();
        const uint32_t idesc = ptx::make_idesc_bf16(2 * TILE_M, 64);
        const uint32_t w_lo = ptx::sdesc_lo(w_sm), ring_lo = ptx::sdesc_lo(ring_sm);
        ptx::mbar_wait(ptx::smem_u32(&bars->w_full), 0);
        mbar_wait_cluster(ptx::smem_u32(&bars->peer_w_full), 0);
        int slot = 0;
        uint32_t ph = 0;
        int it = 0;
        for (int tp = pair; tp < npair_tiles; tp += npairs, ++it) {
            const int set = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait_cluster(ptx::smem_u32(&bars->acc_empty[set]), aphase ^ 1);
            ptx::tc_fence_after();
            const uint32_t acc0 = tmem_base + set * (TILE_R * 64);
#pragma unroll
            for (int j = 0; j < nsteps; ++j) {
                ptx::mbar_wait(ptx::smem_u32(&bars->full[slot]), ph);
                mbar_wait_cluster(ptx::smem_u32(&bars->peer_full[slot]), ph);
                ptx::tc_fence_after();
                const uint32_t a0 = ring_lo + ((slot * units_per_step * UNIT_BYTES) >> 4);
                if (S == 1) {
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int r = j - ky;
                        if (r < 0 || r >= TILE_R) continue;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                const uint32_t ad = a0 + ((kx * 128 + k4 * 32) >> 4);
                                const uint32_t bd = w_lo + (((ky * 3 + kx) * W_HALF_TAP + k4 * 32) >> 4);
                                if ((ky | kx | k4) != 0) umma2_bf16_lo<1>(acc0 + r * 64, ad, bd, idesc, leader);
                                else umma2_bf16_lo<0>(acc0 + r * 64, ad, bd, idesc, leader);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int r = (j >> 1) - q;
                        const int ky = j - 2 * r;
                        if (r < 0 || r >= TILE_R || ky > 2) continue;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const int aoff = kx == 1 ? 0 : UNIT_BYTES + (kx == 2 ? 128 : 0);
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                const uint32_t ad = a0 + ((aoff + k4 * 32) >> 4);
                                const uint32_t bd = w_lo + (((ky * 3 + kx) * W_HALF_TAP + k4 * 32) >> 4);
                                if ((ky | kx | k4) != 0) umma2_bf16_lo<1>(acc0 + r * 64, ad, bd, idesc, leader);
                                else umma2_bf16_lo<0>(acc0 + r * 64, ad, bd, idesc, leader);
                            }
                        }
                    }
                }
                umma2_commit_mc(ptx::smem_u32(&bars->empty[slot]), leader);      // both CTAs' slot is reusable
                if (++slot == nslots) { slot = 0; ph ^= 1; }
            }
            umma2_commit_mc(ptx::smem_u32(&bars->acc_full[set]), leader);         // both CTAs' accumulators are complete
        }
    } else if (warp >= 4) {
        // ================================ epilogue (each CTA: its own 128 pixels x 64 channels) ================================
        const int q = warp - 4;
        int it = 0;
        uint32_t nstore = 0;
        uint8_t *stg_w = smem_al + W_BYTES + RING_UNITS * UNIT_BYTES + q * 8192;
        const uint32_t stg_w_sm = stg_sm + q * 8192;
        const uint32_t acc_empty0[2] = {map_to_rank(ptx::smem_u32(&bars->acc_empty[0]), 0), map_to_rank(ptx::smem_u32(&bars->acc_empty[1]), 0)};
        for (int tp = pair; tp < npair_tiles; tp += npairs, ++it) {
            int tx, ty, b;
            tile_coords(tp, tx, ty, b);
            const int set = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            ptx::mbar_wait(ptx::smem_u32(&bars->acc_full[set]), aphase);
            ptx::tc_fence_after();
            const int px0 = tx * TILE_M + q * 32;
#pragma unroll 1
            for (int r = 0; r < TILE_R; ++r) {
                const int y = ty * TILE_R + r;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + set * (TILE_R * 64) + r * 64;
                uint32_t v0[32], v1[32];
                ptx::tmem_ld_x32(taddr, v0);
                ptx::tmem_ld_x32(taddr + 32, v1);
                ptx::tmem_ld_wait();
                if (r == TILE_R - 1) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(set ? acc_empty0[1] : acc_empty0[0]);
                }
                const uint32_t buf = nstore & 1;
                if (lane == 0) ptx::bulk_wait_read<1>();
                __syncwarp();
                uint8_t *rowp = stg_w + buf * 4096 + lane * 128;
#pragma unroll
                for (int c = 0; c < 64; c += 8) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float a = __uint_as_float(c < 32 ? v0[c + e] : v1[c - 32 + e]) + (p.bias ? __ldg(p.bias + c + e) : 0.f);
                        f[e] = p.relu ? fmaxf(a, 0.f) : a;
                    }
                    uint4 u;
                    __nv_bfloat162 h;
                    h = __floats2bfloat162_rn(f[0], f[1]); u.x = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(f[2], f[3]); u.y = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(f[4], f[5]); u.z = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(f[6], f[7]); u.w = *reinterpret_cast<uint32_t *>(&h);
                    *reinterpret_cast<uint4 *>(rowp + ((((c >> 3) ^ (lane & 7))) << 4)) = u;
                }
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if (b < p.B && y < p.Ho && px0 < p.Wo) ptx::tma_store_4d(&tmap_out, stg_w_sm + buf * 4096, 0, px0, y, b);
                    ptx::bulk_commit();
                }
                ++nstore;
            }
        }
        if (lane == 0) ptx::bulk_wait<0>();
        __syncwarp();
    }
    ptx::tc_fence_before();
    __syncthreads();
    cluster_sync_all();           // the peer's shared memory and barriers stay alive until the leader is done with them
    if (warp == 2) tmem_dealloc2(tmem_base, 512);
}

PerDeviceFlag g_attr_set2;
thread_local int g_enable_2cta = 0;       // measured slower than the single-CTA kernel on B200 (see the header comment): off unless tu_debug_set("conv_2cta", 1)

}  // namespace

void tc_set_conv_2cta(int on) { g_enable_2cta = on; }

// Returns TU_TC_UNSUPPORTED for anything but a plain 64 -> 64 convolution (the single-CTA kernel handles the rest).
int tc_conv3x3_c64_pair(const bf16 *in, const bf16 *w, const float *bias, bf16 *out, int B, int H, int W, int stride, int relu,
                        cudaStream_t st) {
    if (!g_enable_2cta) return TU_TC_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(in) & 127) || (reinterpret_cast<uintptr_t>(w) & 127) || (reinterpret_cast<uintptr_t>(out) & 15))
        return TU_TC_UNSUPPORTED;
    if (stride == 2 && (W & 1)) return TU_TC_UNSUPPORTED;
    TcEncodeFn enc = tc_encode_fn();
    if (!enc) return TU_TC_UNSUPPORTED;
    const int g_sm_count2 = device_sm_count();
    if (!g_attr_set2.is_set()) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "conv3x3_tc2 smem attribute");
        g_attr_set2.set();
    }
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    CUtensorMap tm_act, tm_w, tm_out;
    {
        cuuint64_t dims[4], strides[3];
        cuuint32_t box[4] = {64, (cuuint32_t)BOXW, 1, 1}, estr[4] = {1, 1, 1, 1};
        if (stride == 1) {
            dims[0] = 64; dims[1] = (cuuint64_t)W; dims[2] = (cuuint64_t)H; dims[3] = (cuuint64_t)B;
            strides[0] = 128; strides[1] = (cuuint64_t)W * 128; strides[2] = (cuuint64_t)H * W * 128;
        } else {
            dims[0] = 128; dims[1] = (cuuint64_t)(W / 2); dims[2] = (cuuint64_t)H; dims[3] = (cuuint64_t)B;
            strides[0] = 256; strides[1] = (cuuint64_t)W * 128; strides[2] = (cuuint64_t)H * W * 128;
        }
        CUresult r = enc(&tm_act, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t wd[2] = {64, 9 * 64}, ws[1] = {128};
        cuuint32_t wb[2] = {64, 32}, we[2] = {1, 1};
        if (r == CUDA_SUCCESS)
            r = enc(&tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)w, wd, ws, wb, we, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t od[4] = {64, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)B}, os[3] = {128, (cuuint64_t)Wo * 128, (cuuint64_t)Ho * Wo * 128};
        cuuint32_t ob[4] = {64, 32, 1, 1}, oe[4] = {1, 1, 1, 1};
        if (r == CUDA_SUCCESS)
            r = enc(&tm_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)out, od, os, ob, oe, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("tu: cuTensorMapEncodeTiled(conv pair) failed with code " + std::to_string((int)r));
            return TU_ERR_CUDA;
        }
    }
    ConvParams2 p;
    p.B = B; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo; p.relu = relu;
    p.tiles_x = ceil_div(Wo, TILE_M); p.tiles_y = ceil_div(Ho, TILE_R);
    p.total_tiles = p.tiles_x * p.tiles_y * B;
    p.bias = bias;
    const int npair_tiles = (p.total_tiles + 1) / 2;
    int grid = 2 * (npair_tiles < g_sm_count2 / 2 ? npair_tiles : g_sm_count2 / 2);
    if (stride == 1)
        conv3x3_tc2_kernel<1><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tm_act, tm_w, tm_out, p);
    else
        conv3x3_tc2_kernel<2><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tm_act, tm_w, tm_out, p);
    TU_CHECK_LAUNCH("conv3x3_tc2");
    return TU_OK;
}

}  // namespace tu
