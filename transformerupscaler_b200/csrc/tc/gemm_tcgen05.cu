// C[M,N] = A[M,K] * W[N,K]^T on the 5th-gen tensor cores (bf16 operands, fp32 accumulation in TMEM), with the
// epilogues the transformer part of the models needs.
//
// Reference call sites: qkv / proj / mlp Linear layers (WindowTransformer/model.py:77-79,144-148,113,129,168;
// ResidualTransformer/model.py:31-37), patch_embed = Conv2d k8 s8 (W:208,251; F:215,268; R:93,135) and
// patch_unembed = ConvTranspose2d k8 s8 + crop + skip add (W:218,287-294; F:225,302-309; R:108,150-153).
//
// Per CTA (persistent over output tiles of 128 rows x BN columns, n fastest so an A block is re-read from L2):
//   warp 0: TMA producer — A tile (128 rows x 64 k, 128-byte swizzle) and W tile (BN rows x 64 k) per stage, 4 stages.
//           A is either a plain row-major matrix or, for patch embed, a rank-5 view of the NHWC feature map
//           (c, kx, tx, ky, b*ty) so that the 8x8x64 patch gather is done by the TMA engine: stage s = pixel
//           (ky,kx) = s/8, s%8 of every patch of the tile, 64 channels = one swizzle row.
//   warp 1: one thread issues 4 tcgen05.mma (128 x BN x 16) per stage into one of two TMEM accumulator sets.
//   warps 4-7: epilogue, thread = output row: tcgen05.ld 32 columns at a time, then
//           STORE  : + bias, optional exact GELU, bf16 row-major
//           RESID  : x[m][n] += acc + bias on the fp32 token stream (optionally also a bf16 copy for the next GEMM)
//           EMBED  : + bias (+ pos_embed), fp32 token written at its window-ordered row
//           UNEMBED: column n = (ky,kx,c) scattered to pixel (8ty+ky, 8tx+kx), + bias + skip, cropped, bf16 NHWC
#include <cuda.h>

#include <mutex>

#include "ptx.cuh"
#include "tc_api.cuh"

namespace tu {

namespace {

constexpr int BM = 128, BK = 64, NSTAGE = 4, NUM_THREADS = 384;   // warps 0-3: TMA, MMA, TMEM alloc, spare; warps 4-11: epilogue
constexpr int A_STAGE = BM * BK * 2;   // 16384
constexpr int UE_PITCH = 68;           // floats per staged row of the unembed epilogue (64 + 4: conflict-free 16-byte accesses)
constexpr int UE_BYTES = 8 * 32 * UE_PITCH * 4;
constexpr int UT_BYTES = 8 * 2 * 4096;     // UNEMBED_TMA: per epilogue warp two 4 KB boxes (32 pixels x 64 channels), skip in / result out

enum { EPI_STORE = 0, EPI_RESID = 1, EPI_EMBED = 2, EPI_UNEMBED = 3, EPI_UNEMBED_TMA = 4 };

struct GemmParams {
    int M, N, K, BN;
    int tiles_m, tiles_n, total_tiles;
    int amode;          // 0: plain 2D A; 1: patch-embed rank-5 A
    int epi, act;       // act: 0 none, 2 gelu (STORE only)
    const float *bias;
    // STORE
    bf16 *out;
    // RESID
    float *x;
    bf16 *x_bf16;       // optional bf16 copy of the updated stream
    // EMBED / UNEMBED geometry
    const float *pos;
    float *tok;
    int B, Ht, Wt, nWy, nWx, window, dim;
    int tiles_tx;       // embed: x-tiles of 16 tokens
    const bf16 *skip;
    int skipH, skipW, Hc, Wc;
    // UNEMBED_TMA behind the fused window stack: tiles are handed out through a global counter and an M tile is touched only
    // after the stack has published it (tile_flags[tm] != 0), so that this kernel can start -- programmatic dependent launch,
    // no griddepcontrol.wait -- on the SMs the stack's last, partly filled wave leaves idle
    int *dyn_ctr;
    const int *tile_flags;
    unsigned long long *trace;      // debug (tu_debug_trace)
    unsigned int trace_cap;
    int a_reuse;        // UNEMBED_TMA behind the stack, K = 128: the A tile (128 tokens) is loaded once per M tile into one of two buffers
                        // (the A halves of the four stage slots) and only the filter streams through the ring: a draw of four n-tiles
                        // moves 104 KB through the TMA engine instead of 128 KB per tile
    int rev;            // tiles are walked last to first (debug key "snake", bit 2: unembed behind a reversed window stack)
};
thread_local int g_unembed_areuse = 0;  // debug key "unembed_areuse"
constexpr int QD = 4;      // depth of the tile-index queue between the scheduler thread and the three roles

struct Barriers {
    uint64_t full[NSTAGE];
    uint64_t empty[NSTAGE];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
    uint64_t skip_full[8][2];   // UNEMBED_TMA: private to each epilogue warp
    uint64_t a_full[2], a_empty[2];   // a_reuse: the two A-tile buffers
    uint64_t q_full[QD], q_empty[QD];
    int tq[QD];
    uint32_t tmem_base;
};
static_assert(sizeof(Barriers) <= 512, "barrier block too large");

__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ long win_row(int b, int ty, int tx, int nWy, int nWx) {
    return (((long)b * nWy + (ty >> 3)) * nWx + (tx >> 3)) * 64 + (ty & 7) * 8 + (tx & 7);
}

// GELU = x * 0.5 (1 + tanh(p(x))), p an odd polynomial fitted to atanh(erf(x / sqrt 2)): |error| <= 2.6e-5 against the
// exact erf form, far below the bf16 resolution of the stored activation (tools/fit_gelu.py)
__device__ __forceinline__ float gelu_f(float x) {
    const float x2 = fminf(x * x, 64.f);
    float q = fmaf(-0.0003515167826820022f, x2, 0.03700564597780192f);
    q = fmaf(q, x2, 0.7975078843613885f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * q));
    const float h = 0.5f * x;
    return fmaf(h, t, h);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
               const __grid_constant__ CUtensorMap tmap_skip, const __grid_constant__ CUtensorMap tmap_o5, const GemmParams p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const int w_stage = p.BN * 128;
    const int stage_bytes = A_STAGE + w_stage;
    uint8_t *smem_al = smem_raw + (smem0 - ptx::smem_u32(smem_raw));
    // UNEMBED_TMA keeps its 1024-byte aligned pixel boxes right after the operand stages; barriers follow
    const int bar_off = NSTAGE * stage_bytes + (p.epi == EPI_UNEMBED_TMA ? UT_BYTES : 0);
    Barriers *bars = reinterpret_cast<Barriers *>(smem_al + bar_off);
    float *stage_f32 = reinterpret_cast<float *>(smem_al + bar_off + 512);   // UNEMBED only: 8 warps x 32 rows x 68 floats
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
    const int nk = p.K / BK;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            ptx::mbar_init(ptx::smem_u32(&bars->full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->empty[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(ptx::smem_u32(&bars->acc_full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->acc_empty[i]), 8);
        }
        for (int i = 0; i < 16; ++i) ptx::mbar_init(ptx::smem_u32(&bars->skip_full[i >> 1][i & 1]), 1);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(ptx::smem_u32(&bars->a_full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->a_empty[i]), 1);
        }
        for (int i = 0; i < QD; ++i) {
            ptx::mbar_init(ptx::smem_u32(&bars->q_full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->q_empty[i]), 10);       // TMA producer, MMA warp, 8 epilogue warps
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), 512);
        ptx::tmem_relinquish();
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_w);
        if (p.epi == EPI_UNEMBED_TMA) { ptx::prefetch_tmap(&tmap_skip); ptx::prefetch_tmap(&tmap_o5); }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const bool dyn = p.dyn_ctr != nullptr;
    if (!dyn) pdl_wait();       // dynamic mode orders itself behind the window stack tile by tile (tile_flags)

    // k-th tile of this CTA: static round-robin, or the k-th index the scheduler thread drew from the global counter
    auto tile_at = [&](int k) -> int {
        if (!dyn) {
            const int t = blockIdx.x + k * gridDim.x;
            return t < p.total_tiles ? (p.rev ? p.total_tiles - 1 - t : t) : -1;
        }
        ptx::mbar_wait(ptx::smem_u32(&bars->q_full[k % QD]), (k / QD) & 1);
        return *reinterpret_cast<volatile int *>(&bars->tq[k % QD]);
    };
    auto tile_done = [&](int k) {        // one arrival per consumer (the calling thread)
        if (dyn) ptx::mbar_arrive(ptx::smem_u32(&bars->q_empty[k % QD]));
    };
    auto wait_published = [&](int tm) {  // the window stack has written (and fenced) the tokens of M tile tm
        if (!p.tile_flags) return;
        const long long t0 = clock64();
        while (ld_acquire_gpu(p.tile_flags + tm) == 0) {
            __nanosleep(200);
            if (clock64() - t0 > 4000000000LL) __trap();      // ~2 s: a protocol bug traps instead of hanging the GPU
        }
        asm volatile("fence.proxy.async.global;" ::: "memory");      // generic-proxy acquire -> the TMA reads that follow
    };

    // the scheduler thread has seen the flag of a tile before it hands the tile out; a role that is about to read the tile through the
    // async proxy (TMA) only needs the proxy fence (generic-proxy acquire, passed on through the queue's mbarrier -> TMA reads)
    auto proxy_fence = [&]() {
        if (p.tile_flags) asm volatile("fence.proxy.async.global;" ::: "memory");
    };

    if (warp == 3 && lane == 0 && dyn) {
        // ================================ tile scheduler ================================
        // Tiles are drawn DRAW at a time (one atomic and one flag check per draw: the DRAW n-tiles of a draw belong to one M tile, whose
        // tokens then also stay in L2 for the draw), and the wait for the window stack's flag happens HERE, up to QD tiles ahead of
        // the roles: a global acquire load is ~0.8 us, which the TMA producer and the eight skip-box prefetches used to pay per tile
        constexpr int DRAW = 4;
        int last_tm = -1;
        for (int k = 0;;) {
            const int base = (p.tiles_n % DRAW == 0) ? atomicAdd(p.dyn_ctr, DRAW) : atomicAdd(p.dyn_ctr, 1);
            const int cnt = (p.tiles_n % DRAW == 0) ? DRAW : 1;
            bool last = false;
            for (int j = 0; j < cnt && !last; ++j, ++k) {
                const int t0 = base + j;
                last = t0 >= p.total_tiles;
                const int t = last ? -1 : (p.rev ? p.total_tiles - 1 - t0 : t0);
                if (!last) {
                    trace_event(p.trace, p.trace_cap, 3, (unsigned)t);          // tile drawn by this CTA's scheduler
                    const int tm = t / p.tiles_n;
                    if (tm != last_tm) { wait_published(tm); last_tm = tm; trace_event(p.trace, p.trace_cap, 5, (unsigned)tm); }
                }
                if (k >= QD) ptx::mbar_wait(ptx::smem_u32(&bars->q_empty[k % QD]), ((k / QD) & 1) ^ 1);
                *reinterpret_cast<volatile int *>(&bars->tq[k % QD]) = t;
                ptx::mbar_arrive(ptx::smem_u32(&bars->q_full[k % QD]));      // release: the index is visible to whoever passes the wait
            }
            if (last) break;
        }
    } else if (warp == 0 && lane == 0) {
        // ================================ TMA producer ================================
        int stage = 0;
        uint32_t phase = 0;
        int ready_tm = -1, na = 0;
        for (int k = 0;; ++k) {
            const int t = tile_at(k);
            if (t < 0) break;
            const int tn = t % p.tiles_n, tm = t / p.tiles_n;
            const int n0 = tn * p.BN;
            if (tm != ready_tm) {
                if (dyn) proxy_fence(); else wait_published(tm);
                ready_tm = tm;
                if (p.a_reuse) {          // the M tile's tokens -> A buffer na & 1 (free once the MMAs of the M tile before last have retired)
                    const int ab = na & 1;
                    ptx::mbar_wait(ptx::smem_u32(&bars->a_empty[ab]), ((na >> 1) & 1) ^ 1);
                    const uint32_t fa = ptx::smem_u32(&bars->a_full[ab]);
                    ptx::mbar_expect_tx(fa, nk * A_STAGE);
                    for (int s = 0; s < nk; ++s) ptx::tma_load_2d(smem0 + (2 * ab + s) * stage_bytes, &tmap_a, fa, s * BK, tm * BM);
                    ++na;
                }
            }
            if (p.a_reuse) {
                for (int s = 0; s < nk; ++s) {
                    ptx::mbar_wait(ptx::smem_u32(&bars->empty[stage]), phase ^ 1);
                    const uint32_t fb = ptx::smem_u32(&bars->full[stage]);
                    ptx::mbar_expect_tx(fb, w_stage);
                    ptx::tma_load_2d(smem0 + stage * stage_bytes + A_STAGE, &tmap_w, fb, s * BK, n0);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
                tile_done(k);
                continue;
            }
            for (int s = 0; s < nk; ++s) {
                ptx::mbar_wait(ptx::smem_u32(&bars->empty[stage]), phase ^ 1);
                const uint32_t dst = smem0 + stage * stage_bytes;
                const uint32_t fb = ptx::smem_u32(&bars->full[stage]);
                ptx::mbar_expect_tx(fb, stage_bytes);
                if (p.amode == 0) {
                    ptx::tma_load_2d(dst, &tmap_a, fb, s * BK, tm * BM);
                } else {
                    const int tx0 = (tm % p.tiles_tx) * 16, r0 = (tm / p.tiles_tx) * 8;
                    ptx::tma_load_5d(dst, &tmap_a, fb, 0, s & 7, tx0, s >> 3, r0);
                }
                ptx::tma_load_2d(dst + A_STAGE, &tmap_w, fb, s * BK, n0);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
            tile_done(k);
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (whole warp converged, elected lane issues) ================================
        const uint32_t leader = ptx::elect_one();
        const uint32_t idesc = ptx::make_idesc_bf16(BM, p.BN);
        const uint32_t smem_lo = ptx::sdesc_lo(smem0);
        int stage = 0;
        uint32_t phase = 0;
        int cur_tm = -1, na = 0, ab = 0;
        for (int it = 0;; ++it) {
            const int t = tile_at(it);
            if (t < 0) break;
            __syncwarp();
            if (lane == 0) tile_done(it);
            if (p.a_reuse && t / p.tiles_n != cur_tm) {
                // every MMA that reads the previous M tile's A buffer has been issued: it is free once they retire
                if (cur_tm >= 0) ptx::umma_commit_pred(ptx::smem_u32(&bars->a_empty[ab]), leader);
                ab = na & 1;
                ptx::mbar_wait(ptx::smem_u32(&bars->a_full[ab]), (na >> 1) & 1);
                ++na;
                cur_tm = t / p.tiles_n;
            }
            const int set = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            ptx::mbar_wait(ptx::smem_u32(&bars->acc_empty[set]), aphase ^ 1);
            ptx::tc_fence_after();
            const uint32_t acc = tmem_base + set * 256;
            for (int s = 0; s < nk; ++s) {
                ptx::mbar_wait(ptx::smem_u32(&bars->full[stage]), phase);
                ptx::tc_fence_after();
                const uint32_t w_lo = smem_lo + ((stage * stage_bytes + A_STAGE) >> 4);
                const uint32_t a_lo = p.a_reuse ? smem_lo + (((2 * ab + s) * stage_bytes) >> 4) : smem_lo + ((stage * stage_bytes) >> 4);
                ptx::umma_bf16_lo_rt(acc, a_lo, w_lo, idesc, s != 0, leader);
#pragma unroll
                for (int k4 = 1; k4 < 4; ++k4) ptx::umma_bf16_lo<1>(acc, a_lo + k4 * 2, w_lo + k4 * 2, idesc, leader);
                ptx::umma_commit_pred(ptx::smem_u32(&bars->empty[stage]), leader);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
            ptx::umma_commit_pred(ptx::smem_u32(&bars->acc_full[set]), leader);
        }
    } else if (warp >= 4) {
        // ================================ epilogue ================================
        // 8 warps: warp w drains TMEM lane quadrant q = w % 4 (a warp may only touch its own 32 lanes), columns of half w / 4
        const int q = (warp - 4) & 3, half = (warp - 4) >> 2;
        const int cbeg = half * (p.BN >> 1), cend = cbeg + (p.BN >> 1);
        const int i = q * 32 + lane;                  // row of the tile owned by this thread
        if (p.epi == EPI_UNEMBED_TMA) {
            // Window-ordered tokens, BN = 128 (two pixels of every token), no crop inside a patch.  This warp's 32 token rows are
            // half a window: 4 token rows x 8 token columns, and it owns pixel (dy, dx) of each.  Those 32 pixels are a strided
            // box (c 64, dx 1, tx 8, y 4 rows at stride 8, b 1) of the NHWC tensor: the skip connection arrives by TMA (the box of
            // the NEXT tile is requested before this tile is processed), the sum is written over it in shared memory and
            // leaves by a TMA store.  Rows / columns outside the image are clipped by the TMA engine in both directions, so pad
            // tokens need no special case and no thread computes a global address.
            const int w8 = warp - 4;
            uint8_t *box = smem_al + NSTAGE * stage_bytes + w8 * 8192;
            const uint32_t box_sm = smem0 + NSTAGE * stage_bytes + w8 * 8192;
            auto coords = [&](int t, int &dx, int &tx0, int &y0, int &b) -> bool {
                const int tn = t % p.tiles_n, tm = t / p.tiles_n;
                const int wi = tm * 2 + (q >> 1);                 // window of this warp's rows
                const int wx = wi % p.nWx, rest = wi / p.nWx;
                const int wy = rest % p.nWy;
                b = rest / p.nWy;
                const int pix = ((tn * p.BN) >> 6) + half;
                dx = pix & 7;
                tx0 = wx * 8;
                y0 = (wy * 8 + (q & 1) * 4) * 8 + (pix >> 3);
                return b < p.B;
            };
            auto request = [&](int t, int buf) {                  // lane 0: skip box of tile t -> buffer buf
                int dx, tx0, y0, b;
                if (!coords(t, dx, tx0, y0, b)) return;
                proxy_fence();                                    // (dynamic mode) the scheduler handed the tile out after the stack published it
                const uint32_t fb = ptx::smem_u32(&bars->skip_full[w8][buf]);
                ptx::mbar_expect_tx(fb, 4096);
                ptx::tma_load_5d(box_sm + buf * 4096, &tmap_skip, fb, 0, dx, tx0, y0, b);
            };
            const float4 *bias4 = reinterpret_cast<const float4 *>(p.bias);
            int t = tile_at(0);
            if (lane == 0 && t >= 0) request(t, 0);
            uint32_t sph[2] = {0u, 0u};
            for (int it = 0; t >= 0; ++it) {
                const int set = it & 1, buf = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                int dx, tx0, y0, b;
                const bool live = coords(t, dx, tx0, y0, b);
                const int t_next = tile_at(it + 1);
                __syncwarp();
                if (lane == 0) {
                    tile_done(it);
                    if (t_next >= 0) request(t_next, buf ^ 1);   // that buffer's store has been read (below)
                }
                t = t_next;                                       // (everything below uses the coordinates computed above)
                ptx::mbar_wait(ptx::smem_u32(&bars->acc_full[set]), aphase);
                ptx::tc_fence_after();
                const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + set * 256 + cbeg;
                uint32_t v0[32], v1[32];
                ptx::tmem_ld_x32(tbase, v0);
                ptx::tmem_ld_x32(tbase + 32, v1);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->acc_empty[set]));
                if (!live) continue;                              // warp-uniform
                ptx::mbar_wait(ptx::smem_u32(&bars->skip_full[w8][buf]), sph[buf]);
                sph[buf] ^= 1;
                uint8_t *rowp = box + buf * 4096 + lane * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint4 *sp = reinterpret_cast<uint4 *>(rowp + ((c ^ (lane & 7)) << 4));
                    const uint4 su = *sp;
                    const __nv_bfloat162 *sh = reinterpret_cast<const __nv_bfloat162 *>(&su);
                    const float2 s0 = __bfloat1622float2(sh[0]), s1 = __bfloat1622float2(sh[1]), s2 = __bfloat1622float2(sh[2]), s3 = __bfloat1622float2(sh[3]);
                    const float4 ba = __ldg(bias4 + c * 2), bb = __ldg(bias4 + c * 2 + 1);
                    const uint32_t *v = c < 4 ? &v0[c * 8] : &v1[(c - 4) * 8];
                    uint4 u;
                    __nv_bfloat162 h;
                    h = __floats2bfloat162_rn(__uint_as_float(v[0]) + ba.x + s0.x, __uint_as_float(v[1]) + ba.y + s0.y); u.x = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(__uint_as_float(v[2]) + ba.z + s1.x, __uint_as_float(v[3]) + ba.w + s1.y); u.y = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(__uint_as_float(v[4]) + bb.x + s2.x, __uint_as_float(v[5]) + bb.y + s2.y); u.z = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(__uint_as_float(v[6]) + bb.z + s3.x, __uint_as_float(v[7]) + bb.w + s3.y); u.w = *reinterpret_cast<uint32_t *>(&h);
                    *sp = u;
                }
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_5d(&tmap_o5, box_sm + buf * 4096, 0, dx, tx0, y0, b);
                    ptx::bulk_commit();
                    ptx::bulk_wait_read<0>();                     // the box may be refilled by the request two tiles ahead
                }
                __syncwarp();
            }
            if (lane == 0) ptx::bulk_wait<0>();
            __syncwarp();
        } else {
        int it = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
            const int tn = t % p.tiles_n, tm = t / p.tiles_n;
            const int n0 = tn * p.BN;
            const int set = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            // ---- row geometry
            bool valid;
            long row = 0;            // STORE/RESID: m ; EMBED: token stream row
            int b = 0, ty = 0, tx = 0;
            if (p.epi == EPI_EMBED) {
                tx = (tm % p.tiles_tx) * 16 + (i & 15);
                const int bty = (tm / p.tiles_tx) * 8 + (i >> 4);
                b = bty / p.Ht;
                ty = bty - b * p.Ht;
                valid = tx < p.Wt && b < p.B;
                row = p.window ? win_row(b, ty, tx, p.nWy, p.nWx) : ((long)b * p.Ht + ty) * p.Wt + tx;
            } else {
                const int m = tm * BM + i;
                valid = m < p.M;
                row = m;
                if (p.epi == EPI_UNEMBED) {
                    if (p.window) {
                        const int wi = m >> 6, tk = m & 63;
                        const int wx = wi % p.nWx, rest = wi / p.nWx;
                        const int wy = rest % p.nWy;
                        b = rest / p.nWy;
                        ty = wy * 8 + (tk >> 3);
                        tx = wx * 8 + (tk & 7);
                    } else {
                        tx = m % p.Wt;
                        const int rest = m / p.Wt;
                        ty = rest % p.Ht;
                        b = rest / p.Ht;
                    }
                    valid = valid && ty < p.Ht && tx < p.Wt;
                }
            }
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + set * 256;
            if (p.epi == EPI_UNEMBED) {
                // BN = 128: this warp owns 64 columns = the 64 channels of ONE output pixel (dy,dx) of each of its 32 tokens.
                // A thread owning a whole pixel would store it as eight 16-byte pieces, 32 different cache lines per warp
                // instruction; instead the warp transposes through shared memory so that 8 lanes cover one pixel (128
                // contiguous bytes) and an instruction touches 4 lines.  The skip-connection loads do not depend on the
                // accumulator, so all 8 are issued before waiting for the MMAs (one DRAM latency per tile, overlapped).
                {   // L2 prefetch of the skip pixel this thread's token row needs in the CTA's NEXT tile (every 128-byte line of
                    // the skip tensor is used exactly once, so this only moves its DRAM fetch one tile earlier)
                    const int t2 = t + gridDim.x;
                    if (t2 < p.total_tiles) {
                        const int m2 = (t2 / p.tiles_n) * BM + i;
                        int b2, ty2, tx2;
                        if (p.window) {
                            const int wi = m2 >> 6, tk = m2 & 63;
                            const int wx = wi % p.nWx, rest = wi / p.nWx;
                            b2 = rest / p.nWy;
                            ty2 = (rest % p.nWy) * 8 + (tk >> 3);
                            tx2 = wx * 8 + (tk & 7);
                        } else {
                            tx2 = m2 % p.Wt;
                            const int rest = m2 / p.Wt;
                            ty2 = rest % p.Ht;
                            b2 = rest / p.Ht;
                        }
                        const int pix2 = ((t2 % p.tiles_n) * p.BN + cbeg) >> 6;
                        const int gy2 = ty2 * 8 + (pix2 >> 3), gx2 = tx2 * 8 + (pix2 & 7);
                        if (m2 < p.M && ty2 < p.Ht && tx2 < p.Wt && gy2 < p.Hc && gx2 < p.Wc) {
                            const bf16 *a = p.skip + (((long)b2 * p.skipH + gy2) * p.skipW + gx2) * 64;
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
                        }
                    }
                }
                float *stg = stage_f32 + (warp - 4) * (32 * UE_PITCH);
                const uint32_t geo = valid ? ((uint32_t)b << 20) | ((uint32_t)ty << 10) | (uint32_t)tx : 0xFFFFFFFFu;
                const int sub = lane >> 3, chunk = lane & 7;
                const float4 bs0 = *reinterpret_cast<const float4 *>(p.bias + chunk * 8), bs1 = *reinterpret_cast<const float4 *>(p.bias + chunk * 8 + 4);
                const int pix = (n0 + cbeg) >> 6;
                const int dy = pix >> 3, dx = pix & 7;
                bool ok[8];
                long oo[8];
                uint4 su[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {      // row j*4 + sub of the warp's 32 token rows
                    const uint32_t g = __shfl_sync(0xffffffffu, geo, j * 4 + sub);
                    const int gb = g >> 20, gy = ((g >> 10) & 1023) * 8 + dy, gx = (g & 1023) * 8 + dx;
                    ok[j] = g != 0xFFFFFFFFu && gy < p.Hc && gx < p.Wc;
                    oo[j] = (((long)gb * p.Hc + gy) * p.Wc + gx) * 64 + chunk * 8;
                    const long so = ok[j] ? (((long)gb * p.skipH + gy) * p.skipW + gx) * 64 + chunk * 8 : (long)chunk * 8;
                    su[j] = *reinterpret_cast<const uint4 *>(p.skip + so);
                }
                ptx::mbar_wait(ptx::smem_u32(&bars->acc_full[set]), aphase);
                ptx::tc_fence_after();
                {
                    uint32_t v0[32], v1[32];
                    ptx::tmem_ld_x32(tbase + cbeg, v0);
                    ptx::tmem_ld_x32(tbase + cbeg + 32, v1);
                    ptx::tmem_ld_wait();
                    float *row = stg + lane * UE_PITCH;
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        *reinterpret_cast<float4 *>(row + c) = make_float4(__uint_as_float(v0[c]), __uint_as_float(v0[c + 1]), __uint_as_float(v0[c + 2]), __uint_as_float(v0[c + 3]));
                        *reinterpret_cast<float4 *>(row + 32 + c) = make_float4(__uint_as_float(v1[c]), __uint_as_float(v1[c + 1]), __uint_as_float(v1[c + 2]), __uint_as_float(v1[c + 3]));
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->acc_empty[set]));      // accumulator columns are in shared memory now
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float *sr = stg + (j * 4 + sub) * UE_PITCH + chunk * 8;
                    const float4 a0 = *reinterpret_cast<const float4 *>(sr), a1 = *reinterpret_cast<const float4 *>(sr + 4);
                    const __nv_bfloat162 *sh = reinterpret_cast<const __nv_bfloat162 *>(&su[j]);
                    const float2 s0 = __bfloat1622float2(sh[0]), s1 = __bfloat1622float2(sh[1]), s2 = __bfloat1622float2(sh[2]), s3 = __bfloat1622float2(sh[3]);
                    uint4 u;
                    __nv_bfloat162 h;
                    h = __floats2bfloat162_rn(a0.x + bs0.x + s0.x, a0.y + bs0.y + s0.y); u.x = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(a0.z + bs0.z + s1.x, a0.w + bs0.w + s1.y); u.y = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(a1.x + bs1.x + s2.x, a1.y + bs1.y + s2.y); u.z = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(a1.z + bs1.z + s3.x, a1.w + bs1.w + s3.y); u.w = *reinterpret_cast<uint32_t *>(&h);
                    if (ok[j]) *reinterpret_cast<uint4 *>(p.out + oo[j]) = u;
                }
                __syncwarp();                      // staging rows are free for the next tile
                continue;
            }
            ptx::mbar_wait(ptx::smem_u32(&bars->acc_full[set]), aphase);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int c0 = cbeg; c0 < cend; c0 += 32) {
                uint32_t v[32];
                ptx::tmem_ld_x32(tbase + c0, v);
                ptx::tmem_ld_wait();
                if (!valid) continue;
                const int n = n0 + c0;
                if (p.epi == EPI_STORE) {
                    bf16 *o = p.out + row * p.N + n;
#pragma unroll
                    for (int c = 0; c < 32; c += 8) {
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float a = __uint_as_float(v[c + e]) + __ldg(p.bias + n + c + e);
                            f[e] = p.act == 2 ? gelu_f(a) : a;
                        }
                        uint4 u;
                        __nv_bfloat162 h;
                        h = __floats2bfloat162_rn(f[0], f[1]); u.x = *reinterpret_cast<uint32_t *>(&h);
                        h = __floats2bfloat162_rn(f[2], f[3]); u.y = *reinterpret_cast<uint32_t *>(&h);
                        h = __floats2bfloat162_rn(f[4], f[5]); u.z = *reinterpret_cast<uint32_t *>(&h);
                        h = __floats2bfloat162_rn(f[6], f[7]); u.w = *reinterpret_cast<uint32_t *>(&h);
                        *reinterpret_cast<uint4 *>(o + c) = u;
                    }
                } else if (p.epi == EPI_RESID) {
                    float *xr = p.x + row * p.N + n;
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        float4 o = *reinterpret_cast<const float4 *>(xr + c);
                        o.x += __uint_as_float(v[c + 0]) + __ldg(p.bias + n + c + 0);
                        o.y += __uint_as_float(v[c + 1]) + __ldg(p.bias + n + c + 1);
                        o.z += __uint_as_float(v[c + 2]) + __ldg(p.bias + n + c + 2);
                        o.w += __uint_as_float(v[c + 3]) + __ldg(p.bias + n + c + 3);
                        *reinterpret_cast<float4 *>(xr + c) = o;
                        if (p.x_bf16) {
                            __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
                            uint2 u;
                            u.x = *reinterpret_cast<uint32_t *>(&h0);
                            u.y = *reinterpret_cast<uint32_t *>(&h1);
                            *reinterpret_cast<uint2 *>(p.x_bf16 + row * p.N + n + c) = u;
                        }
                    }
                } else if (p.epi == EPI_EMBED) {
                    float *o = p.tok + row * p.dim + n;
                    const float *pe = p.pos ? p.pos + ((long)ty * p.Wt + tx) * p.dim + n : nullptr;
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
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->acc_empty[set]));
        }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

PerDeviceMax g_smem_set;

int pick_bn(int N) {
    if (N % 256 == 0) return 256;
    if (N % 192 == 0) return 192;
    if (N % 128 == 0) return 128;
    return 0;
}

int encode_2d(CUtensorMap *tm, const void *ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    TcEncodeFn enc = tc_encode_fn();
    cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows}, es[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("tu: cuTensorMapEncodeTiled(gemm 2d) failed with code " + std::to_string((int)r));
        return TU_ERR_CUDA;
    }
    return TU_OK;
}

int launch(const CUtensorMap &ta, const CUtensorMap &tw, GemmParams &p, cudaStream_t st, const CUtensorMap *tskip = nullptr,
           const CUtensorMap *to5 = nullptr) {
    const int g_sm_count = device_sm_count();
    const int smem = NSTAGE * (A_STAGE + p.BN * 128) + 512 + (p.epi == EPI_UNEMBED ? UE_BYTES : p.epi == EPI_UNEMBED_TMA ? UT_BYTES : 0) + 1024;
    if (smem > g_smem_set.get()) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return cuda_fail(e, "gemm_tc smem attribute");
        g_smem_set.set(smem);
    }
    p.tiles_n = p.N / p.BN;
    p.total_tiles = p.tiles_m * p.tiles_n;
    const int grid = p.total_tiles < g_sm_count ? p.total_tiles : g_sm_count;
    launch_pdl(gemm_tc_kernel, dim3(grid), dim3(NUM_THREADS), (size_t)smem, st, ta, tw, tskip ? *tskip : ta, to5 ? *to5 : ta, p);
    TU_CHECK_LAUNCH("gemm_tc");
    return TU_OK;
}

}  // namespace

void tc_set_unembed_areuse(int on) { g_unembed_areuse = on; }

int tc_linear(const bf16 *A, const bf16 *W, const float *bias, int M, int N, int K, int act, bf16 *out, float *resid_x,
              bf16 *resid_bf16, cudaStream_t st) {
    const int BN = pick_bn(N);
    if (!tc_encode_fn() || !BN || K % 64 || (reinterpret_cast<uintptr_t>(A) & 127) || (reinterpret_cast<uintptr_t>(W) & 127))
        return TU_TC_UNSUPPORTED;
    CUtensorMap ta, tw;
    int rc;
    if ((rc = encode_2d(&ta, A, M, K, BM))) return rc;
    if ((rc = encode_2d(&tw, W, N, K, BN))) return rc;
    GemmParams p = {};
    p.M = M; p.N = N; p.K = K; p.BN = BN;
    p.tiles_m = ceil_div(M, BM);
    p.amode = 0;
    p.bias = bias;
    if (resid_x) {
        p.epi = EPI_RESID; p.x = resid_x; p.x_bf16 = resid_bf16;
    } else {
        p.epi = EPI_STORE; p.act = act; p.out = out;
    }
    return launch(ta, tw, p, st);
}

int tc_patch_embed(const bf16 *feat, const bf16 *W, const float *bias, const float *pos, float *tok, int B, int H, int Wd,
                   int Ht, int Wt, int dim, int window, cudaStream_t st) {
    const int BN = pick_bn(dim);
    if (!tc_encode_fn() || !BN || H != 8 * Ht || Wd < 8 * Wt || (reinterpret_cast<uintptr_t>(feat) & 127) ||
        (reinterpret_cast<uintptr_t>(W) & 127))
        return TU_TC_UNSUPPORTED;
    CUtensorMap ta, tw;
    {
        // rank-5 view of NHWC(64): (c, kx, tx, ky, b*ty)
        cuuint64_t dims[5] = {64, 8, (cuuint64_t)Wt, 8, (cuuint64_t)B * Ht};
        cuuint64_t strides[4] = {128, 1024, (cuuint64_t)Wd * 128, (cuuint64_t)Wd * 128 * 8};
        cuuint32_t box[5] = {64, 1, 16, 1, 8}, es[5] = {1, 1, 1, 1, 1};
        CUresult r = tc_encode_fn()(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void *)feat, dims, strides, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("tu: cuTensorMapEncodeTiled(patch embed) failed with code " + std::to_string((int)r));
            return TU_ERR_CUDA;
        }
    }
    int rc;
    if ((rc = encode_2d(&tw, W, dim, 4096, BN))) return rc;
    GemmParams p = {};
    p.M = B * Ht * Wt; p.N = dim; p.K = 4096; p.BN = BN;
    p.tiles_tx = ceil_div(Wt, 16);
    p.tiles_m = p.tiles_tx * ceil_div(B * Ht, 8);
    p.amode = 1;
    p.epi = EPI_EMBED;
    p.bias = bias; p.pos = pos; p.tok = tok;
    p.B = B; p.Ht = Ht; p.Wt = Wt; p.nWy = (Ht + 7) / 8; p.nWx = (Wt + 7) / 8; p.window = window; p.dim = dim;
    return launch(ta, tw, p, st);
}

// dyn_ctr / tile_flags (both or neither; zeroed by the caller earlier in the stream): the launch follows the fused window stack,
// which publishes every 128-token tile in tile_flags; used only with the strided-box TMA epilogue
int tc_patch_unembed(const bf16 *tok_bf16, const bf16 *W, const float *bias, const bf16 *skip, int skipH, int skipW, bf16 *out,
                     int B, int Ht, int Wt, int Hc, int Wc, int dim, int window, int *dyn_ctr, const int *tile_flags, cudaStream_t st) {
    if (!tc_encode_fn() || dim % 64 || (reinterpret_cast<uintptr_t>(tok_bf16) & 127) || (reinterpret_cast<uintptr_t>(W) & 127) ||
        (reinterpret_cast<uintptr_t>(skip) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || B >= 2048 || Ht >= 1024 || Wt >= 1024)
        return TU_TC_UNSUPPORTED;
    const int nWy = (Ht + 7) / 8, nWx = (Wt + 7) / 8;
    const int M = window ? B * nWy * nWx * 64 : B * Ht * Wt;
    CUtensorMap ta, tw;
    int rc;
    if ((rc = encode_2d(&ta, tok_bf16, M, dim, BM))) return rc;
    if ((rc = encode_2d(&tw, W, 4096, dim, 128))) return rc;
    GemmParams p = {};
    p.M = M; p.N = 4096; p.K = dim; p.BN = 128;      // 2 pixels per token and tile: leaves shared memory for the epilogue transpose
    p.tiles_m = ceil_div(M, BM);
    p.amode = 0;
    p.epi = EPI_UNEMBED;
    p.bias = bias; p.out = out; p.skip = skip; p.skipH = skipH; p.skipW = skipW; p.Hc = Hc; p.Wc = Wc;
    p.B = B; p.Ht = Ht; p.Wt = Wt; p.nWy = nWy; p.nWx = nWx; p.window = window; p.dim = dim;
    // strided-box TMA epilogue when a patch is never cropped horizontally: (c 64, dx 8, tx, y, b) views of skip and out
    if (window && Wc == 8 * Wt && (reinterpret_cast<uintptr_t>(skip) & 127) == 0 && (reinterpret_cast<uintptr_t>(out) & 127) == 0) {
        CUtensorMap ts, to;
        cuuint64_t sd[5] = {64, 8, (cuuint64_t)Wt, (cuuint64_t)Hc, (cuuint64_t)B};
        cuuint64_t ss[4] = {128, 1024, (cuuint64_t)skipW * 128, (cuuint64_t)skipH * skipW * 128};
        cuuint64_t os[4] = {128, 1024, (cuuint64_t)Wc * 128, (cuuint64_t)Hc * Wc * 128};
        cuuint32_t box[5] = {64, 1, 8, 32, 1}, es[5] = {1, 1, 1, 8, 1};
        TcEncodeFn enc = tc_encode_fn();
        const bool ok = enc(&ts, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void *)skip, sd, ss, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS &&
                        enc(&to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void *)out, sd, os, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        if (ok) {
            p.epi = EPI_UNEMBED_TMA;
            p.rev = ((g_snake_mask >> 2) & 1) && dim == 128;      // the dim-128 stack publishes its tiles in the same order
            if (dyn_ctr && tile_flags) {
                p.dyn_ctr = dyn_ctr; p.tile_flags = tile_flags; p.trace = g_trace_buf; p.trace_cap = g_trace_cap;
                p.a_reuse = (g_unembed_areuse && dim == 2 * BK) ? 1 : 0;
            }
            return launch(ta, tw, p, st, &ts, &to);
        }
    }
    if (tile_flags) return TU_TC_UNSUPPORTED;      // the caller relaunches without the overlap (this path waits for the whole stack)
    return launch(ta, tw, p, st);
}

}  // namespace tu
