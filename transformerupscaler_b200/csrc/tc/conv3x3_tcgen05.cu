// 3x3 / pad-1 convolution with 64 input channels as an implicit GEMM on the 5th-gen tensor cores.
//
// Reference call sites: conv2 / downsample / decoder_conv1 (WindowTransformer/model.py:202,205,221;
// FastTransformer/model.py:204,228; ResidualTransformer/model.py:85,88,111) and the Upsampler convs
// 64 -> 64 r^2 followed by PixelShuffle(r) (FastTransformer/utils.py:49-98) — 85-95 % of every model's FLOPs.
//
// Mapping (per CTA, persistent over tiles):
//   * NHWC bf16 activations: one pixel = 64 channels = 128 bytes = exactly one row of a 128-byte-swizzled
//     K-major UMMA operand.  A tile is TILE_R output rows x 128 output pixels.
//   * TMA brings each needed input row segment (136 pixels: 128 + halo, zero-filled outside the image, so
//     padding costs nothing) into a ring of shared-memory slots ONCE; the nine filter taps are nine UMMA
//     descriptors into the same bytes: tap (ky,kx) of output row r reads slot row r+ky starting kx pixels
//     (= kx*128 bytes) in.  Stride 2 views the image as (W/2) "super-pixels" of 128 channels and loads the even
//     and the odd pixels of a row as two dense slots.
//   * the 9 x (64 co x 64 ci) filter bank of the current 64-channel output chunk stays resident in shared
//     memory (72 KB); accumulators live in TMEM: 2 sets x TILE_R rows x 64 fp32 columns = all 512 columns, so
//     the epilogue of tile t overlaps the MMAs of tile t+1.
//   * warp 0: TMA producer, warp 1: MMA issuer (one thread), warps 4-7: epilogue (tcgen05.ld -> +bias, ReLU,
//     bf16 -> 128-byte global stores, optionally at PixelShuffle-ed addresses).
#include <cuda.h>

#include <string.h>

#include <mutex>

#include "ptx.cuh"
#include "tc_api.cuh"

namespace tu {

namespace {

constexpr int TILE_R = 4;                 // output rows per tile
constexpr int TILE_M = 128;               // output pixels per row segment = UMMA M
constexpr int BOXW = 136;                 // pixels per TMA box (128 + halo, keeps slots 1024-byte aligned)
constexpr int UNIT_BYTES = BOXW * 128;    // 17408 = 17 * 1024
constexpr int RING_UNITS = 6;
constexpr int W_BYTES_MAX = 9 * 64 * 128;  // 73728 (64 output channels); 9*16*128 for the 64->3 head (3 padded to 16)
constexpr int STG_BYTES = 4 * 2 * 4096;     // epilogue staging: 4 warps x 2 buffers x (32 pixels x 128 B), source of the TMA stores
constexpr int SMEM_BYTES = W_BYTES_MAX + RING_UNITS * UNIT_BYTES + STG_BYTES + 256 + 1024;
constexpr int NUM_THREADS = 256;

struct ConvParams {
    int B, H, W, Ho, Wo;
    int stride, relu, nchunk, ps_r;
    int tiles_x, tiles_y, tiles_per_chunk, total_tiles;
    int rev;            // tiles of a chunk are walked last to first: the tail of the producer kernel's output is still in L2
    const float *bias;
    bf16 *out;
    float *out3;        // NOUT == 16: planar fp32 (B,3,Ho,Wo)
};

struct Barriers {
    uint64_t full[RING_UNITS];
    uint64_t empty[RING_UNITS];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
    uint64_t w_full;
    uint64_t w_free;
    uint32_t tmem_base;
    float xchg[64];          // NOUT == 16 epilogue: seam values between the four epilogue warps, two row parities
};

template <int NOUT, int S>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CUtensorMap tmap_out, const ConvParams p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    // NOUT == 64: nine (co 64 x ci 64) filter taps.  NOUT == 16 (the 64 -> 3 heads): three ky blocks of 16 rows n = kx*4 + co, so
    // the tensor core sums over ky and ci on the UNSHIFTED pixel rows and the epilogue adds the three kx-shifted columns
    // (3 x fewer MMAs than nine N = 16 taps, whose cost is the 4 KB shared-memory read of the A operand, not the math).
    constexpr int NTAPS = NOUT == 64 ? 9 : 3;
    constexpr int W_TAP = NOUT * 128, W_BYTES = NTAPS * W_TAP;
    constexpr int TILE_W = NOUT == 64 ? TILE_M : TILE_M - 2;      // output pixels per tile (the head needs a 1-pixel halo of rows)
    const uint32_t w_sm = smem0;
    const uint32_t ring_sm = smem0 + W_BYTES_MAX;
    const uint32_t stg_sm = ring_sm + RING_UNITS * UNIT_BYTES;      // 1024-byte aligned (UNIT_BYTES = 17 * 1024)
    uint8_t *smem_al = smem_raw + (smem0 - ptx::smem_u32(smem_raw));
    Barriers *bars = reinterpret_cast<Barriers *>(smem_al + W_BYTES_MAX + RING_UNITS * UNIT_BYTES + STG_BYTES);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
    constexpr int units_per_step = S;                       // stride 2: even + odd slot per input row
    constexpr int nslots = RING_UNITS / units_per_step;     // 6 or 3
    constexpr int nsteps = S == 1 ? TILE_R + 2 : 2 * TILE_R + 1;
    static_assert(nsteps % nslots == 0, "a tile must consume whole turns of the ring (slot index is then compile-time)");

    if (threadIdx.x == 0) {
        for (int i = 0; i < RING_UNITS; ++i) {
            ptx::mbar_init(ptx::smem_u32(&bars->full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->empty[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(ptx::smem_u32(&bars->acc_full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->acc_empty[i]), 4);
        }
        ptx::mbar_init(ptx::smem_u32(&bars->w_full), 1);
        ptx::mbar_init(ptx::smem_u32(&bars->w_free), 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), 512);
        ptx::tmem_relinquish();
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_act);
        ptx::prefetch_tmap(&tmap_w);
        if (NOUT == 64) ptx::prefetch_tmap(&tmap_out);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_wait();

    if (warp == 0 && lane == 0) {
        // ================================ TMA producer ================================
        int slot = 0;
        uint32_t phase = 0;
        int cur_chunk = -1;
        uint32_t wphase = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const int chunk = t / p.tiles_per_chunk;
            int rem = t - chunk * p.tiles_per_chunk;
            if (p.rev) rem = p.tiles_per_chunk - 1 - rem;
            const int tx = rem % p.tiles_x;
            rem /= p.tiles_x;
            const int ty = rem % p.tiles_y;
            const int b = rem / p.tiles_y;
            if (chunk != cur_chunk) {
                if (cur_chunk >= 0) {   // all MMAs that read the old filter bank must have retired
                    ptx::mbar_wait(ptx::smem_u32(&bars->w_free), wphase);
                    wphase ^= 1;
                }
                ptx::mbar_expect_tx(ptx::smem_u32(&bars->w_full), W_BYTES);
                for (int tap = 0; tap < NTAPS; ++tap)
                    ptx::tma_load_2d(w_sm + tap * W_TAP, &tmap_w, ptx::smem_u32(&bars->w_full), 0, (chunk * NTAPS + tap) * NOUT);
                cur_chunk = chunk;
            }
            const int x0 = tx * TILE_W, y0 = ty * TILE_R;
            for (int j = 0; j < nsteps; ++j) {
                ptx::mbar_wait(ptx::smem_u32(&bars->empty[slot]), phase ^ 1);
                const uint32_t dst = ring_sm + slot * units_per_step * UNIT_BYTES;
                const uint32_t fb = ptx::smem_u32(&bars->full[slot]);
                ptx::mbar_expect_tx(fb, units_per_step * UNIT_BYTES);
                if (S == 1) {
                    ptx::tma_load_4d(dst, &tmap_act, fb, 0, x0 - 1, y0 - 1 + j, b);
                } else {
                    const int iy = 2 * y0 - 1 + j;
                    ptx::tma_load_4d(dst, &tmap_act, fb, 0, x0, iy, b);                     // even pixels 2x
                    ptx::tma_load_4d(dst + UNIT_BYTES, &tmap_act, fb, 64, x0 - 1, iy, b);   // odd pixels 2x+1, from x0-1
                }
                if (++slot == nslots) { slot = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        // The whole warp walks the loop converged; the elected lane's predicate guards each tcgen05 instruction.
        // Slot, tap and k offsets are compile-time, so an MMA costs two adds on pre-encoded descriptor words.
        const uint32_t leader = ptx::elect_one();
        const uint32_t idesc = ptx::make_idesc_bf16(TILE_M, NOUT);
        const uint32_t w_lo = ptx::sdesc_lo(w_sm), ring_lo = ptx::sdesc_lo(ring_sm);
        uint32_t tphase = 0;                     // ring phase at the start of the tile (flips once per tile)
        int cur_chunk = -1;
        uint32_t wphase = 0;
        int it = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
            const int chunk = t / p.tiles_per_chunk;
            const int set = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            if (chunk != cur_chunk) {
                ptx::mbar_wait(ptx::smem_u32(&bars->w_full), wphase);
                wphase ^= 1;
                cur_chunk = chunk;
            }
            ptx::mbar_wait(ptx::smem_u32(&bars->acc_empty[set]), aphase ^ 1);
            ptx::tc_fence_after();
            const uint32_t acc0 = tmem_base + set * (TILE_R * 64);
#pragma unroll
            for (int j = 0; j < nsteps; ++j) {
                const int slot = j % nslots;
                ptx::mbar_wait(ptx::smem_u32(&bars->full[slot]), tphase ^ ((j / nslots) & 1));
                ptx::tc_fence_after();
                const uint32_t a0 = ring_lo + ((slot * units_per_step * UNIT_BYTES) >> 4);
                if (S == 1 && NOUT == 16) {
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int r = j - ky;
                        if (r < 0 || r >= TILE_R) continue;
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            const uint32_t ad = a0 + ((k4 * 32) >> 4);                       // slot pixel 0 = image pixel x0 - 1
                            const uint32_t bd = w_lo + ((ky * W_TAP + k4 * 32) >> 4);
                            if ((ky | k4) != 0) ptx::umma_bf16_lo<1>(acc0 + r * 64, ad, bd, idesc, leader);
                            else ptx::umma_bf16_lo<0>(acc0 + r * 64, ad, bd, idesc, leader);
                        }
                    }
                } else if (S == 1) {
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int r = j - ky;
                        if (r < 0 || r >= TILE_R) continue;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                const uint32_t ad = a0 + ((kx * 128 + k4 * 32) >> 4);
                                const uint32_t bd = w_lo + (((ky * 3 + kx) * W_TAP + k4 * 32) >> 4);
                                if ((ky | kx | k4) != 0) ptx::umma_bf16_lo<1>(acc0 + r * 64, ad, bd, idesc, leader);
                                else ptx::umma_bf16_lo<0>(acc0 + r * 64, ad, bd, idesc, leader);
                            }
                        }
                    }
                } else {
                    // input row iy = 2*y0 - 1 + j feeds output row r with ky = j - 2r in {0,1,2}
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int r = (j >> 1) - q;
                        const int ky = j - 2 * r;
                        if (r < 0 || r >= TILE_R || ky > 2) continue;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            // kx=0: odd slot row 0; kx=1: even slot row 0; kx=2: odd slot row 1
                            const int aoff = kx == 1 ? 0 : UNIT_BYTES + (kx == 2 ? 128 : 0);
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                const uint32_t ad = a0 + ((aoff + k4 * 32) >> 4);
                                const uint32_t bd = w_lo + (((ky * 3 + kx) * W_TAP + k4 * 32) >> 4);
                                if ((ky | kx | k4) != 0) ptx::umma_bf16_lo<1>(acc0 + r * 64, ad, bd, idesc, leader);
                                else ptx::umma_bf16_lo<0>(acc0 + r * 64, ad, bd, idesc, leader);
                            }
                        }
                    }
                }
                ptx::umma_commit_pred(ptx::smem_u32(&bars->empty[slot]), leader);     // slot reusable once these MMAs retire
            }
            tphase ^= (nsteps / nslots) & 1;
            ptx::umma_commit_pred(ptx::smem_u32(&bars->acc_full[set]), leader);       // accumulators of this tile complete
            // does the next tile of this CTA switch to another filter bank?
            const int tn = t + gridDim.x;
            if (tn < p.total_tiles && tn / p.tiles_per_chunk != chunk) ptx::umma_commit_pred(ptx::smem_u32(&bars->w_free), leader);
        }
    } else if (warp >= 4) {
        // ================================ epilogue ================================
        // NOUT == 64: TMEM -> (+bias, ReLU, bf16) -> the pixel's 128 bytes into a swizzled staging buffer -> one TMA store per
        // 32-pixel row segment and warp (a thread storing its own pixel would touch 32 different cache lines per instruction).
        const int q = warp - 4;                     // TMEM lane quadrant of this warp
        int it = 0;
        uint32_t nstore = 0;
        float *xchg = bars->xchg;
        uint8_t *stg_w = smem_al + W_BYTES_MAX + RING_UNITS * UNIT_BYTES + q * 8192;
        const uint32_t stg_w_sm = stg_sm + q * 8192;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
            const int chunk = t / p.tiles_per_chunk;
            int rem = t - chunk * p.tiles_per_chunk;
            if (p.rev) rem = p.tiles_per_chunk - 1 - rem;
            const int tx = rem % p.tiles_x;
            rem /= p.tiles_x;
            const int ty = rem % p.tiles_y;
            const int b = rem / p.tiles_y;
            const int set = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            ptx::mbar_wait(ptx::smem_u32(&bars->acc_full[set]), aphase);
            ptx::tc_fence_after();
            const int px0 = tx * TILE_W + q * 32, px = px0 + lane;
            const float *bias = p.bias ? p.bias + chunk * 64 : nullptr;
#pragma unroll 1
            for (int r = 0; r < TILE_R; ++r) {
                const int y = ty * TILE_R + r;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + set * (TILE_R * 64) + r * 64;
                if (NOUT == 16) {
                    // lane i holds D[q][kx*4 + co] for input pixel q = x0 - 1 + i; out[x] = D[x-1][kx 0] + D[x][kx 1] + D[x+1][kx 2]
                    uint32_t v[32];
                    ptx::tmem_ld_x32(taddr, v);       // columns 16..31 belong to nobody (row stride is 64 columns)
                    ptx::tmem_ld_wait();
                    if (r == TILE_R - 1) {            // accumulators are in registers: hand the TMEM set back to the MMA warp
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->acc_empty[set]));
                    }
                    float lft[3], rgt[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        lft[c] = __shfl_up_sync(0xffffffffu, __uint_as_float(v[c]), 1);
                        rgt[c] = __shfl_down_sync(0xffffffffu, __uint_as_float(v[8 + c]), 1);
                    }
                    // warp seams go through shared memory (double-buffered by row parity; one named barrier per row)
                    float *xs = xchg + (r & 1) * 32;
                    if (lane == 31) { xs[q * 8 + 0] = __uint_as_float(v[0]); xs[q * 8 + 1] = __uint_as_float(v[1]); xs[q * 8 + 2] = __uint_as_float(v[2]); }
                    if (lane == 0) { xs[q * 8 + 4] = __uint_as_float(v[8]); xs[q * 8 + 5] = __uint_as_float(v[9]); xs[q * 8 + 6] = __uint_as_float(v[10]); }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (lane == 0 && q > 0) { lft[0] = xs[(q - 1) * 8 + 0]; lft[1] = xs[(q - 1) * 8 + 1]; lft[2] = xs[(q - 1) * 8 + 2]; }
                    if (lane == 31 && q < 3) { rgt[0] = xs[(q + 1) * 8 + 4]; rgt[1] = xs[(q + 1) * 8 + 5]; rgt[2] = xs[(q + 1) * 8 + 6]; }
                    const int i = q * 32 + lane, xo = tx * TILE_W + i - 1;
                    if (i >= 1 && i <= TILE_W && y < p.Ho && xo < p.Wo) {
                        const long plane = (long)p.Ho * p.Wo;
                        float *o = p.out3 + (long)b * 3 * plane + (long)y * p.Wo + xo;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            float a = (lft[c] + __uint_as_float(v[4 + c])) + rgt[c] + (bias ? __ldg(bias + c) : 0.f);
                            o[c * plane] = p.relu ? fmaxf(a, 0.f) : a;
                        }
                    }
                    continue;
                }
                uint32_t v0[32], v1[32];
                ptx::tmem_ld_x32(taddr, v0);
                ptx::tmem_ld_x32(taddr + 32, v1);
                ptx::tmem_ld_wait();
                if (r == TILE_R - 1) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->acc_empty[set]));
                }
                const uint32_t buf = nstore & 1;
                if (lane == 0) ptx::bulk_wait_read<1>();      // the store that last used this buffer has read it
                __syncwarp();
                uint8_t *rowp = stg_w + buf * 4096 + lane * 128;
#pragma unroll
                for (int c = 0; c < 64; c += 8) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float a = __uint_as_float(c < 32 ? v0[c + e] : v1[c - 32 + e]) + (bias ? __ldg(bias + c + e) : 0.f);
                        f[e] = p.relu ? fmaxf(a, 0.f) : a;
                    }
                    uint4 u;
                    __nv_bfloat162 h;
                    h = __floats2bfloat162_rn(f[0], f[1]); u.x = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(f[2], f[3]); u.y = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(f[4], f[5]); u.z = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(f[6], f[7]); u.w = *reinterpret_cast<uint32_t *>(&h);
                    *reinterpret_cast<uint4 *>(rowp + ((((c >> 3) ^ (lane & 7))) << 4)) = u;      // 128-byte swizzle: chunk ^= row % 8
                }
                ptx::fence_proxy_async();         // generic-proxy writes -> visible to the TMA engine
                __syncwarp();
                if (lane == 0) {
                    if (y < p.Ho && px0 < p.Wo) {
                        if (p.ps_r) {
                            const int rr = p.ps_r, pi = chunk / rr, pj = chunk % rr;
                            ptx::tma_store_4d(&tmap_out, stg_w_sm + buf * 4096, 0, px0, pj, (b * p.Ho + y) * rr + pi);
                        } else {
                            ptx::tma_store_4d(&tmap_out, stg_w_sm + buf * 4096, 0, px0, y, b);
                        }
                    }
                    ptx::bulk_commit();
                }
                ++nstore;
            }
        }
        if (lane == 0) ptx::bulk_wait<0>();          // all stores of this warp have left shared memory and are visible
        __syncwarp();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------- host side
thread_local int g_base_off_mode = 0;   // measured on B200: the UMMA swizzle is a function of the absolute smem address (like TMA's),
                           // so row-shifted operand views need base_offset 0 (tools/tc_probe.py, profiles/r01_tc_probe.log)
PerDeviceFlag g_attr_set;

}  // namespace

// debug key "snake": bit k set = kernel k walks its work items last to first, so that it starts on the part of its input the
// previous kernel wrote last (still in L2).  bit 0 downsample, 1 the 64 -> 3 head, 2 window stack + unembed, 3 tile conv 64 -> 64
thread_local int g_snake_mask = 0;
void tc_set_snake(int mask) { g_snake_mask = mask; }

TcEncodeFn tc_encode_fn() {
    static TcEncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (TcEncodeFn)p;
    });
    return fn;
}

int tc_available() { return tc_encode_fn() != nullptr; }
void tc_set_base_off_mode(int m) { g_base_off_mode = m; }

static int launch_conv(const bf16 *in, const bf16 *w, const float *bias, bf16 *out, float *out3, int nout, int B, int H, int W,
                       int stride, int relu, int nchunk, int ps_r, cudaStream_t st) {
    if ((reinterpret_cast<uintptr_t>(in) & 127) || (reinterpret_cast<uintptr_t>(w) & 127) || (reinterpret_cast<uintptr_t>(out) & 15))
        return TU_TC_UNSUPPORTED;
    TcEncodeFn enc = tc_encode_fn();
    if (!enc) return TU_TC_UNSUPPORTED;
    const int g_sm_count = device_sm_count();
    if (!g_attr_set.is_set()) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_tc_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_tc_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "conv3x3_tc smem attribute");
        g_attr_set.set();
    }
    CUtensorMap tm_act, tm_w;
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
        CUresult r = enc(&tm_act, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)in, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("tu: cuTensorMapEncodeTiled(activations) failed with code " + std::to_string((int)r));
            return TU_ERR_CUDA;
        }
        cuuint64_t wd[2] = {64, (cuuint64_t)nchunk * (nout == 64 ? 9 : 3) * nout}, ws[1] = {128};
        cuuint32_t wb[2] = {64, (cuuint32_t)nout}, we[2] = {1, 1};
        r = enc(&tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)w, wd, ws, wb, we, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("tu: cuTensorMapEncodeTiled(weights) failed with code " + std::to_string((int)r));
            return TU_ERR_CUDA;
        }
    }
    CUtensorMap tm_out;
    memset(&tm_out, 0, sizeof(tm_out));
    const int Ho_ = (H - 1) / stride + 1, Wo_ = (W - 1) / stride + 1;
    if (nout == 64) {
        cuuint64_t od[4], os[3];
        cuuint32_t ob[4] = {64, 32, 1, 1}, oe[4] = {1, 1, 1, 1};
        if (ps_r) {      // (c, px [stride r pixels], pj, (b*Ho + y)*r + pi): PixelShuffle addressing done by the TMA engine
            od[0] = 64; od[1] = (cuuint64_t)Wo_; od[2] = (cuuint64_t)ps_r; od[3] = (cuuint64_t)B * Ho_ * ps_r;
            os[0] = (cuuint64_t)ps_r * 128; os[1] = 128; os[2] = (cuuint64_t)Wo_ * ps_r * 128;
        } else {
            od[0] = 64; od[1] = (cuuint64_t)Wo_; od[2] = (cuuint64_t)Ho_; od[3] = (cuuint64_t)B;
            os[0] = 128; os[1] = (cuuint64_t)Wo_ * 128; os[2] = (cuuint64_t)Ho_ * Wo_ * 128;
        }
        CUresult r = enc(&tm_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)out, od, os, ob, oe, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("tu: cuTensorMapEncodeTiled(conv output) failed with code " + std::to_string((int)r));
            return TU_ERR_CUDA;
        }
    }
    ConvParams p;
    p.B = B; p.H = H; p.W = W;
    p.Ho = (H - 1) / stride + 1; p.Wo = (W - 1) / stride + 1;
    p.stride = stride; p.relu = relu; p.nchunk = nchunk; p.ps_r = ps_r;
    p.tiles_x = ceil_div(p.Wo, nout == 64 ? TILE_M : TILE_M - 2); p.tiles_y = ceil_div(p.Ho, TILE_R);
    p.tiles_per_chunk = p.tiles_x * p.tiles_y * B;
    p.total_tiles = p.tiles_per_chunk * nchunk;
    p.bias = bias; p.out = out; p.out3 = out3;
    p.rev = (g_snake_mask >> (nout == 64 ? (stride == 2 ? 0 : 3) : 1)) & 1;
    const int grid = p.total_tiles < g_sm_count ? p.total_tiles : g_sm_count;
    if (nout == 64 && stride == 1)
        launch_pdl(conv3x3_tc_kernel<64, 1>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, st, tm_act, tm_w, tm_out, p);
    else if (nout == 64)
        launch_pdl(conv3x3_tc_kernel<64, 2>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, st, tm_act, tm_w, tm_out, p);
    else
        launch_pdl(conv3x3_tc_kernel<16, 1>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, st, tm_act, tm_w, tm_out, p);
    TU_CHECK_LAUNCH("conv3x3_tc");
    return TU_OK;
}

int tc_conv3x3_c64(const bf16 *in, const bf16 *w, const float *bias, bf16 *out, int B, int H, int W, int stride, int relu,
                   int nchunk, int ps_r, cudaStream_t st) {
    if (nchunk > 1 && ps_r == 0) return TU_TC_UNSUPPORTED;
    if (stride == 2 && (W & 1)) return TU_TC_UNSUPPORTED;
    return launch_conv(in, w, bias, out, nullptr, 64, B, H, W, stride, relu, nchunk, ps_r, st);
}

int tc_conv3x3_c64_to3(const bf16 *in, const bf16 *w16, const float *bias, float *out, int B, int H, int W, int relu,
                       cudaStream_t st) {
    return launch_conv(in, w16, bias, reinterpret_cast<bf16 *>(out), out, 16, B, H, W, 1, relu, 1, 0, st);
}

}  // namespace tu
