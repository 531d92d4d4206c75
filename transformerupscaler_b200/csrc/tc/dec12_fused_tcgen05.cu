// decoder_conv1 (64 -> 64, 3x3, pad 1, + bias, ReLU) and decoder_conv2 (64 -> 3, 3x3, pad 1, + bias) in ONE kernel: the
// 64-channel map between them never reaches HBM.
//
// Reference call sites: WindowTransformer/model.py:221-222,297-298; FastTransformer/model.py:228-229,312-313;
// ResidualTransformer/model.py:111-112,156-157.
//
// Why: decoder_conv2 alone is a memory-bound kernel (it reads decoder_conv1's 64-channel output, 236 MB for 8 frames at
// 360x640, to produce 3 channels: 0.062 ms, 4.5 % of a WindowTransformer forward) and decoder_conv1 has to write those bytes
// first.  The streaming kernel (conv3x3_stream_tcgen05.cu) already turns every finished decoder_conv1 row into swizzled
// K-major bf16 rows in shared memory for its TMA store -- exactly the A operand of a second MMA.  So the row stays there:
//   * the dec1 epilogue (warps 4-7) drains a finished row from TMEM (+bias folded into the accumulator reset, ReLU, bf16)
//     into a staging buffer and signals the MMA warp instead of storing it;
//   * the MMA warp multiplies the staged row (128 pixels x 64 ci) with the head filter stacked as N = 48 =
//     [ky 2; ky 1; ky 0] x 16 rows (n = kx*4 + co): one dec1 row d adds its share to the accumulators of head rows
//     d-1, d, d+1 at once (the same vertical stacking the 64 -> 64 part uses), into a second ring of eight 16-column slots;
//   * the head epilogue (warps 8-11) drains a finished head row and adds the three kx-shifted columns
//     out[x] = P[x-1][kx 0] + P[x][kx 1] + P[x+1][kx 2] with warp shuffles (warp seams through shared memory).
// Seams.  Vertically an item recomputes one dec1 row above and below its R head rows (R + 2 dec1 rows from R + 4 input
// rows).  Horizontally the strips stay 128 pixels wide: the first / last pixel of a strip misses one term that the
// neighbouring strip holds, so both strips add their part of those two pixels with atomicAdd into memory zeroed beforehand
// (tc_dec12_zero_seams).  Each seam pixel receives exactly two addends, so the result does not depend on their order.
// dec1 pixels right of the image are forced to zero before staging (they are the head's zero padding).
//   warp 0      TMA producer (input row segments; both filter banks once)
//   warp 1      MMA issuer (dec1: 12 N<=192 MMAs per input row; head: 4 N<=48 MMAs per dec1 row, one row behind)
//   warp 2      TMEM allocation
//   warps 4-7   dec1 epilogue        warps 8-11  head epilogue
#include <cuda.h>
#include <string.h>

#include "ptx.cuh"
#include "tc_api.cuh"

namespace tu {

namespace {

constexpr int TILE_M = 128, BOXW = 136;
constexpr int UNIT_BYTES = BOXW * 128;        // 17408: one input row segment
constexpr int NACC = 6;                       // dec1 row accumulators: 6 x 64 TMEM columns
constexpr int HN = 8;                         // head row accumulators: 8 x 16 TMEM columns behind them
constexpr int HCOL0 = NACC * 64;
constexpr int W_BLK = 64 * 128;               // one (kx, ky) filter block: 64 co x 64 ci
constexpr int W_KX = 3 * W_BLK;               // per kx: [ky=2; ky=1; ky=0] stacked = 192 rows
constexpr int W_BYTES = 3 * W_KX;             // 73728
constexpr int W2_BLK = 16 * 128;              // head filter, one ky: 16 rows n = kx*4 + co
constexpr int W2_BYTES = 3 * W2_BLK;          // 6144
constexpr int STG_ROW = TILE_M * 128;         // 16384: one staged dec1 row
constexpr int OFF_W2 = W_BYTES, OFF_RING = OFF_W2 + W2_BYTES;
constexpr int MAX_RING = 6, MAX_STG = 3;
constexpr int NUM_THREADS = 384;
static_assert(HCOL0 + HN * 16 <= 512, "TMEM columns");
static_assert(OFF_RING % 1024 == 0 && UNIT_BYTES % 1024 == 0, "swizzled operands need 1024-byte alignment");
// RING input row slots, NSTG staged dec1 rows (A operand of the head MMAs); the head MMAs of a dec1 row are issued LAG input
// rows after the row completes (the dec1 epilogue has that long to drain and stage it before the issuing warp would stall)
template <int RING, int NSTG>
constexpr int smem_bytes() { return OFF_RING + RING * UNIT_BYTES + NSTG * STG_ROW + 1024 + 1024; }

struct DecParams {
    int B, H, W;
    int R;                  // head rows per work item
    int tiles_x, chunks_y, total_items;
    int defer;              // the last head rows of an item are issued during the next item (debug: 0 = at the end of the item)
    const float *bias1;     // 64
    const float *bias2;     // 3
    float *out3;            // planar fp32 (B, 3, H, W)
};

struct DecBarriers {
    uint64_t full[MAX_RING], empty[MAX_RING];
    uint64_t acc_full[NACC], acc_empty[NACC];
    uint64_t hacc_full[HN], hacc_empty[HN];
    uint64_t stg_full[MAX_STG], stg_free[MAX_STG];
    uint64_t w_full;
    uint32_t tmem_base;
    uint32_t pad;
    float xchg[2][32];      // head epilogue: seam values between its four warps, two row parities
};
static_assert(sizeof(DecBarriers) <= 1024, "barrier block too large");

template <int RING, int NSTG, int LAG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dec12_fused_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_w,
                   const __grid_constant__ CUtensorMap tmap_w2, const DecParams p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr int OFF_STG = OFF_RING + RING * UNIT_BYTES, OFF_BARS = OFF_STG + NSTG * STG_ROW;
    static_assert(smem_bytes<RING, NSTG>() <= 227 * 1024 && RING <= MAX_RING && NSTG <= MAX_STG && NSTG > LAG, "configuration");
    const uint32_t w_sm = smem0, w2_sm = smem0 + OFF_W2, ring_sm = smem0 + OFF_RING, stg_sm = smem0 + OFF_STG;
    uint8_t *smem_al = smem_raw + (smem0 - ptx::smem_u32(smem_raw));
    DecBarriers *bars = reinterpret_cast<DecBarriers *>(smem_al + OFF_BARS);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < RING; ++i) {
            ptx::mbar_init(ptx::smem_u32(&bars->full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->empty[i]), 1);
        }
        for (int i = 0; i < NACC; ++i) {
            ptx::mbar_init(ptx::smem_u32(&bars->acc_full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->acc_empty[i]), 4);
        }
        for (int i = 0; i < HN; ++i) {
            ptx::mbar_init(ptx::smem_u32(&bars->hacc_full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->hacc_empty[i]), 4);
        }
        for (int i = 0; i < NSTG; ++i) {
            ptx::mbar_init(ptx::smem_u32(&bars->stg_full[i]), 4);
            ptx::mbar_init(ptx::smem_u32(&bars->stg_free[i]), 1);
        }
        ptx::mbar_init(ptx::smem_u32(&bars->w_full), 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), 512);
        ptx::tmem_relinquish();
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_act);
        ptx::prefetch_tmap(&tmap_w);
        ptx::prefetch_tmap(&tmap_w2);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_wait();

    // work item -> frame b, head rows y0 .. y0+rows-1, pixels x0 .. x0+127, dec1 rows d_lo .. d_lo+nd-1 (one halo row each side,
    // clipped to the image: the head's zero padding above / below the image is "no contribution", not a computed row)
    auto item_geom = [&](int it, int &b, int &y0, int &rows, int &x0, int &d_lo, int &nd) {
        const int tx = it % p.tiles_x;
        int rem = it / p.tiles_x;
        const int cy = rem % p.chunks_y;
        b = rem / p.chunks_y;
        y0 = cy * p.R;
        rows = min(p.R, p.H - y0);
        x0 = tx * TILE_M;
        d_lo = max(y0 - 1, 0);
        nd = min(y0 + rows, p.H - 1) - d_lo + 1;
    };

    if (warp == 0 && lane == 0) {
        // ================================ TMA producer ================================
        // dec1 bank -> [kx][ky = 2, 1, 0][co][ci]; head bank -> [ky = 2, 1, 0][16 rows n][ci]
        ptx::mbar_expect_tx(ptx::smem_u32(&bars->w_full), W_BYTES + W2_BYTES);
        for (int kx = 0; kx < 3; ++kx)
            for (int j = 0; j < 3; ++j)
                ptx::tma_load_2d(w_sm + kx * W_KX + j * W_BLK, &tmap_w, ptx::smem_u32(&bars->w_full), 0, ((2 - j) * 3 + kx) * 64);
        for (int j = 0; j < 3; ++j) ptx::tma_load_2d(w2_sm + j * W2_BLK, &tmap_w2, ptx::smem_u32(&bars->w_full), 0, (2 - j) * 16);
        int slot = 0;
        uint32_t phase = 0;
        for (int it = blockIdx.x; it < p.total_items; it += gridDim.x) {
            int b, y0, rows, x0, d_lo, nd;
            item_geom(it, b, y0, rows, x0, d_lo, nd);
            for (int u = 0; u < nd + 2; ++u) {            // input rows d_lo - 1 .. d_lo + nd (zero-filled outside the image)
                ptx::mbar_wait(ptx::smem_u32(&bars->empty[slot]), phase ^ 1);
                const uint32_t fb = ptx::smem_u32(&bars->full[slot]);
                ptx::mbar_expect_tx(fb, UNIT_BYTES);
                ptx::tma_load_4d(ring_sm + slot * UNIT_BYTES, &tmap_act, fb, 0, x0 - 1, d_lo - 1 + u, b);
                if (++slot == RING) { slot = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (whole warp converged, elected lane issues) ================================
        const uint32_t leader = ptx::elect_one();
        const uint32_t w_lo = ptx::sdesc_lo(w_sm), w2_lo = ptx::sdesc_lo(w2_sm), ring_lo = ptx::sdesc_lo(ring_sm), stg_lo = ptx::sdesc_lo(stg_sm);
        int slot = 0;
        uint32_t phase = 0;
        // head rows trail the dec1 rows by LAG input rows, also across work items: the last rows of an item are issued during the
        // first steps of the next one (waiting for them at the end of the item would drain the tensor pipe once per item)
        struct HItem { int y0, rows, d_lo, nd, h_opened; uint32_t g0, hg0; };
        HItem cur = {0, 0, 0, 0, 0, 0u, 0u}, prv = cur;
        uint32_t head_next = 0;         // per-CTA running index of the next dec1 row whose head MMAs are to be issued
        ptx::mbar_wait(ptx::smem_u32(&bars->w_full), 0);
        // head MMAs of dec1 row g (running index; image row d of item I): it feeds head rows d-1, d, d+1 (clipped to the item) through
        // the filter blocks ky = 2, 1, 0.  Head slots are reset by the head epilogue after each drain, so every MMA accumulates.
        auto head_issue = [&](uint32_t g) {
            HItem &I = g < cur.g0 ? prv : cur;
            const int m = (int)(g - I.g0);
            const uint32_t buf = g % NSTG, par = (g / NSTG) & 1;
            const int d = I.d_lo + m;
            const int hlo = max(d - 1, I.y0), hhi = min(d + 1, I.y0 + I.rows - 1);
            while (I.h_opened <= hhi - I.y0) {            // a head row opens: its slot must have been drained and reset
                const uint32_t hg = I.hg0 + I.h_opened;
                ptx::mbar_wait(ptx::smem_u32(&bars->hacc_empty[hg & (HN - 1)]), (hg >> 3) & 1);
                ++I.h_opened;
            }
            ptx::mbar_wait(ptx::smem_u32(&bars->stg_full[buf]), par);
            ptx::tc_fence_after();
            const int n = hhi - hlo + 1, blk0 = hlo - d + 1;
            const int s0 = (I.hg0 + hlo - I.y0) & (HN - 1);
            const int n1 = min(n, HN - s0), n2 = n - n1;
            const uint32_t a_lo = stg_lo + ((buf * STG_ROW) >> 4);
            const uint32_t d1 = tmem_base + HCOL0 + s0 * 16, b1 = w2_lo + ((uint32_t)(blk0 * W2_BLK) >> 4);
            const uint32_t id1 = ptx::make_idesc_bf16(TILE_M, 16 * n1);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) ptx::umma_bf16_lo<1>(d1, a_lo + k4 * 2, b1 + k4 * 2, id1, leader);
            if (n2 > 0) {
                const uint32_t b2 = b1 + ((uint32_t)(n1 * W2_BLK) >> 4), id2 = ptx::make_idesc_bf16(TILE_M, 16 * n2);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) ptx::umma_bf16_lo<1>(tmem_base + HCOL0, a_lo + k4 * 2, b2 + k4 * 2, id2, leader);
            }
            ptx::umma_commit_pred(ptx::smem_u32(&bars->stg_free[buf]), leader);
            if (d - 1 >= I.y0) ptx::umma_commit_pred(ptx::smem_u32(&bars->hacc_full[(I.hg0 + d - 1 - I.y0) & (HN - 1)]), leader);
            if (m == I.nd - 1 && d <= I.y0 + I.rows - 1)     // bottom of the image: the last head row has no dec1 row below it
                ptx::umma_commit_pred(ptx::smem_u32(&bars->hacc_full[(I.hg0 + d - I.y0) & (HN - 1)]), leader);
        };
        for (int it = blockIdx.x; it < p.total_items; it += gridDim.x) {
            int b, y0, rows, x0, d_lo, nd;
            item_geom(it, b, y0, rows, x0, d_lo, nd);
            while (head_next < cur.g0) head_issue(head_next++);      // rows of the item before the previous one (tiny items only)
            prv = cur;
            cur.y0 = y0; cur.rows = rows; cur.d_lo = d_lo; cur.nd = nd; cur.h_opened = 0;
            cur.g0 = prv.g0 + prv.nd; cur.hg0 = prv.hg0 + prv.rows;
            const uint32_t g0 = cur.g0;
            for (int u = 0; u < nd + 2; ++u) {
                // input row u - 1 (relative to d_lo) feeds dec1 rows lo..hi with ky = u - row; filter block of row m is 2 - (u - m)
                const int lo = max(u - 2, 0), hi = min(u, nd - 1);
                ptx::mbar_wait(ptx::smem_u32(&bars->full[slot]), phase);
                if (u <= nd - 1) {                         // dec1 row `u` opens: its slot must have been drained and reset
                    const uint32_t g = g0 + u;
                    ptx::mbar_wait(ptx::smem_u32(&bars->acc_empty[g % NACC]), (g / NACC) & 1);
                }
                ptx::tc_fence_after();
                const uint32_t a_lo = ring_lo + ((uint32_t)(slot * UNIT_BYTES) >> 4);
                const int n = hi - lo + 1, blk0 = 2 - (u - lo);
                const int s0 = (int)((g0 + lo) % NACC);
                const int n1 = min(n, NACC - s0), n2 = n - n1;            // the window of slots may wrap around the ring
                const uint32_t d1 = tmem_base + s0 * 64, b1 = w_lo + ((uint32_t)(blk0 * W_BLK) >> 4);
                const uint32_t id1 = ptx::make_idesc_bf16(TILE_M, 64 * n1);
                if (n2 == 0) {
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            ptx::umma_bf16_lo<1>(d1, a_lo + ((kx * 128 + k4 * 32) >> 4), b1 + ((kx * W_KX + k4 * 32) >> 4), id1, leader);
                } else {
                    const uint32_t b2 = b1 + ((uint32_t)(n1 * W_BLK) >> 4), id2 = ptx::make_idesc_bf16(TILE_M, 64 * n2);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            ptx::umma_bf16_lo<1>(d1, a_lo + ((kx * 128 + k4 * 32) >> 4), b1 + ((kx * W_KX + k4 * 32) >> 4), id1, leader);
                            ptx::umma_bf16_lo<1>(tmem_base, a_lo + ((kx * 128 + k4 * 32) >> 4), b2 + ((kx * W_KX + k4 * 32) >> 4), id2, leader);
                        }
                }
                ptx::umma_commit_pred(ptx::smem_u32(&bars->empty[slot]), leader);
                if (u >= 2) ptx::umma_commit_pred(ptx::smem_u32(&bars->acc_full[(g0 + u - 2) % NACC]), leader);   // dec1 row u-2 is complete
                if (++slot == RING) { slot = 0; phase ^= 1; }
                // dec1 rows up to g0 + u - 2 are complete (committed); their heads follow LAG steps later
                while ((int)(head_next - g0) <= u - 2 - LAG) head_issue(head_next++);
            }
            if (!p.defer)
                while (head_next < cur.g0 + (uint32_t)cur.nd) head_issue(head_next++);
        }
        while (head_next < cur.g0 + (uint32_t)cur.nd) head_issue(head_next++);
    } else if (warp >= 4 && warp < 8) {
        // ================================ dec1 epilogue: one dec1 row at a time ================================
        const int q = warp - 4;
        uint32_t g = 0;
        // a row accumulator is reset to the bias vector after its drain, so the bias is never added here
        uint32_t init0[32], init1[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            init0[c] = __float_as_uint(__ldg(p.bias1 + c));
            init1[c] = __float_as_uint(__ldg(p.bias1 + 32 + c));
        }
        for (int c = 0; c < NACC * 64; c += 64) {
            ptx::tmem_st_x32(tmem_base + ((uint32_t)(q * 32) << 16) + c, init0);
            ptx::tmem_st_x32(tmem_base + ((uint32_t)(q * 32) << 16) + c + 32, init1);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0)
            for (int sl = 0; sl < NACC; ++sl) ptx::mbar_arrive(ptx::smem_u32(&bars->acc_empty[sl]));
        for (int it = blockIdx.x; it < p.total_items; it += gridDim.x) {
            int b, y0, rows, x0, d_lo, nd;
            item_geom(it, b, y0, rows, x0, d_lo, nd);
            const bool inside = x0 + q * 32 + lane < p.W;
#pragma unroll 1
            for (int m = 0; m < nd; ++m, ++g) {
                const uint32_t sl = g % NACC;
                ptx::mbar_wait(ptx::smem_u32(&bars->acc_full[sl]), (g / NACC) & 1);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + sl * 64;
                uint32_t v0[32], v1[32];
                ptx::tmem_ld_x32(taddr, v0);
                ptx::tmem_ld_x32(taddr + 32, v1);
                ptx::tmem_ld_wait();
                ptx::tmem_st_x32(taddr, init0);                                 // reset the slot for the row that opens it next
                ptx::tmem_st_x32(taddr + 32, init1);
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->acc_empty[sl]));      // the row is in registers, the slot is reset
                const uint32_t buf = g % NSTG;
                ptx::mbar_wait(ptx::smem_u32(&bars->stg_free[buf]), ((g / NSTG) & 1) ^ 1);  // the head MMAs that read this buffer have retired
                uint8_t *rowp = smem_al + OFF_STG + buf * STG_ROW + q * 4096 + lane * 128;
#pragma unroll
                for (int c = 0; c < 64; c += 8) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float a = __uint_as_float(c < 32 ? v0[c + e] : v1[c - 32 + e]);
                        f[e] = inside ? fmaxf(a, 0.f) : 0.f;                     // pixels right of the image are the head's zero padding
                    }
                    uint4 uu;
                    __nv_bfloat162 h;
                    h = __floats2bfloat162_rn(f[0], f[1]); uu.x = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(f[2], f[3]); uu.y = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(f[4], f[5]); uu.z = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(f[6], f[7]); uu.w = *reinterpret_cast<uint32_t *>(&h);
                    *reinterpret_cast<uint4 *>(rowp + ((((c >> 3) ^ (lane & 7))) << 4)) = uu;      // 128-byte swizzle: chunk ^= row % 8
                }
                ptx::fence_proxy_async();         // generic-proxy writes -> visible to the tensor core's operand reads
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->stg_full[buf]));
            }
        }
    } else if (warp >= 8) {
        // ================================ head epilogue: one head row at a time ================================
        const int q = warp - 8;
        uint32_t hg = 0;
        // reset value of a head slot: the bias in the kx = 1 columns (n = 4 + co), zero elsewhere
        uint32_t hinit[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) hinit[c] = (c >= 4 && c < 7) ? __float_as_uint(__ldg(p.bias2 + c - 4)) : 0u;
        for (int s = 0; s < HN; ++s) ptx::tmem_st_x16(tmem_base + ((uint32_t)(q * 32) << 16) + HCOL0 + s * 16, hinit);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0)
            for (int s = 0; s < HN; ++s) ptx::mbar_arrive(ptx::smem_u32(&bars->hacc_empty[s]));
        const long plane = (long)p.H * p.W;
        for (int it = blockIdx.x; it < p.total_items; it += gridDim.x) {
            int b, y0, rows, x0, d_lo, nd;
            item_geom(it, b, y0, rows, x0, d_lo, nd);
            const int x = x0 + q * 32 + lane;
            const bool first = q == 0 && lane == 0, last = q == 3 && lane == 31;
            const bool left_seam = first && x0 > 0, right_seam = last && x0 + TILE_M < p.W;
            float *o = p.out3 + (long)b * 3 * plane + (long)y0 * p.W + x;
#pragma unroll 1
            for (int h = 0; h < rows; ++h, ++hg, o += p.W) {
                const uint32_t sl = hg & (HN - 1);
                ptx::mbar_wait(ptx::smem_u32(&bars->hacc_full[sl]), (hg >> 3) & 1);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + HCOL0 + sl * 16;
                uint32_t v[16];
                ptx::tmem_ld_x16(taddr, v);
                ptx::tmem_ld_wait();
                ptx::tmem_st_x16(taddr, hinit);
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->hacc_empty[sl]));
                // lane i holds P[kx*4 + co] of pixel x; out[x] = P[x-1][kx 0] + P[x][kx 1] + P[x+1][kx 2]
                float lft[3], rgt[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    lft[c] = __shfl_up_sync(0xffffffffu, __uint_as_float(v[c]), 1);
                    rgt[c] = __shfl_down_sync(0xffffffffu, __uint_as_float(v[8 + c]), 1);
                }
                float *xs = bars->xchg[hg & 1];
                if (lane == 31) { xs[q * 8 + 0] = __uint_as_float(v[0]); xs[q * 8 + 1] = __uint_as_float(v[1]); xs[q * 8 + 2] = __uint_as_float(v[2]); }
                if (lane == 0) { xs[q * 8 + 4] = __uint_as_float(v[8]); xs[q * 8 + 5] = __uint_as_float(v[9]); xs[q * 8 + 6] = __uint_as_float(v[10]); }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (lane == 0 && q > 0) { lft[0] = xs[(q - 1) * 8 + 0]; lft[1] = xs[(q - 1) * 8 + 1]; lft[2] = xs[(q - 1) * 8 + 2]; }
                if (lane == 31 && q < 3) { rgt[0] = xs[(q + 1) * 8 + 4]; rgt[1] = xs[(q + 1) * 8 + 5]; rgt[2] = xs[(q + 1) * 8 + 6]; }
                if (left_seam) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        atomicAdd(o + c * plane, __uint_as_float(v[4 + c]) + rgt[c]);
                        atomicAdd(o + c * plane - 1, __uint_as_float(v[8 + c]));       // this strip's share of the pixel left of it
                    }
                } else if (right_seam) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        atomicAdd(o + c * plane, lft[c] + __uint_as_float(v[4 + c]));
                        atomicAdd(o + c * plane + 1, __uint_as_float(v[c]));           // this strip's share of the pixel right of it
                    }
                } else if (x < p.W) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float l = first ? 0.f : lft[c], r = last ? 0.f : rgt[c];
                        o[c * plane] = (l + __uint_as_float(v[4 + c])) + r;
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

thread_local int g_enable_dec12 = 1;
thread_local int g_dec12_defer = 1;
thread_local int g_dec12_variant = 1;      // 0: 6 input slots, 2 staged rows, lag 1;  1: 5 input slots, 3 staged rows, lag 2

}  // namespace

void tc_set_dec12_fused(int on) { g_enable_dec12 = on & 1; g_dec12_variant = on & 2 ? 0 : 1; g_dec12_defer = on & 4 ? 0 : 1; }
bool tc_dec12_fused_enabled() { return g_enable_dec12 != 0; }

// Zero the pixels two 128-pixel strips share (columns 127|128, 255|256, ...) of a planar fp32 (B, 3, H, W) image; must be
// ordered before tc_dec12_fused on the same stream.  An ordinary launch (no programmatic serialization): the buffer may still
// be read by the kernel in front of it (the previous forward's bicubic kernel reads the same workspace slot).
__global__ void dec12_zero_seams_kernel(float *out3, long rows, int W, int seams) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * seams) return;
    const long row = i / seams;
    const int s = (int)(i - row * seams) + 1;
    float *o = out3 + row * W + (long)s * TILE_M;
    o[-1] = 0.f;
    o[0] = 0.f;
}

int tc_dec12_zero_seams(float *out3, int B, int H, int W, cudaStream_t st) {
    const int seams = ceil_div(W, TILE_M) - 1;
    if (seams <= 0) return TU_OK;
    const long n = (long)B * 3 * H * seams;
    dec12_zero_seams_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out3, (long)B * 3 * H, W, seams);
    TU_CHECK_LAUNCH("dec12_zero_seams");
    return TU_OK;
}

// in NHWC bf16 (B, H, W, 64); w1 = [9 taps][64 co][64 ci] bf16 (the banks of the 64 -> 64 kernels); w16 = [3 ky][16 rows
// n = kx*4 + co][64 ci] bf16 (the bank of the 64 -> 3 head); out3 planar fp32 (B, 3, H, W) with its seams zeroed beforehand
int tc_dec12_fused(const bf16 *in, const bf16 *w1, const float *bias1, const bf16 *w16, const float *bias2, float *out3, int B, int H,
                   int W, cudaStream_t st) {
    if (!g_enable_dec12 || !bias1 || !bias2) return TU_TC_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(in) & 127) || (reinterpret_cast<uintptr_t>(w1) & 127) || (reinterpret_cast<uintptr_t>(w16) & 127) ||
        (reinterpret_cast<uintptr_t>(out3) & 3) || (long)B * 3 * H * W >= (1L << 31))
        return TU_TC_UNSUPPORTED;
    TcEncodeFn enc = tc_encode_fn();
    if (!enc) return TU_TC_UNSUPPORTED;
    const int g_sm_count_d = device_sm_count();
    CUtensorMap tm_act, tm_w, tm_w2;
    {
        cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
        cuuint32_t box[4] = {64, (cuuint32_t)BOXW, 1, 1}, estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm_act, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t wd[2] = {64, 9 * 64}, ws[1] = {128};
        cuuint32_t wb[2] = {64, 64}, we[2] = {1, 1};
        if (r == CUDA_SUCCESS)
            r = enc(&tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)w1, wd, ws, wb, we, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t w2d[2] = {64, 3 * 16};
        cuuint32_t w2b[2] = {64, 16};
        if (r == CUDA_SUCCESS)
            r = enc(&tm_w2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)w16, w2d, ws, w2b, we, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("tu: cuTensorMapEncodeTiled(dec12 fused) failed with code " + std::to_string((int)r));
            return TU_ERR_CUDA;
        }
    }
    DecParams p;
    p.B = B; p.H = H; p.W = W;
    p.tiles_x = ceil_div(W, TILE_M);
    // head rows per work item: tall items amortise the four halo input rows, but the item count should fill whole waves of SMs
    int bestR = H < 8 ? H : 8;
    double best = 1e30;
    for (int R = 8; R <= 64 && R <= (H > 8 ? H : 8); ++R) {
        const long items = (long)p.tiles_x * ceil_div(H, R) * B;
        const long waves = (items + g_sm_count_d - 1) / g_sm_count_d;
        const double cost = (double)waves * (R + 4);       // steps executed by the busiest SM
        if (cost < best - 1e-9) { best = cost; bestR = R; }
    }
    p.R = bestR;
    p.chunks_y = ceil_div(H, p.R);
    p.total_items = p.tiles_x * p.chunks_y * B;
    p.bias1 = bias1; p.bias2 = bias2; p.out3 = out3;
    p.defer = g_dec12_defer;
    const int grid = p.total_items < g_sm_count_d ? p.total_items : g_sm_count_d;
    static PerDeviceFlag attr_set;
    if (!attr_set.is_set()) {
        cudaError_t e = cudaFuncSetAttribute(dec12_fused_kernel<6, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<6, 2>());
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(dec12_fused_kernel<5, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<5, 3>());
        if (e != cudaSuccess) return cuda_fail(e, "dec12_fused smem attribute");
        attr_set.set();
    }
    if (g_dec12_variant == 0)
        launch_pdl(dec12_fused_kernel<6, 2, 1>, dim3(grid), dim3(NUM_THREADS), smem_bytes<6, 2>(), st, tm_act, tm_w, tm_w2, p);
    else
        launch_pdl(dec12_fused_kernel<5, 3, 2>, dim3(grid), dim3(NUM_THREADS), smem_bytes<5, 3>(), st, tm_act, tm_w, tm_w2, p);
    TU_CHECK_LAUNCH("dec12_fused");
    return TU_OK;
}

}  // namespace tu
