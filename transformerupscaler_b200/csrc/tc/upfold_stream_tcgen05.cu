// FastTransformer's up-sampling branch with its last three linear steps folded into one convolution:
//   up1 stage  Conv2d(64 -> 64 r^2, 3x3, bias)  ->  nn.PixelShuffle(r)  ->  up1_conv = Conv2d(64 -> 3, 3x3, no bias) + ReLU
//   (FastTransformer/utils.py:43-98 with act=False, utils.py:13-40; model.py:207-208,264-265).
// No non-linearity sits between the two convolutions, so their composition is ONE convolution from the low-resolution
// 64-channel map to the 3 r^2 sub-pixel outputs of every low-resolution pixel, with a 5x5 low-resolution footprint
// (packing.py::fold_up1 builds its filter and bias in fp64 from the state_dict): 2*64*25*3r^2 FLOP per pixel instead of
// 2*64*9*64r^2 + 2*64*9*3r^2 (8.04x fewer at every r), and the (B, rH, rW, 64) intermediate -- 1.9 GB for four 720p frames
// at r = 2, written once and read once -- never exists.  The fold is exact in real arithmetic except on the outermost
// ring of high-resolution pixels, where the reference zero-pads the INTERMEDIATE map: those pixels are recomputed with
// border variants of the folded filter by upfold_ring_kernel (below).
//
// Kernel = the streaming ky-stacked convolution (conv3x3_stream_tcgen05.cu) generalised to 5x5 taps and a narrow N:
//   * a work item is a column strip of 128 low-res pixels x R_rows rows; input row s feeds output rows s-2..s+2 through
//     ky = 4..0 with the SAME operand view, so one 128 x (5 NO) x 16 MMA against the stacked filter [W(ky=4);..;W(ky=0)]
//     accumulates into five consecutive row accumulators (a ring of sixteen NO-column slots in TMEM); five kx shifts of
//     the operand view x four k-steps = 20 MMAs per input row;
//   * NO = 16 (r = 2: 12 outputs), 32 (r = 3: 27 outputs), 48 (r = 6: three chunks, one colour = 36 outputs each; eight row
//     accumulators, three input-row slots and one staging buffer per epilogue warp to fit 227 KB);
//   * epilogue: row accumulator -> +bias, ReLU -> fp32 staging laid out as the high-resolution row segments
//     [(c, i)][32 pixels x r] -> one TMA store per (c, i) into the planar (B, 3, rH, rW) image: PixelShuffle is address math.
#include <cuda.h>
#include <string.h>

#include "ptx.cuh"
#include "tc_api.cuh"

namespace tu {

namespace {

constexpr int TILE_M = 128, BOXW = 136;
constexpr int UNIT_BYTES = BOXW * 128;        // one input row segment (136 pixels from x0 - 2)
constexpr int NACC_MAX = 16;                  // output-row accumulators in TMEM (8 when a row needs 48 columns)
constexpr int NUM_THREADS = 256;

// KS = filter size: 5 for the folded up-sampling branch, 3 for the plain 64 -> 3 heads (R = 1: decoder_conv2, up1_conv)
template <int NO, int R, int KS>
struct FoldCfg {
    static constexpr int RPC = R == 1 ? 3 : R == 2 ? 6 : R == 3 ? 9 : NO == 48 ? 6 : 5;   // (c, i) output rows per chunk
    static constexpr int NCHUNK = (3 * R + RPC - 1) / RPC;           // 1, 1, 1; r = 6: 3 (one colour per chunk, NO = 48) or 4
    static constexpr int RING = NO == 16 ? 6 : NO == 32 ? 5 : 3;     // input row slots
    static constexpr int NACC = NO == 48 ? 8 : 16;                   // row accumulators: NACC * NO <= 512 TMEM columns
    static constexpr int NBUF = NO == 48 ? 1 : 2;                    // staging buffers per epilogue warp
    static constexpr int W_BLK = NO * 128;                           // one (kx, ky) filter block: NO rows x 64 ci
    static constexpr int W_KX = KS * W_BLK;                          // per kx: [ky = KS-1 .. 0] stacked
    static constexpr int W_BYTES = KS * W_KX;
    static constexpr int ROW_BYTES = 32 * R * 4;                     // one staged high-res row segment of a warp
    static constexpr int STG_WARP = RPC * ROW_BYTES;                 // per buffer
    static constexpr int STG_BYTES = 4 * NBUF * STG_WARP;
    static constexpr int SMEM_BYTES = W_BYTES + RING * UNIT_BYTES + STG_BYTES + 512 + 1024;
    static_assert(RPC * R <= NO, "chunk does not fit its accumulator");
    static_assert(W_BYTES % 1024 == 0 && SMEM_BYTES <= 232448, "shared memory layout");
};

struct FoldParams {
    int B, H, W, relu;
    int R_rows;             // output (low-res) rows per work item
    int tiles_x, chunks_y, items_per_chunk, total_items;
    const float *bias;      // (NCHUNK * NO), zero in unused columns
};

struct FoldBarriers {
    uint64_t full[6], empty[6];
    uint64_t acc_full[NACC_MAX], acc_empty[NACC_MAX];
    uint64_t w_full, w_free;
    uint32_t tmem_base;
};
static_assert(sizeof(FoldBarriers) <= 512, "barrier block too large");

// NO consecutive TMEM columns of this warp's 32 lanes <-> registers
template <int NO> __device__ __forceinline__ void tm_ld(uint32_t taddr, uint32_t (&v)[NO]) {
    if constexpr (NO == 16) {
        ptx::tmem_ld_x16(taddr, v);
    } else if constexpr (NO == 32) {
        ptx::tmem_ld_x32(taddr, v);
    } else {
        uint32_t a[32], b[16];
        ptx::tmem_ld_x32(taddr, a);
        ptx::tmem_ld_x16(taddr + 32, b);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = a[j];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[32 + j] = b[j];
    }
}
template <int NO> __device__ __forceinline__ void tm_st(uint32_t taddr, const uint32_t (&v)[NO]) {
    if constexpr (NO == 16) {
        ptx::tmem_st_x16(taddr, v);
    } else if constexpr (NO == 32) {
        ptx::tmem_st_x32(taddr, v);
    } else {
        uint32_t a[32], b[16];
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = v[j];
#pragma unroll
        for (int j = 0; j < 16; ++j) b[j] = v[32 + j];
        ptx::tmem_st_x32(taddr, a);
        ptx::tmem_st_x16(taddr + 32, b);
    }
}

template <int NO, int R, int KS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
upfold_stream_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_w,
                     const __grid_constant__ CUtensorMap tmap_out, const FoldParams p) {
    using Cfg = FoldCfg<NO, R, KS>;
    constexpr int PAD = KS / 2;
    constexpr int NACC = Cfg::NACC, NBUF = Cfg::NBUF;
    constexpr int RING = Cfg::RING, W_BLK = Cfg::W_BLK, W_KX = Cfg::W_KX, W_BYTES = Cfg::W_BYTES, RPC = Cfg::RPC;
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_sm = smem0, ring_sm = smem0 + W_BYTES, stg_sm = ring_sm + RING * UNIT_BYTES;
    uint8_t *smem_al = smem_raw + (smem0 - ptx::smem_u32(smem_raw));
    FoldBarriers *bars = reinterpret_cast<FoldBarriers *>(smem_al + W_BYTES + RING * UNIT_BYTES + Cfg::STG_BYTES);
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
        ptx::prefetch_tmap(&tmap_out);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_wait();

    // work item -> (output chunk, frame b, first row y0, rows, first pixel x0); chunk-major, so a CTA swaps filter banks rarely
    auto item_geom = [&](int it, int &b, int &y0, int &rows, int &x0) -> int {
        const int chunk = it / p.items_per_chunk;
        it -= chunk * p.items_per_chunk;
        const int tx = it % p.tiles_x;
        int rem = it / p.tiles_x;
        const int cy = rem % p.chunks_y;
        b = rem / p.chunks_y;
        y0 = cy * p.R_rows;
        rows = min(p.R_rows, p.H - y0);
        x0 = tx * TILE_M;
        return chunk;
    };

    if (warp == 0 && lane == 0) {
        // ================================ TMA producer ================================
        int slot = 0, cur_chunk = -1;
        uint32_t phase = 0, wfree_ph = 0;
        for (int it = blockIdx.x; it < p.total_items; it += gridDim.x) {
            int b, y0, rows, x0;
            const int chunk = item_geom(it, b, y0, rows, x0);
            if (chunk != cur_chunk) {
                if (cur_chunk >= 0) {          // all MMAs that read the old bank must have retired
                    ptx::mbar_wait(ptx::smem_u32(&bars->w_free), wfree_ph);
                    wfree_ph ^= 1;
                }
                ptx::mbar_expect_tx(ptx::smem_u32(&bars->w_full), W_BYTES);
                for (int kx = 0; kx < KS; ++kx)     // global order = shared order: [chunk][kx][ky = KS-1..0][NO][64]
                    ptx::tma_load_2d(w_sm + kx * W_KX, &tmap_w, ptx::smem_u32(&bars->w_full), 0, (chunk * KS + kx) * KS * NO);
                cur_chunk = chunk;
            }
            for (int u = 0; u < rows + 2 * PAD; ++u) {    // input rows y0 - PAD .. y0 + rows + PAD - 1
                ptx::mbar_wait(ptx::smem_u32(&bars->empty[slot]), phase ^ 1);
                const uint32_t fb = ptx::smem_u32(&bars->full[slot]);
                ptx::mbar_expect_tx(fb, UNIT_BYTES);
                ptx::tma_load_4d(ring_sm + slot * UNIT_BYTES, &tmap_act, fb, 0, x0 - PAD, y0 - PAD + u, b);
                if (++slot == RING) { slot = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (whole warp converged, elected lane issues) ================================
        const uint32_t leader = ptx::elect_one();
        const uint32_t w_lo = ptx::sdesc_lo(w_sm), ring_lo = ptx::sdesc_lo(ring_sm);
        int slot = 0, cur_chunk = -1;
        uint32_t phase = 0, wfull_ph = 0;
        uint32_t g0 = 0;                                   // per-CTA running index of the item's first output row
        for (int it = blockIdx.x; it < p.total_items; it += gridDim.x) {
            int b, y0, rows, x0;
            const int chunk = item_geom(it, b, y0, rows, x0);
            if (chunk != cur_chunk) {
                ptx::mbar_wait(ptx::smem_u32(&bars->w_full), wfull_ph);
                wfull_ph ^= 1;
                cur_chunk = chunk;
            }
            for (int u = 0; u < rows + 2 * PAD; ++u) {
                // input row y0 - PAD + u feeds output rows m in [lo, hi] with ky = u - m; the stacked block of row m is KS-1 - (u - m).
                // A slot is zero when its row opens (the epilogue clears it after draining), so every MMA accumulates.
                const int lo = max(u - (KS - 1), 0), hi = min(u, rows - 1);
                ptx::mbar_wait(ptx::smem_u32(&bars->full[slot]), phase);
                if (u <= rows - 1) {                       // row u opens: its slot must have been drained and cleared
                    const uint32_t g = g0 + u;
                    ptx::mbar_wait(ptx::smem_u32(&bars->acc_empty[g & (NACC - 1)]), (g / NACC) & 1);
                }
                ptx::tc_fence_after();
                const uint32_t a_lo = ring_lo + ((uint32_t)(slot * UNIT_BYTES) >> 4);
                const int n = hi - lo + 1, blk0 = KS - 1 - (u - lo);
                const int s0 = (g0 + lo) & (NACC - 1);
                const int n1 = min(n, NACC - s0), n2 = n - n1;            // the window of slots may wrap around the ring
                const uint32_t d1 = tmem_base + s0 * NO, b1 = w_lo + ((uint32_t)(blk0 * W_BLK) >> 4);
                const uint32_t id1 = ptx::make_idesc_bf16(TILE_M, NO * n1);
                if (n2 == 0) {
#pragma unroll
                    for (int kx = 0; kx < KS; ++kx)
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            ptx::umma_bf16_lo<1>(d1, a_lo + ((kx * 128 + k4 * 32) >> 4), b1 + ((kx * W_KX + k4 * 32) >> 4), id1, leader);
                } else {
                    const uint32_t b2 = b1 + ((uint32_t)(n1 * W_BLK) >> 4), id2 = ptx::make_idesc_bf16(TILE_M, NO * n2);
#pragma unroll
                    for (int kx = 0; kx < KS; ++kx)
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            ptx::umma_bf16_lo<1>(d1, a_lo + ((kx * 128 + k4 * 32) >> 4), b1 + ((kx * W_KX + k4 * 32) >> 4), id1, leader);
                            ptx::umma_bf16_lo<1>(tmem_base, a_lo + ((kx * 128 + k4 * 32) >> 4), b2 + ((kx * W_KX + k4 * 32) >> 4), id2, leader);
                        }
                }
                ptx::umma_commit_pred(ptx::smem_u32(&bars->empty[slot]), leader);
                if (u >= KS - 1) ptx::umma_commit_pred(ptx::smem_u32(&bars->acc_full[(g0 + u - (KS - 1)) & (NACC - 1)]), leader);   // row u-(KS-1) is complete
                if (++slot == RING) { slot = 0; phase ^= 1; }
            }
            g0 += rows;
            const int nx = it + gridDim.x;      // does this CTA's next item use another filter bank?
            if (nx < p.total_items && nx / p.items_per_chunk != chunk) ptx::umma_commit_pred(ptx::smem_u32(&bars->w_free), leader);
        }
    } else if (warp >= 4) {
        // ================================ epilogue: one low-res output row at a time ================================
        const int q = warp - 4;
        uint32_t g = 0, nstore = 0;
        // value a row accumulator is reset to after its drain: with a single filter bank the bias vector itself, so that the drain
        // does not add it; zero otherwise
        constexpr bool ACC_BIAS = Cfg::NCHUNK == 1;
        uint32_t zero[NO];
#pragma unroll
        for (int c = 0; c < NO; ++c) zero[c] = ACC_BIAS ? __float_as_uint(__ldg(p.bias + c)) : 0u;
        for (int c = 0; c < NACC * NO; c += NO) tm_st<NO>(tmem_base + ((uint32_t)(q * 32) << 16) + c, zero);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0)
            for (int sl = 0; sl < NACC; ++sl) ptx::mbar_arrive(ptx::smem_u32(&bars->acc_empty[sl]));
        uint8_t *stg_w = smem_al + W_BYTES + RING * UNIT_BYTES + q * NBUF * Cfg::STG_WARP;
        const uint32_t stg_w_sm = stg_sm + q * NBUF * Cfg::STG_WARP;
        float bch[RPC * R];                          // bias of the current chunk (several chunks: added in the drain, from registers)
        int bch_chunk = -1;
        const int oH = p.H * R;
        for (int it = blockIdx.x; it < p.total_items; it += gridDim.x) {
            int b, y0, rows, x0;
            const int chunk = item_geom(it, b, y0, rows, x0);
            const int px0 = x0 + q * 32;
            if (!ACC_BIAS && chunk != bch_chunk) {
#pragma unroll
                for (int e = 0; e < RPC * R; ++e) bch[e] = __ldg(p.bias + chunk * NO + e);
                bch_chunk = chunk;
            }
            const int nrow = min(RPC, 3 * R - chunk * RPC);          // (c, i) rows this chunk really has
#pragma unroll 1
            for (int m = 0; m < rows; ++m, ++g) {
                const int sl = g & (NACC - 1);
                ptx::mbar_wait(ptx::smem_u32(&bars->acc_full[sl]), (g / NACC) & 1);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + sl * NO;
                uint32_t v[NO];
                tm_ld<NO>(taddr, v);
                ptx::tmem_ld_wait();
                tm_st<NO>(taddr, zero);                                         // reset the slot for the row that opens it next
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(ptx::smem_u32(&bars->acc_empty[sl]));      // the row is in registers, the slot is zero
                    if constexpr (NBUF == 2) ptx::bulk_wait_read<1>();          // the stores that last used this buffer have read it
                    else ptx::bulk_wait_read<0>();
                }
                __syncwarp();
                const uint32_t buf = NBUF == 2 ? (nstore & 1) : 0;
                float *rowp = reinterpret_cast<float *>(stg_w + buf * Cfg::STG_WARP) + lane * R;
#pragma unroll
                for (int qq = 0; qq < RPC; ++qq)
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const int nidx = qq * R + j;
                        {
                            const float a = ACC_BIAS ? __uint_as_float(v[nidx]) : __uint_as_float(v[nidx]) + bch[nidx];
                            rowp[qq * 32 * R + j] = p.relu ? fmaxf(a, 0.f) : a;
                        }
                    }
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if (px0 < p.W) {
                        for (int qq = 0; qq < nrow; ++qq) {
                            const int ci = chunk * RPC + qq, c = ci / R, i = ci - c * R;
                            ptx::tma_store_2d(&tmap_out, stg_w_sm + buf * Cfg::STG_WARP + qq * Cfg::ROW_BYTES, px0 * R,
                                              (b * 3 + c) * oH + (y0 + m) * R + i);
                        }
                    }
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
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

// ---- border ring: the reference zero-pads the high-resolution intermediate, so an output pixel in the first / last row or
// column sums only the up1_conv taps that fall inside the image.  ring_w holds the folded filter for the nine (row, column)
// border cases [vy*3 + vx][o = (c*r + i)*r + j][tap = dy*5 + dx][ci] in fp32, ring_b the matching bias.  One warp per pixel.
__global__ void __launch_bounds__(256)
upfold_ring_kernel(const bf16 *__restrict__ in, const float *__restrict__ ring_w, const float *__restrict__ ring_b,
                   float *__restrict__ out, int B, int H, int W, int r) {
    pdl_trigger();
    const int oH = H * r, oW = W * r;
    const int per_frame = 2 * oW + 2 * (oH - 2);
    const int wid = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    pdl_wait();
    if (wid >= B * per_frame) return;
    const int b = wid / per_frame;
    int k = wid - b * per_frame, Y, X;
    if (k < oW) { Y = 0; X = k; }
    else if (k < 2 * oW) { Y = oH - 1; X = k - oW; }
    else { k -= 2 * oW; Y = 1 + (k >> 1); X = (k & 1) ? oW - 1 : 0; }
    const int vy = Y == 0 ? 0 : Y == oH - 1 ? 2 : 1, vx = X == 0 ? 0 : X == oW - 1 ? 2 : 1;
    const int y = Y / r, i = Y - y * r, x = X / r, j = X - x * r;
    const int nco = 3 * r * r;
    float acc[3] = {0.f, 0.f, 0.f};
    const float *wv = ring_w + ((long)(vy * 3 + vx) * nco) * 25 * 64 + lane * 2;
    const int o0 = i * r + j;                        // output o = c*r*r + o0
    for (int dy = 0; dy < 5; ++dy) {
        const int iy = y + dy - 2;
        if (iy < 0 || iy >= H) continue;
        const bf16 *row = in + ((long)b * H + iy) * W * 64 + lane * 2;
        // the five taps of a row are independent: all their loads are in flight together (the kernel is latency bound)
        float2 fv[5];
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
            const int ix = x + dx - 2;
            fv[dx] = (ix >= 0 && ix < W) ? __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(row + (long)ix * 64)) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float2 w2[5];
#pragma unroll
            for (int dx = 0; dx < 5; ++dx)
                w2[dx] = *reinterpret_cast<const float2 *>(wv + ((long)(c * r * r + o0) * 25 + dy * 5 + dx) * 64);
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) acc[c] = fmaf(fv[dx].x, w2[dx].x, fmaf(fv[dx].y, w2[dx].y, acc[c]));
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], s);
    }
    if (lane < 3) {
        const int o = (lane * r + i) * r + j;
        const float a = (lane == 0 ? acc[0] : lane == 1 ? acc[1] : acc[2]) + ring_b[(vy * 3 + vx) * nco + o];
        out[(((long)b * 3 + lane) * oH + Y) * oW + X] = fmaxf(a, 0.f);
    }
}


template <int NO, int R, int KS>
int launch_fold(const bf16 *in, const void *wbank, const float *bias, int relu, float *out, int B, int H, int W, cudaStream_t st) {
    using Cfg = FoldCfg<NO, R, KS>;
    static PerDeviceFlag attr;
    TcEncodeFn enc = tc_encode_fn();
    if (!enc) return TU_TC_UNSUPPORTED;
    const int g_sm_count_f = device_sm_count();
    if (!attr.is_set()) {
        cudaError_t e = cudaFuncSetAttribute(upfold_stream_kernel<NO, R, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "upfold_stream smem attribute");
        attr.set();
    }
    CUtensorMap tm_act, tm_w, tm_out;
    {
        cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
        cuuint32_t box[4] = {64, (cuuint32_t)BOXW, 1, 1}, estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm_act, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t wd[2] = {64, (cuuint64_t)Cfg::NCHUNK * KS * KS * NO}, ws[1] = {128};
        cuuint32_t wb[2] = {64, (cuuint32_t)(KS * NO)}, we[2] = {1, 1};
        if (r == CUDA_SUCCESS)
            r = enc(&tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)wbank, wd, ws, wb, we, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        // planar fp32 image as (x, plane row): a warp stores one (c, i) row segment of 32 r pixels at a time
        cuuint64_t od[2] = {(cuuint64_t)W * R, (cuuint64_t)B * 3 * H * R}, os[1] = {(cuuint64_t)W * R * 4};
        cuuint32_t ob[2] = {(cuuint32_t)(32 * R), 1};
        if (r == CUDA_SUCCESS)
            r = enc(&tm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)out, od, os, ob, we, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("tu: cuTensorMapEncodeTiled(upfold) failed with code " + std::to_string((int)r));
            return TU_ERR_CUDA;
        }
    }
    FoldParams p;
    p.B = B; p.H = H; p.W = W; p.relu = relu;
    p.tiles_x = ceil_div(W, TILE_M);
    // rows per work item: tall items amortise the halo rows, but the item count should fill whole waves of SMs
    int bestR = H < 8 ? H : 8;
    double best = 1e30;
    for (int Rr = 8; Rr <= 96 && Rr <= (H > 8 ? H : 8); ++Rr) {
        const long items = (long)p.tiles_x * ceil_div(H, Rr) * B * Cfg::NCHUNK;
        const long waves = (items + g_sm_count_f - 1) / g_sm_count_f;
        const double cost = (double)waves * (Rr + KS - 1);
        if (cost < best - 1e-9) { best = cost; bestR = Rr; }
    }
    p.R_rows = bestR;
    p.chunks_y = ceil_div(H, p.R_rows);
    p.items_per_chunk = p.tiles_x * p.chunks_y * B;
    p.total_items = p.items_per_chunk * Cfg::NCHUNK;
    p.bias = bias;
    const int grid = p.total_items < g_sm_count_f ? p.total_items : g_sm_count_f;
    launch_pdl(upfold_stream_kernel<NO, R, KS>, dim3(grid), dim3(NUM_THREADS), Cfg::SMEM_BYTES, st, tm_act, tm_w, tm_out, p);
    TU_CHECK_LAUNCH("upfold_stream");
    return TU_OK;
}

}  // namespace

// relu(up1_conv(PixelShuffle_r(up1_stage(in)))) with the folded filter f: NHWC bf16 (B,H,W,64) -> planar fp32 (B,3,rH,rW)
int tc_upfold(const bf16 *in, const TuUpFold *f, float *out, int B, int H, int W, cudaStream_t st) {
    if ((reinterpret_cast<uintptr_t>(in) & 127) || (reinterpret_cast<uintptr_t>(f->w) & 127) || (reinterpret_cast<uintptr_t>(out) & 15))
        return TU_TC_UNSUPPORTED;
    if (((long)W * f->r * 4) % 16) return TU_TC_UNSUPPORTED;       // TMA row pitch of the output image
    int rc;
    switch (f->r) {
        case 2: rc = launch_fold<16, 2, 5>(in, f->w, f->b, 1, out, B, H, W, st); break;
        case 3: rc = launch_fold<32, 3, 5>(in, f->w, f->b, 1, out, B, H, W, st); break;
        case 6: rc = launch_fold<48, 6, 5>(in, f->w, f->b, 1, out, B, H, W, st); break;
        default: return TU_TC_UNSUPPORTED;
    }
    if (rc) return rc;
    const int R = f->r;
    const long ring = (long)B * (2L * W * R + 2L * (H * R - 2));
    launch_pdl(upfold_ring_kernel, dim3((unsigned)((ring + 7) / 8)), dim3(256), 0, st, in, f->ring_w, f->ring_b, out, B, H, W, R);
    TU_CHECK_LAUNCH("upfold_ring");
    return TU_OK;
}

// 64 -> 3 head (decoder_conv2, up1_conv) on the same streaming kernel with a 3x3 filter and no PixelShuffle: wst = bf16
// (3 kx, 3 blocks holding ky = 2..0, 16 rows [co < 3, rest zero], 64 ci); bias16 = fp32 (16) or nullptr.  Needs W % 4 == 0.
int tc_conv3x3_c64_to3_stream(const bf16 *in, const bf16 *wst, const float *bias16, float *out, int B, int H, int W, int relu,
                              cudaStream_t st) {
    if (!wst || !bias16 || (reinterpret_cast<uintptr_t>(in) & 127) || (reinterpret_cast<uintptr_t>(wst) & 127) ||
        (reinterpret_cast<uintptr_t>(out) & 15) || (W & 3))
        return TU_TC_UNSUPPORTED;
    return launch_fold<16, 1, 3>(in, wst, bias16, relu, out, B, H, W, st);
}

}  // namespace tu
