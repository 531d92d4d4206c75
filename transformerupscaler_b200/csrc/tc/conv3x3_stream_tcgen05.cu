// 3x3 / pad-1 / stride-1 convolution 64 -> 64 channels, streaming down the image with the three VERTICAL taps stacked into N.
//
// Reference call sites: conv2 and decoder_conv1 (WindowTransformer/model.py:202,221; FastTransformer/model.py:204,228;
// ResidualTransformer/model.py:85,111) -- 54 % + 13 % of WindowTransformer's FLOPs.
//
// Why: the tap-by-tap kernel (conv3x3_tcgen05.cu) issues nine 128x64x16 MMAs per 16 input channels; each reads 4 KB of
// activations + 2 KB of filter from shared memory for 32 cycles of tensor work = 192 B/clk against a 128 B/clk shared
// memory port, and ncu shows exactly that (sm__throughput 88 %, tensor pipe 58 %; 9.0 k cycles per 4-row tile against an
// 8.7 k-cycle shared-memory bound).  An input row s feeds output rows s+1, s, s-1 through the taps ky = 0, 1, 2 with the
// SAME operand view, so one 128x192x16 MMA against the stacked filter [W(ky=2); W(ky=1); W(ky=0)] accumulates into the
// accumulators of three consecutive output rows at once: 4 KB + 6 KB of shared memory per 96 cycles = 104 B/clk, under
// the port limit.  No cross-lane work is needed (a TMEM lane is still one pixel), so the epilogue is a plain drain.
//
// Streaming: a work item is a column strip of 128 pixels x R output rows of one frame.  The row accumulators form a ring
// of eight 64-column slots in TMEM (all 512 columns); input row s accumulates into the window of slots of rows s-1..s+1,
// row y is complete after input row y+1 and is drained by the epilogue while the MMAs move on, so every input row is
// loaded once (plus two halo rows per item) and up to eight output rows are in flight.
//   warp 0   TMA producer: one input row segment (136 pixels from x0-1; zero-filled outside the image) per step
//   warp 1   MMA issuer: per input row 3 kx x 4 k-steps N<=192 MMAs (the first one split so that the newest row overwrites)
//   warps 4-7 epilogue: tcgen05.ld -> +bias, ReLU, bf16 -> swizzled staging -> TMA store, per output row
#include <cuda.h>
#include <string.h>

#include "ptx.cuh"
#include "tc_api.cuh"

namespace tu {

namespace {

constexpr int TILE_M = 128, BOXW = 136;
constexpr int UNIT_BYTES = BOXW * 128;        // 17408: one input row segment
constexpr int RING = 6;                       // input row slots
constexpr int NACC = 8;                       // output-row accumulators in TMEM (8 x 64 columns)
constexpr int W_BLK = 64 * 128;               // one (kx, ky) filter block: 64 co x 64 ci
constexpr int W_KX = 3 * W_BLK;               // per kx: [ky=2; ky=1; ky=0] stacked = 192 rows
constexpr int W_BYTES = 3 * W_KX;             // 73728
constexpr int STG_BYTES = 4 * 2 * 4096;       // 4 epilogue warps x 2 buffers x (32 pixels x 128 B)
constexpr int SMEM_BYTES = W_BYTES + RING * UNIT_BYTES + STG_BYTES + 512 + 1024;
constexpr int NUM_THREADS = 256;

struct StreamParams {
    int B, H, W, relu;
    int R;                  // output rows per work item
    int tiles_x, chunks_y, items_per_chunk, total_items;
    int ps_r;               // PixelShuffle factor (0 = plain): output chunk c holds sub-pixel phase (c / r, c % r)
    const float *bias;      // nchunk * 64, or nullptr
    int acc_bias;           // single bank: the bias vector is folded into the accumulator "clear" instead of being added in the drain
};

struct StreamBarriers {
    uint64_t full[RING], empty[RING];
    uint64_t acc_full[NACC], acc_empty[NACC];
    uint64_t w_full, w_free;
    uint32_t tmem_base;
};
static_assert(sizeof(StreamBarriers) <= 512, "barrier block too large");

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_stream_kernel(const __grid_constant__ CUtensorMap tmap_act, const __grid_constant__ CUtensorMap tmap_w,
                      const __grid_constant__ CUtensorMap tmap_out, const StreamParams p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_sm = smem0, ring_sm = smem0 + W_BYTES, stg_sm = ring_sm + RING * UNIT_BYTES;
    uint8_t *smem_al = smem_raw + (smem0 - ptx::smem_u32(smem_raw));
    StreamBarriers *bars = reinterpret_cast<StreamBarriers *>(smem_al + W_BYTES + RING * UNIT_BYTES + STG_BYTES);
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

    // work item -> (output-channel chunk, frame b, first output row y0, rows in the item, first pixel x0); items are ordered
    // chunk-major, so a CTA switches filter banks at most nchunk - 1 times
    auto item_geom = [&](int it, int &b, int &y0, int &rows, int &x0) -> int {
        const int chunk = it / p.items_per_chunk;
        it -= chunk * p.items_per_chunk;
        const int tx = it % p.tiles_x;
        int rem = it / p.tiles_x;
        const int cy = rem % p.chunks_y;
        b = rem / p.chunks_y;
        y0 = cy * p.R;
        rows = min(p.R, p.H - y0);
        x0 = tx * TILE_M;
        return chunk;
    };

    if (warp == 0 && lane == 0) {
        // ================================ TMA producer ================================
        // filter bank -> smem as [kx][ky = 2, 1, 0][co][ci]: block (kx, j) holds tap (ky = 2 - j, kx)
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
                for (int kx = 0; kx < 3; ++kx)
                    for (int j = 0; j < 3; ++j)
                        ptx::tma_load_2d(w_sm + kx * W_KX + j * W_BLK, &tmap_w, ptx::smem_u32(&bars->w_full), 0,
                                         (chunk * 9 + (2 - j) * 3 + kx) * 64);
                cur_chunk = chunk;
            }
            for (int u = 0; u < rows + 2; ++u) {          // input rows y0 - 1 .. y0 + rows
                ptx::mbar_wait(ptx::smem_u32(&bars->empty[slot]), phase ^ 1);
                const uint32_t fb = ptx::smem_u32(&bars->full[slot]);
                ptx::mbar_expect_tx(fb, UNIT_BYTES);
                ptx::tma_load_4d(ring_sm + slot * UNIT_BYTES, &tmap_act, fb, 0, x0 - 1, y0 - 1 + u, b);
                if (++slot == RING) { slot = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (whole warp converged, elected lane issues) ================================
        const uint32_t leader = ptx::elect_one();
        const uint32_t w_lo = ptx::sdesc_lo(w_sm), ring_lo = ptx::sdesc_lo(ring_sm);
        int slot = 0, cur_chunk = -1;
        uint32_t phase = 0, wfull_ph = 0;
        uint32_t g0 = 0;                                   // global index (per CTA) of the item's first output row
        for (int it = blockIdx.x; it < p.total_items; it += gridDim.x) {
            int b, y0, rows, x0;
            const int chunk = item_geom(it, b, y0, rows, x0);
            if (chunk != cur_chunk) {
                ptx::mbar_wait(ptx::smem_u32(&bars->w_full), wfull_ph);
                wfull_ph ^= 1;
                cur_chunk = chunk;
            }
            for (int u = 0; u < rows + 2; ++u) {
                // input row u - 1 (relative) feeds output rows lo..hi with ky = u - row; filter block of row m is 2 - (u - m).
                // Every accumulator slot is zero when a row opens (the epilogue clears it after draining), so all MMAs accumulate.
                const int lo = max(u - 2, 0), hi = min(u, rows - 1);
                ptx::mbar_wait(ptx::smem_u32(&bars->full[slot]), phase);
                if (u <= rows - 1) {                       // row `u` opens: its ring slot must have been drained and cleared
                    const uint32_t g = g0 + u;
                    ptx::mbar_wait(ptx::smem_u32(&bars->acc_empty[g & (NACC - 1)]), (g >> 3) & 1);
                }
                ptx::tc_fence_after();
                const uint32_t a_lo = ring_lo + ((uint32_t)(slot * UNIT_BYTES) >> 4);
                const int n = hi - lo + 1, blk0 = 2 - (u - lo);
                const int s0 = (g0 + lo) & (NACC - 1);
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
                if (u >= 2) ptx::umma_commit_pred(ptx::smem_u32(&bars->acc_full[(g0 + u - 2) & (NACC - 1)]), leader);   // row u-2 is complete
                if (++slot == RING) { slot = 0; phase ^= 1; }
            }
            g0 += rows;
            // does this CTA's next item use another filter bank?
            const int nx = it + gridDim.x;
            if (nx < p.total_items && nx / p.items_per_chunk != chunk) ptx::umma_commit_pred(ptx::smem_u32(&bars->w_free), leader);
        }
    } else if (warp >= 4) {
        // ================================ epilogue: one output row at a time ================================
        const int q = warp - 4;
        uint32_t g = 0, nstore = 0;
        // value a row accumulator is reset to after its drain: zero, or (single filter bank) the bias vector, which then never has
        // to be added in the drain (64 global loads per row and thread otherwise)
        uint32_t init0[32], init1[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            init0[c] = p.acc_bias ? __float_as_uint(__ldg(p.bias + c)) : 0u;
            init1[c] = p.acc_bias ? __float_as_uint(__ldg(p.bias + 32 + c)) : 0u;
        }
        for (int c = 0; c < 512; c += 64) {
            ptx::tmem_st_x32(tmem_base + ((uint32_t)(q * 32) << 16) + c, init0);
            ptx::tmem_st_x32(tmem_base + ((uint32_t)(q * 32) << 16) + c + 32, init1);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0)
            for (int sl = 0; sl < NACC; ++sl) ptx::mbar_arrive(ptx::smem_u32(&bars->acc_empty[sl]));
        uint8_t *stg_w = smem_al + W_BYTES + RING * UNIT_BYTES + q * 8192;
        const uint32_t stg_w_sm = stg_sm + q * 8192;
        for (int it = blockIdx.x; it < p.total_items; it += gridDim.x) {
            int b, y0, rows, x0;
            const int chunk = item_geom(it, b, y0, rows, x0);
            const int px0 = x0 + q * 32;
            const float *bias = (p.bias && !p.acc_bias) ? p.bias + chunk * 64 : nullptr;
#pragma unroll 1
            for (int m = 0; m < rows; ++m, ++g) {
                const int sl = g & (NACC - 1);
                ptx::mbar_wait(ptx::smem_u32(&bars->acc_full[sl]), (g >> 3) & 1);
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
                if (lane == 0) {
                    ptx::mbar_arrive(ptx::smem_u32(&bars->acc_empty[sl]));      // the row is in registers, the slot is zero
                    ptx::bulk_wait_read<1>();                                   // the store that last used this buffer has read it
                }
                __syncwarp();
                const uint32_t buf = nstore & 1;
                uint8_t *rowp = stg_w + buf * 4096 + lane * 128;
#pragma unroll
                for (int c = 0; c < 64; c += 8) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float a = __uint_as_float(c < 32 ? v0[c + e] : v1[c - 32 + e]) + (bias ? __ldg(bias + c + e) : 0.f);
                        f[e] = p.relu ? fmaxf(a, 0.f) : a;
                    }
                    uint4 uu;
                    __nv_bfloat162 h;
                    h = __floats2bfloat162_rn(f[0], f[1]); uu.x = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(f[2], f[3]); uu.y = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(f[4], f[5]); uu.z = *reinterpret_cast<uint32_t *>(&h);
                    h = __floats2bfloat162_rn(f[6], f[7]); uu.w = *reinterpret_cast<uint32_t *>(&h);
                    *reinterpret_cast<uint4 *>(rowp + ((((c >> 3) ^ (lane & 7))) << 4)) = uu;
                }
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if (px0 < p.W) {
                        if (p.ps_r) {
                            const int rr = p.ps_r;
                            ptx::tma_store_4d(&tmap_out, stg_w_sm + buf * 4096, 0, px0, chunk % rr, (b * p.H + y0 + m) * rr + chunk / rr);
                        } else {
                            ptx::tma_store_4d(&tmap_out, stg_w_sm + buf * 4096, 0, px0, y0 + m, b);
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

PerDeviceFlag g_attr_set_s;
thread_local int g_enable_stream = 1;

}  // namespace

void tc_set_conv_stream(int on) { g_enable_stream = on; }

// stride 1 only; w = [chunk][9 taps][64 co][64 ci] bf16 (the same banks the tap-by-tap kernel uses); nchunk > 1 needs ps_r
// with ps_r * ps_r == nchunk: chunk c is written as sub-pixel phase (c / r, c % r) of a (B, H r, W r, 64) tensor
int tc_conv3x3_c64_stream(const bf16 *in, const bf16 *w, const float *bias, bf16 *out, int B, int H, int W, int relu, int nchunk,
                          int ps_r, cudaStream_t st) {
    if (!g_enable_stream) return TU_TC_UNSUPPORTED;
    if (nchunk > 1 && (ps_r == 0 || ps_r * ps_r != nchunk)) return TU_TC_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(in) & 127) || (reinterpret_cast<uintptr_t>(w) & 127) || (reinterpret_cast<uintptr_t>(out) & 15))
        return TU_TC_UNSUPPORTED;
    TcEncodeFn enc = tc_encode_fn();
    if (!enc) return TU_TC_UNSUPPORTED;
    const int g_sm_count_s = device_sm_count();
    if (!g_attr_set_s.is_set()) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "conv3x3_stream smem attribute");
        g_attr_set_s.set();
    }
    CUtensorMap tm_act, tm_w, tm_out;
    {
        cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
        cuuint32_t box[4] = {64, (cuuint32_t)BOXW, 1, 1}, estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm_act, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t wd[2] = {64, (cuuint64_t)nchunk * 9 * 64}, ws[1] = {128};
        cuuint32_t wb[2] = {64, 64}, we[2] = {1, 1};
        if (r == CUDA_SUCCESS)
            r = enc(&tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)w, wd, ws, wb, we, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint32_t ob[4] = {64, 32, 1, 1};
        cuuint64_t od[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t os[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
        if (ps_r) {      // (c, px [stride r pixels], pj, (b*H + y)*r + pi): PixelShuffle addressing done by the TMA engine
            od[1] = (cuuint64_t)W; od[2] = (cuuint64_t)ps_r; od[3] = (cuuint64_t)B * H * ps_r;
            os[0] = (cuuint64_t)ps_r * 128; os[1] = 128; os[2] = (cuuint64_t)W * ps_r * 128;
        }
        if (r == CUDA_SUCCESS)
            r = enc(&tm_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)out, od, os, ob, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("tu: cuTensorMapEncodeTiled(conv stream) failed with code " + std::to_string((int)r));
            return TU_ERR_CUDA;
        }
    }
    StreamParams p;
    p.B = B; p.H = H; p.W = W; p.relu = relu;
    p.tiles_x = ceil_div(W, TILE_M);
    // rows per work item: tall items amortise the two halo rows, but the item count should fill whole waves of SMs
    int bestR = H < 8 ? H : 8;
    double best = 1e30;
    for (int R = 8; R <= 64 && R <= (H > 8 ? H : 8); ++R) {
        const long items = (long)p.tiles_x * ceil_div(H, R) * B * nchunk;
        const long waves = (items + g_sm_count_s - 1) / g_sm_count_s;
        const double cost = (double)waves * (R + 2);       // steps executed by the busiest SM
        if (cost < best - 1e-9) { best = cost; bestR = R; }
    }
    p.R = bestR;
    p.chunks_y = ceil_div(H, p.R);
    p.items_per_chunk = p.tiles_x * p.chunks_y * B;
    p.total_items = p.items_per_chunk * nchunk;
    p.ps_r = ps_r;
    p.bias = bias;
    p.acc_bias = (bias != nullptr && nchunk == 1) ? 1 : 0;
    const int grid = p.total_items < g_sm_count_s ? p.total_items : g_sm_count_s;
    launch_pdl(conv3x3_stream_kernel, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, st, tm_act, tm_w, tm_out, p);
    TU_CHECK_LAUNCH("conv3x3_stream");
    return TU_OK;
}

}  // namespace tu
