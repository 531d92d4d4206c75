// Memory-bound resampling kernels.
//  * fused  out = clamp( bicubic(x) + bicubic(residual) )  — replaces two F.interpolate(mode='bicubic',
//    align_corners=False) calls, the add and the clamp (WindowTransformer/model.py:241,301,304-305;
//    ResidualTransformer/model.py:125,160,163-164) with one pass: x and residual read once, out written once.
//    Tap arithmetic follows ATen bit-for-bit in fp32 (ATen/native/UpSample.h:259-312,400-448):
//    scale = (float)in/out, src = fma(scale, dst+0.5, -0.5), i = floor(src), t = src - i,
//    Keys cubic A = -0.75, taps clamped to [0, in-1].
//  * antialiased bilinear resize (torchvision Resize on a tensor = ATen _upsample_bilinear2d_aa), used by
//    FastTransformer when the integer factor overshoots res_out (FastTransformer/model.py:323-325).
#include <cuda.h>

#include <algorithm>

#include "tc/ptx.cuh"
#include "tc/tc_api.cuh"
#include "tu_common.cuh"

namespace tu {

struct Cubic {
    int idx[4];
    float w[4];
};

__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

__device__ __forceinline__ Cubic cubic_taps(int dst, int in_size, int out_size) {
    const float scale = (float)in_size / (float)out_size;
    const float src = fmaf(scale, (float)dst + 0.5f, -0.5f);
    int i0 = min((int)floorf(src), in_size - 1);
    float t = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
    const float A = -0.75f;
    Cubic c;
    c.w[0] = cubic2(t + 1.f, A);
    c.w[1] = cubic1(t, A);
    const float u = 1.f - t;
    c.w[2] = cubic1(u, A);
    c.w[3] = cubic2(u + 1.f, A);
#pragma unroll
    for (int j = 0; j < 4; ++j) c.idx[j] = min(max(i0 + j - 1, 0), in_size - 1);
    return c;
}

// A thread owns one output column of a 32-row strip and slides a 4-row window of horizontally-resampled source rows
// down the strip, for the 3 channels and both sources at once.  A source row is resampled horizontally exactly once
// per strip, the per-row vertical taps come from a small shared table computed once per CTA.  Operation order per
// output is ATen's: horizontal 4-tap sum inside each source row, then the vertical 4-tap sum.
//
// Source pixels come from one of two places:
//   * TMA variant (row pitch of both sources a multiple of 16 bytes): one thread issues two cp.async.bulk.tensor loads
//     that bring the whole source footprint of the tile (3 channels x rows x cols) into shared memory; out-of-image
//     parts are rewritten with the border pixel (replicate padding == clamped taps), so a thread's four taps are four
//     contiguous elements and the strip walk runs at shared-memory latency with almost no address arithmetic.
//   * direct variant (any shape): taps are read from global memory.
constexpr int BS_W = 128, BS_H = 32;

// element as stored -> float, and the factor applied once per 4-tap sum (uint8 frames: 1/255, ToTensor)
__device__ __forceinline__ float raw_f(float v) { return v; }
__device__ __forceinline__ float raw_f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float raw_f(uint8_t v) { return u8_to_float(v); }
template <typename TS> __device__ __forceinline__ float src_scale(float t) { return t; }
template <> __device__ __forceinline__ float src_scale<uint8_t>(float t) { return t * 0.00392156862745098f; }

template <typename TS>
struct GlobalSrc {          // image of one frame: (3, H, W); taps and rows clamped to the image like ATen
    const TS *img;
    long plane;
    int H, W;
    __device__ __forceinline__ float hrow(int ch, int row, const Cubic &cx) const {
        const TS *p = img + ch * plane + (long)min(max(row, 0), H - 1) * W;
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) t += raw_f(p[cx.idx[j]]) * cx.w[j];
        return src_scale<TS>(t);
    }
};
// slide the 3-channel window so that it covers source rows base-1 .. base+2
template <typename Src>
__device__ __forceinline__ void slide(float (&win)[3][4], int &cur, int base, const Src &src, const Cubic &cx) {
    if (cur != base) {
        const int step = base - cur;
        if (step < 0 || step > 3) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int i = 0; i < 4; ++i) win[c][i] = src.hrow(c, base - 1 + i, cx);
        } else {
            for (int s = 0; s < step; ++s) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    win[c][0] = win[c][1]; win[c][1] = win[c][2]; win[c][2] = win[c][3];
                    win[c][3] = src.hrow(c, cur + s + 3, cx);
                }
            }
        }
        cur = base;
    }
}

// The TMA zero-fills the parts of the box that lie outside the image; bicubic taps clamp to the border instead.
// Rewrite those columns / rows with the nearest image pixel so that a thread's four taps are always contiguous.
template <typename TS>
__device__ __forceinline__ void replicate_pad(TS *s, int r0, int c0, int nr, int nc, int H, int W) {
    const int padL = max(0, -c0), padR = max(0, c0 + nc - W), padT = max(0, -r0), padB = max(0, r0 + nr - H);
    if ((padL | padR | padT | padB) == 0) return;           // interior tile (uniform across the CTA)
    if (padL | padR) {
        for (int e = threadIdx.x; e < 3 * nr; e += blockDim.x) {
            TS *p = s + e * nc;
            const TS l = p[min(padL, nc - 1)], r = p[max(nc - 1 - padR, 0)];
            for (int c = 0; c < padL; ++c) p[c] = l;
            for (int c = nc - padR; c < nc; ++c) p[c] = r;
        }
        __syncthreads();
    }
    if (padT | padB) {
        for (int e = threadIdx.x; e < 3 * nc; e += blockDim.x) {
            TS *p = s + (e / nc) * nr * nc + (e % nc);
            const TS tp = p[min(padT, nr - 1) * nc], bt = p[max(nr - 1 - padB, 0) * nc];
            for (int r = 0; r < padT; ++r) p[r * nc] = tp;
            for (int r = nr - padB; r < nr; ++r) p[r * nc] = bt;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ int src_floor(int dst, int in_size, int out_size) {
    const float scale = (float)in_size / (float)out_size;
    return min((int)floorf(fmaf(scale, (float)dst + 0.5f, -0.5f)), in_size - 1);
}

struct BicubicTileGeom {
    int xr, xc, rr, rc;             // TMA box sizes: rows / cols of the x tile and of the residual tile
    float sxh, sxw, srh, srw;       // (float)in / (float)out per axis and source, divided once on the host (same IEEE quotient)
};
__device__ __forceinline__ int src_floor_s(int dst, float scale, int in_size) {
    return min((int)floorf(fmaf(scale, (float)dst + 0.5f, -0.5f)), in_size - 1);
}

// shared-memory element -> float with an immediate byte offset (the four taps of a row are contiguous)
template <int OFF> __device__ __forceinline__ float lds_raw(uint32_t a, float) {
    float v;
    asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF * 4));
    return v;
}
template <int OFF> __device__ __forceinline__ float lds_raw(uint32_t a, bf16) {
    unsigned short h;
    asm("ld.shared.u16 %0, [%1+%2];" : "=h"(h) : "r"(a), "n"(OFF * 2));
    return __uint_as_float((uint32_t)h << 16);
}
template <int OFF> __device__ __forceinline__ float lds_raw(uint32_t a, uint8_t) {
    unsigned short h;
    asm("ld.shared.u8 %0, [%1+%2];" : "=h"(h) : "r"(a), "n"(OFF));
    return u8_to_float(h);
}
// 4-tap horizontal sum of one source row of the replicate-padded tile (ATen's left-to-right order)
template <typename TS>
__device__ __forceinline__ float hrow_tile(uint32_t a, const float (&w)[4]) {
    float t = 0.f;
    t += lds_raw<0>(a, TS()) * w[0];
    t += lds_raw<1>(a, TS()) * w[1];
    t += lds_raw<2>(a, TS()) * w[2];
    t += lds_raw<3>(a, TS()) * w[3];
    return src_scale<TS>(t);
}
// slide a 3-channel window of horizontally resampled rows to source rows base-1 .. base+2; ca[] = per-channel shared
// address of the thread's first tap in tile row 0, trow0 = tile row of source row 0 (= -r0)
template <typename TS>
__device__ __forceinline__ void slide_tile(float (&win)[3][4], int &cur, int base, const uint32_t (&ca)[3], int trow0, uint32_t pitchB,
                                           const float (&w)[4]) {
    if (base == cur) return;
    if (base == cur + 1) {                  // the only step an up-scaling strip ever takes after its first row
        const uint32_t ro = (uint32_t)(base + 2 + trow0) * pitchB;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            win[c][0] = win[c][1]; win[c][1] = win[c][2]; win[c][2] = win[c][3];
            win[c][3] = hrow_tile<TS>(ca[c] + ro, w);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t ro = (uint32_t)(base - 1 + i + trow0) * pitchB;
#pragma unroll
            for (int c = 0; c < 3; ++c) win[c][i] = hrow_tile<TS>(ca[c] + ro, w);
        }
    }
    cur = base;
}

// MODE 0: taps from global memory (any shape, residual optional at run time); MODE 1 / 2: TMA-staged tile with / without
// the residual source (compile-time, so the strip loop has no dead branches)
template <typename TI, typename TO, int MODE>
__global__ void __launch_bounds__(BS_W) bicubic_add_clamp_strip_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                                       const __grid_constant__ CUtensorMap tmap_r,
                                                                       const BicubicTileGeom g, const TI *__restrict__ x, int H, int W,
                                                                       const float *__restrict__ res, int rH, int rW,
                                                                       TO *__restrict__ out, int oH, int oW, int clamp, int layout) {
    pdl_trigger();
    pdl_wait();          // the residual image is written by the previous kernel of the stream
    __shared__ float yw[2][BS_H][4];
    __shared__ int ybase[2][BS_H];
    __shared__ __align__(8) uint64_t bar;
    extern __shared__ uint8_t tile_dyn[];
    uint8_t *tile_raw = tile_dyn + ((128u - (ptx::smem_u32(tile_dyn) & 127u)) & 127u);     // TMA destinations are 128-byte aligned
    constexpr bool TMA = MODE != 0;
    if (MODE == 2) res = nullptr;
    const int t = threadIdx.x;
    const int ox0 = blockIdx.x * BS_W, ox = ox0 + t, oy0 = blockIdx.y * BS_H, b = blockIdx.z;
    int xr0 = 0, xc0 = 0, rr0 = 0, rc0 = 0;
    const uint32_t x_bytes = 3u * g.xr * g.xc * sizeof(TI);
    const uint32_t x_bytes_al = (x_bytes + 127u) & ~127u;
    if (TMA) {
        // the innermost box coordinate must start on a 16-byte boundary (measured: tools/probes/tma_probe.cu): round the
        // first column down to a multiple of 16 / sizeof(element); the box is that much wider (see the host side)
        constexpr int XA = 16 / (int)sizeof(TI);
        xr0 = src_floor(oy0, H, oH) - 1; xc0 = (src_floor(ox0, W, oW) - 1) & ~(XA - 1);
        if (res) { rr0 = src_floor(oy0, rH, oH) - 1; rc0 = (src_floor(ox0, rW, oW) - 1) & ~3; }
        if (t == 0) {
            const uint32_t bar_a = ptx::smem_u32(&bar);
            ptx::mbar_init(bar_a, 1);
            ptx::fence_barrier_init();
            ptx::mbar_expect_tx(bar_a, x_bytes + (res ? 3u * g.rr * g.rc * 4u : 0u));
            ptx::tma_load_4d(ptx::smem_u32(tile_raw), &tmap_x, bar_a, xc0, xr0, 0, b);
            if (res) ptx::tma_load_4d(ptx::smem_u32(tile_raw) + x_bytes_al, &tmap_r, bar_a, rc0, rr0, 0, b);
        }
    }
    if (t < 2 * BS_H) {
        const int s = t / BS_H, r = t % BS_H;
        const int in_size = s ? rH : H;
        const float scale = (float)in_size / (float)oH;
        const float src = fmaf(scale, (float)min(oy0 + r, oH - 1) + 0.5f, -0.5f);
        const int i0 = min((int)floorf(src), in_size - 1);
        const float tt = fminf(fmaxf(src - (float)i0, 0.f), 1.f), u = 1.f - tt;
        const float A = -0.75f;
        ybase[s][r] = i0;
        yw[s][r][0] = cubic2(tt + 1.f, A); yw[s][r][1] = cubic1(tt, A); yw[s][r][2] = cubic1(u, A); yw[s][r][3] = cubic2(u + 1.f, A);
    }
    __syncthreads();
    if (TMA) {
        ptx::mbar_wait(ptx::smem_u32(&bar), 0);
        replicate_pad(reinterpret_cast<TI *>(tile_raw), xr0, xc0, g.xr, g.xc, H, W);
        if (res) replicate_pad(reinterpret_cast<float *>(tile_raw + x_bytes_al), rr0, rc0, g.rr, g.rc, rH, rW);
    }
    if (ox >= oW) return;
    const Cubic cx = cubic_taps(ox, W, oW);
    Cubic cr = cx;
    if (res) cr = cubic_taps(ox, rW, oW);
    const long xplane = (long)H * W, rplane = (long)rH * rW, oplane = (long)oH * oW;
    TO *ob = out + (long)b * 3 * oplane + ox;
    float wx[3][4], wr[3][4];
    int curx = -1000000, curr = -1000000;
    const int nrows = min(BS_H, oH - oy0);
    auto run = [&](const auto &sx, const auto &sr) {
        for (int r = 0; r < nrows; ++r) {
            slide(wx, curx, ybase[0][r], sx, cx);
            if (res) slide(wr, curr, ybase[1][r], sr, cr);
            const float4 a = *reinterpret_cast<const float4 *>(yw[0][r]);
            const float4 gg = *reinterpret_cast<const float4 *>(yw[1][r]);
#pragma unroll
            float v3[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v = 0.f;
                v += wx[c][0] * a.x; v += wx[c][1] * a.y; v += wx[c][2] * a.z; v += wx[c][3] * a.w;
                if (res) {
                    float u = 0.f;
                    u += wr[c][0] * gg.x; u += wr[c][1] * gg.y; u += wr[c][2] * gg.z; u += wr[c][3] * gg.w;
                    v += u;
                }
                if (clamp) v = fminf(fmaxf(v, 0.f), 1.f);
                v3[c] = v;
            }
            store_rgb<TO>(out, b, oplane, (long)(oy0 + r) * oW + ox, layout, v3[0], v3[1], v3[2]);
        }
    };
    if (TMA) {
        // ---- fast path: explicit shared-memory addresses, immediate tap offsets, running output pointers
        constexpr bool RES = MODE == 1;
        const uint32_t xpitch = g.xc * sizeof(TI), rpitch = g.rc * 4u;
        const uint32_t xs = ptx::smem_u32(tile_raw) + (uint32_t)(src_floor(ox, W, oW) - 1 - xc0) * (uint32_t)sizeof(TI);
        const uint32_t xa[3] = {xs, xs + g.xr * xpitch, xs + 2u * g.xr * xpitch};
        uint32_t ra[3] = {0u, 0u, 0u};
        if (RES) {
            const uint32_t rs = ptx::smem_u32(tile_raw) + x_bytes_al + (uint32_t)(src_floor(ox, rW, oW) - 1 - rc0) * 4u;
            ra[0] = rs; ra[1] = rs + g.rr * rpitch; ra[2] = rs + 2u * g.rr * rpitch;
        }
        TO *o0 = ob + (long)oy0 * oW, *o1 = o0 + oplane, *o2 = o1 + oplane;
        for (int r = 0; r < nrows; ++r) {
            slide_tile<TI>(wx, curx, ybase[0][r], xa, -xr0, xpitch, cx.w);
            if (RES) slide_tile<float>(wr, curr, ybase[1][r], ra, -rr0, rpitch, cr.w);
            const float4 a = *reinterpret_cast<const float4 *>(yw[0][r]);
            const float4 gg = *reinterpret_cast<const float4 *>(yw[1][r]);
            float v[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                v[c] = 0.f;
                v[c] += wx[c][0] * a.x; v[c] += wx[c][1] * a.y; v[c] += wx[c][2] * a.z; v[c] += wx[c][3] * a.w;
                if (RES) {
                    float u = 0.f;
                    u += wr[c][0] * gg.x; u += wr[c][1] * gg.y; u += wr[c][2] * gg.z; u += wr[c][3] * gg.w;
                    v[c] += u;
                }
                if (clamp) v[c] = fminf(fmaxf(v[c], 0.f), 1.f);
            }
            if (layout == 0) {
                *o0 = from_f<TO>(v[0]); *o1 = from_f<TO>(v[1]); *o2 = from_f<TO>(v[2]);
                o0 += oW; o1 += oW; o2 += oW;
            } else {
                store_rgb<TO>(out, b, oplane, (long)(oy0 + r) * oW + ox, layout, v[0], v[1], v[2]);
            }
        }
    } else {
        const GlobalSrc<TI> sx{x + (long)b * 3 * xplane, xplane, H, W};
        const GlobalSrc<float> sr{res ? res + (long)b * 3 * rplane : nullptr, rplane, rH, rW};
        run(sx, sr);
    }
}

// ---- two output columns per thread -----------------------------------------------------------------------------------
// The strip kernel above is bound by instruction issue (145 instructions per output pixel, 81 % issue slots, 29 % of the HBM
// bandwidth).  Here a thread owns TWO adjacent output columns: the horizontal and vertical 4-tap sums of the pair run on packed
// fp32 pairs (FFMA2: one issue slot for both columns, bit-identical to the scalar FMAs in the same order), the per-row vertical
// weights, row bookkeeping and window shifts are shared by the pair, and the pair leaves as one 4-byte (bf16) / 8-byte (fp32) /
// 2-byte (uint8) store.  TMA-staged tiles only (PAIR_H x 128 outputs per 64-thread CTA), even output width.
constexpr int PAIR_H = 36;

template <typename TS>
__device__ __forceinline__ ptx::f32x2 hrow_pair(uint32_t aA, uint32_t aB, const ptx::f32x2 (&w)[4]) {
    ptx::f32x2 t = ptx::mul2(ptx::pk2(lds_raw<0>(aA, TS()), lds_raw<0>(aB, TS())), w[0]);
    t = ptx::fma2(ptx::pk2(lds_raw<1>(aA, TS()), lds_raw<1>(aB, TS())), w[1], t);
    t = ptx::fma2(ptx::pk2(lds_raw<2>(aA, TS()), lds_raw<2>(aB, TS())), w[2], t);
    t = ptx::fma2(ptx::pk2(lds_raw<3>(aA, TS()), lds_raw<3>(aB, TS())), w[3], t);
    if (sizeof(TS) == 1) t = ptx::mul2(t, ptx::pk2(0.00392156862745098f, 0.00392156862745098f));
    return t;
}
template <typename TS>
__device__ __forceinline__ void slide_pair(ptx::f32x2 (&win)[3][4], int &cur, int base, const uint32_t (&caA)[3], const uint32_t (&caB)[3],
                                           int trow0, uint32_t pitchB, const ptx::f32x2 (&w)[4]) {
    if (base == cur) return;
    if (base == cur + 1) {                  // the only step an up-scaling strip ever takes after its first row
        const uint32_t ro = (uint32_t)(base + 2 + trow0) * pitchB;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            win[c][0] = win[c][1]; win[c][1] = win[c][2]; win[c][2] = win[c][3];
            win[c][3] = hrow_pair<TS>(caA[c] + ro, caB[c] + ro, w);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t ro = (uint32_t)(base - 1 + i + trow0) * pitchB;
#pragma unroll
            for (int c = 0; c < 3; ++c) win[c][i] = hrow_pair<TS>(caA[c] + ro, caB[c] + ro, w);
        }
    }
    cur = base;
}
__device__ __forceinline__ void store_pair(float *o, float a, float b) { *reinterpret_cast<float2 *>(o) = make_float2(a, b); }
__device__ __forceinline__ void store_pair(bf16 *o, float a, float b) { *reinterpret_cast<__nv_bfloat162 *>(o) = __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ void store_pair(uint8_t *o, float a, float b) {
    *reinterpret_cast<unsigned short *>(o) = (unsigned short)((unsigned)from_f<uint8_t>(a) | ((unsigned)from_f<uint8_t>(b) << 8));
}

template <typename TI, typename TO, bool RES>
__global__ void __launch_bounds__(BS_W / 2) bicubic_add_clamp_pair_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                                          const __grid_constant__ CUtensorMap tmap_r,
                                                                          const BicubicTileGeom g, int H, int W, int rH, int rW,
                                                                          TO *__restrict__ out, int oH, int oW, int clamp, int layout) {
    pdl_trigger();
    pdl_wait();          // the residual image is written by the previous kernel of the stream
    __shared__ __align__(16) float yw[2][PAIR_H][4];
    __shared__ int ybase[2][PAIR_H];
    __shared__ __align__(8) uint64_t bar;
    extern __shared__ uint8_t tile_dyn[];
    uint8_t *tile_raw = tile_dyn + ((128u - (ptx::smem_u32(tile_dyn) & 127u)) & 127u);     // TMA destinations are 128-byte aligned
    const int t = threadIdx.x;
    const int ox0 = blockIdx.x * BS_W, ox = ox0 + 2 * t, oy0 = blockIdx.y * PAIR_H, b = blockIdx.z;
    const uint32_t x_bytes = 3u * g.xr * g.xc * sizeof(TI);
    const uint32_t x_bytes_al = (x_bytes + 127u) & ~127u;
    constexpr int XA = 16 / (int)sizeof(TI);       // the innermost box coordinate must start on a 16-byte boundary
    const int xr0 = src_floor(oy0, H, oH) - 1, xc0 = (src_floor(ox0, W, oW) - 1) & ~(XA - 1);
    int rr0 = 0, rc0 = 0;
    if (RES) { rr0 = src_floor(oy0, rH, oH) - 1; rc0 = (src_floor(ox0, rW, oW) - 1) & ~3; }
    if (t == 0) {
        const uint32_t bar_a = ptx::smem_u32(&bar);
        ptx::mbar_init(bar_a, 1);
        ptx::fence_barrier_init();
        ptx::mbar_expect_tx(bar_a, x_bytes + (RES ? 3u * g.rr * g.rc * 4u : 0u));
        ptx::tma_load_4d(ptx::smem_u32(tile_raw), &tmap_x, bar_a, xc0, xr0, 0, b);
        if (RES) ptx::tma_load_4d(ptx::smem_u32(tile_raw) + x_bytes_al, &tmap_r, bar_a, rc0, rr0, 0, b);
    }
    for (int i = t; i < 2 * PAIR_H; i += BS_W / 2) {
        const int s = i / PAIR_H, r = i % PAIR_H;
        const int in_size = s ? rH : H;
        const float scale = (float)in_size / (float)oH;
        const float src = fmaf(scale, (float)min(oy0 + r, oH - 1) + 0.5f, -0.5f);
        const int i0 = min((int)floorf(src), in_size - 1);
        const float tt = fminf(fmaxf(src - (float)i0, 0.f), 1.f), u = 1.f - tt;
        const float A = -0.75f;
        ybase[s][r] = i0;
        yw[s][r][0] = cubic2(tt + 1.f, A); yw[s][r][1] = cubic1(tt, A); yw[s][r][2] = cubic1(u, A); yw[s][r][3] = cubic2(u + 1.f, A);
    }
    __syncthreads();
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    replicate_pad(reinterpret_cast<TI *>(tile_raw), xr0, xc0, g.xr, g.xc, H, W);
    if (RES) replicate_pad(reinterpret_cast<float *>(tile_raw + x_bytes_al), rr0, rc0, g.rr, g.rc, rH, rW);
    if (ox >= oW) return;           // oW is even: a pair is inside or outside as a whole
    const Cubic cA = cubic_taps(ox, W, oW), cB = cubic_taps(ox + 1, W, oW);
    ptx::f32x2 wxp[4], wrp[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) wxp[j] = ptx::pk2(cA.w[j], cB.w[j]);
    const uint32_t xpitch = g.xc * sizeof(TI), rpitch = g.rc * 4u;
    const uint32_t tile_a = ptx::smem_u32(tile_raw);
    const uint32_t xsA = tile_a + (uint32_t)(src_floor(ox, W, oW) - 1 - xc0) * (uint32_t)sizeof(TI);
    const uint32_t xsB = tile_a + (uint32_t)(src_floor(ox + 1, W, oW) - 1 - xc0) * (uint32_t)sizeof(TI);
    const uint32_t xaA[3] = {xsA, xsA + g.xr * xpitch, xsA + 2u * g.xr * xpitch};
    const uint32_t xaB[3] = {xsB, xsB + g.xr * xpitch, xsB + 2u * g.xr * xpitch};
    uint32_t raA[3] = {0u, 0u, 0u}, raB[3] = {0u, 0u, 0u};
    if (RES) {
        const Cubic rA = cubic_taps(ox, rW, oW), rB = cubic_taps(ox + 1, rW, oW);
#pragma unroll
        for (int j = 0; j < 4; ++j) wrp[j] = ptx::pk2(rA.w[j], rB.w[j]);
        const uint32_t rsA = tile_a + x_bytes_al + (uint32_t)(src_floor(ox, rW, oW) - 1 - rc0) * 4u;
        const uint32_t rsB = tile_a + x_bytes_al + (uint32_t)(src_floor(ox + 1, rW, oW) - 1 - rc0) * 4u;
        raA[0] = rsA; raA[1] = rsA + g.rr * rpitch; raA[2] = rsA + 2u * g.rr * rpitch;
        raB[0] = rsB; raB[1] = rsB + g.rr * rpitch; raB[2] = rsB + 2u * g.rr * rpitch;
    }
    const long oplane = (long)oH * oW;
    TO *o0 = out + (long)b * 3 * oplane + (long)oy0 * oW + ox, *o1 = o0 + oplane, *o2 = o1 + oplane;
    ptx::f32x2 wx[3][4], wr[3][4];
    int curx = -1000000, curr = -1000000;
    const int nrows = min(PAIR_H, oH - oy0);
    for (int r = 0; r < nrows; ++r) {
        slide_pair<TI>(wx, curx, ybase[0][r], xaA, xaB, -xr0, xpitch, wxp);
        if (RES) slide_pair<float>(wr, curr, ybase[1][r], raA, raB, -rr0, rpitch, wrp);
        const float4 a = *reinterpret_cast<const float4 *>(yw[0][r]);
        const float4 gg = *reinterpret_cast<const float4 *>(yw[1][r]);
        const ptx::f32x2 a0 = ptx::pk2(a.x, a.x), a1 = ptx::pk2(a.y, a.y), a2 = ptx::pk2(a.z, a.z), a3 = ptx::pk2(a.w, a.w);
        float lo[3], hi[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            ptx::f32x2 v = ptx::mul2(wx[c][0], a0);
            v = ptx::fma2(wx[c][1], a1, v); v = ptx::fma2(wx[c][2], a2, v); v = ptx::fma2(wx[c][3], a3, v);
            if (RES) {
                ptx::f32x2 u = ptx::mul2(wr[c][0], ptx::pk2(gg.x, gg.x));
                u = ptx::fma2(wr[c][1], ptx::pk2(gg.y, gg.y), u); u = ptx::fma2(wr[c][2], ptx::pk2(gg.z, gg.z), u);
                u = ptx::fma2(wr[c][3], ptx::pk2(gg.w, gg.w), u);
                v = ptx::add2(v, u);
            }
            ptx::up2(v, lo[c], hi[c]);
            if (clamp) { lo[c] = fminf(fmaxf(lo[c], 0.f), 1.f); hi[c] = fminf(fmaxf(hi[c], 0.f), 1.f); }
        }
        if (layout == 0) {
            store_pair(o0, lo[0], hi[0]); store_pair(o1, lo[1], hi[1]); store_pair(o2, lo[2], hi[2]);
            o0 += oW; o1 += oW; o2 += oW;
        } else {
            // interleaved frames: the pair's two pixels are six consecutive elements (ox is even: 2-element aligned)
            const int c0 = layout == 2 ? 2 : 0, c2 = 2 - c0;
            TO *o = out + (((long)b * oH + oy0 + r) * oW + ox) * 3;
            store_pair(o, lo[c0], lo[1]); store_pair(o + 2, lo[c2], hi[c0]); store_pair(o + 4, hi[1], hi[c2]);
        }
    }
}

// pair store of values already clamped to [0, 1]: from_f<uint8_t> without its (then no-op) clamp to [0, 255] — the same product
// and the same round-down add, the two result bytes packed by one permute; other output types store as usual
template <typename T> __device__ __forceinline__ void store_pair_unit(T *o, float a, float b) { store_pair(o, a, b); }
template <> __device__ __forceinline__ void store_pair_unit<uint8_t>(uint8_t *o, float a, float b) {
    const uint32_t ua = __float_as_uint(__fadd_rd(__fmul_rn(a, 255.f), 8388608.f)), ub = __float_as_uint(__fadd_rd(__fmul_rn(b, 255.f), 8388608.f));
    *reinterpret_cast<unsigned short *>(o) = (unsigned short)__byte_perm(ua, ub, 0x0040);
}

// clamp a pair to [0, 1] and store it.  bf16: the conversion clamps below (cvt.rn.relu) and one packed min clamps above — rounding is
// monotonic and 1.0 is a bf16, so min(bf16(relu(v)), 1) == bf16(clamp(v, 0, 1)) bit for bit — which takes the two saturating adds off
// the FMA pipe, the busy one in these kernels; other types: saturate, then store
template <typename T> __device__ __forceinline__ void clamp_store_pair(T *o, float lo, float hi) {
    store_pair_unit(o, fminf(fmaxf(lo, 0.f), 1.f), fminf(fmaxf(hi, 0.f), 1.f));
}
template <> __device__ __forceinline__ void clamp_store_pair<bf16>(bf16 *o, float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    asm("min.bf16x2 %0, %0, %1;" : "+r"(d) : "r"(0x3f803f80u));
    *reinterpret_cast<uint32_t *>(o) = d;
}

// one value already clamped to [0, 1] -> the output element (uint8: the same product and round-down add as from_f<uint8_t>, whose clamp
// is a no-op here; the low byte of the sum is the result)
template <typename T> __device__ __forceinline__ T unit_to_u8(float v) { return from_f<T>(v); }
template <> __device__ __forceinline__ uint8_t unit_to_u8<uint8_t>(float v) {
    return (uint8_t)__float_as_uint(__fadd_rd(__fmul_rn(v, 255.f), 8388608.f));
}

// ---- periodic row schedules: x at 3:2 with the residual at 3:1 (720p -> 1080p, every x1.5 output), x at n:1 with the residual at 2n:1
// for n = 2, 3, 4, 6 (720p -> 4K is n = 3) ----------------------------------------------------------------------------------------
// The pair kernel spends more than half of its issue slots on bookkeeping: window moves (28 MOVs per source row), per-row
// compare / branch chains, 16-bit tap loads with one shift each, and a local-memory round trip of the pixel values.  When
// outH = 3/2 H = 3 rH (or n H = 2n rH) the source-row schedule is periodic (RowSched below; checked on the host with the same fp32
// coordinate arithmetic, row_pattern), so the strip loop unrolls over one period of output rows (12, or 8n)
// with the four window rows in FIXED registers (a circular window: no moves, no compares).  Taps are fetched as aligned words:
// the two columns of a pair need at most six consecutive source elements starting at an even index (three 32-bit loads for
// bf16, three 16-bit loads for uint8) against 6-entry weight vectors padded with zeros — fma(e, 0, t) == t, so every sum is
// bitwise the pair kernel's.  Vertical weights enter the packed FMAs as broadcast scalars; a thread works on one channel.
// Planar output, clamp on, outH a multiple of the period (a CTA's rows are whole unrolled periods).

template <typename TS> __device__ __forceinline__ void ld6(uint32_t a, float (&e)[6]);
template <> __device__ __forceinline__ void ld6<bf16>(uint32_t a, float (&e)[6]) {
    uint32_t w0, w1, w2;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(a));
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(a));
    asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(a));
    // low half by a byte permute (ALU pipe; a shift may be issued as an integer multiply on the FMA pipe, which is the busy one here)
    e[0] = __uint_as_float(__byte_perm(w0, 0u, 0x1044)); e[1] = __uint_as_float(w0 & 0xffff0000u);
    e[2] = __uint_as_float(__byte_perm(w1, 0u, 0x1044)); e[3] = __uint_as_float(w1 & 0xffff0000u);
    e[4] = __uint_as_float(__byte_perm(w2, 0u, 0x1044)); e[5] = __uint_as_float(w2 & 0xffff0000u);
}
template <> __device__ __forceinline__ void ld6<uint8_t>(uint32_t a, float (&e)[6]) {
    unsigned short h0, h1, h2;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h0) : "r"(a));
    asm volatile("ld.shared.u16 %0, [%1+2];" : "=h"(h1) : "r"(a));
    asm volatile("ld.shared.u16 %0, [%1+4];" : "=h"(h2) : "r"(a));
    // byte -> 0x4B0000bb (2^23 + b) with one byte permute, minus 2^23: u8_to_float
    e[0] = __uint_as_float(__byte_perm(h0, 0x4B000000u, 0x7440)) - 8388608.f; e[1] = __uint_as_float(__byte_perm(h0, 0x4B000000u, 0x7441)) - 8388608.f;
    e[2] = __uint_as_float(__byte_perm(h1, 0x4B000000u, 0x7440)) - 8388608.f; e[3] = __uint_as_float(__byte_perm(h1, 0x4B000000u, 0x7441)) - 8388608.f;
    e[4] = __uint_as_float(__byte_perm(h2, 0x4B000000u, 0x7440)) - 8388608.f; e[5] = __uint_as_float(__byte_perm(h2, 0x4B000000u, 0x7441)) - 8388608.f;
}
// horizontal sums of a column pair over one staged source row: six elements from an aligned address, zero-padded weights
template <typename TS>
__device__ __forceinline__ ptx::f32x2 hsum6(uint32_t a, const ptx::f32x2 (&w)[6]) {
    float e[6];
    ld6<TS>(a, e);
    ptx::f32x2 t = ptx::mul2(w[0], ptx::pk2(e[0], e[0]));
#pragma unroll
    for (int j = 1; j < 6; ++j) t = ptx::fma2(w[j], ptx::pk2(e[j], e[j]), t);
    if (sizeof(TS) == 1) t = ptx::mul2(t, ptx::pk2(0.00392156862745098f, 0.00392156862745098f));
    return t;
}
// the same over a fp32 row (residual): five elements from the first tap of the left column
__device__ __forceinline__ ptx::f32x2 hsum5(uint32_t a, const ptx::f32x2 (&w)[5]) {
    float e[5];
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(e[0]) : "r"(a));
    asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(e[1]) : "r"(a));
    asm volatile("ld.shared.f32 %0, [%1+8];" : "=f"(e[2]) : "r"(a));
    asm volatile("ld.shared.f32 %0, [%1+12];" : "=f"(e[3]) : "r"(a));
    asm volatile("ld.shared.f32 %0, [%1+16];" : "=f"(e[4]) : "r"(a));
    ptx::f32x2 t = ptx::mul2(w[0], ptx::pk2(e[0], e[0]));
#pragma unroll
    for (int j = 1; j < 5; ++j) t = ptx::fma2(w[j], ptx::pk2(e[j], e[j]), t);
    return t;
}
// w[k] of a 4-tap filter, zero outside 0..3 (k is a compile-time constant after unrolling: no branches, selects only)
__device__ __forceinline__ float tap_or_zero(const float (&w)[4], int k) { return (k >= 0 && k < 4) ? w[k] : 0.f; }

// A thread is (channel, column pair): 192 threads share the staged tile of a 36 x 128 output block — three times the warps of a
// three-channels-per-thread layout for the same shared memory, which is what hides the FMA-chain and shared-memory latencies
// (measured: 2 warps per CTA, 12 per SM -> issue slots 50 % busy).  The per-column filters (four cubic_taps per pair: two
// columns x two sources) are computed once per CTA by 256 tasks spread over the threads and handed over in shared memory, so
// the three channels do not repeat them.
constexpr int R32_T = 3 * (BS_W / 2);

// Row schedules of the tile kernel: over a period of P output rows (a multiple of the 4-row window rotation for both sources), does
// x / the residual step to a new source row before output row q?  PAT 0: outH = 3/2 H = 3 rH (720p -> 1080p; x 3:2, residual 3:1).
// PAT n = 2, 3, 4, 6: outH = n H = 2n rH (inference.py's --scale values on a 2x-downsampled residual; 3 = 720p -> 4K): the source
// coordinate (2 oy + 1 - n) / (2n) passes an integer before oy = n m + n / 2 (n even; n odd: AT oy = n m + (n - 1) / 2, where the
// fp32 rounding decides — the host checks every row), and likewise at ratio 2n for the residual.  A CTA covers whole periods of rows:
// TILE, or BIG when the larger tiles still give every SM a CTA (the per-tile set-up — column filters, vertical
// table, four window rows per source — is 30 % of the instructions at TILE rows; measured at 3:2 / 3:1, profiles/r3_ab_bicubic_tile_rows.log:
// 36 / 48 / 60 / 72 rows = 0.574 / 0.550 / 0.550 / 0.575 of the pair kernel's time, 72 rows leave three CTAs per SM).
template <int PAT> struct RowSched {
    static constexpr int P = 8 * PAT, TILE = PAT == 2 ? 48 : P, BIG = PAT == 2 ? 80 : PAT == 3 ? 72 : 96;
    static constexpr bool xstep(int q) { return q % PAT == PAT / 2; }
    static constexpr bool rstep(int q) { return q % (2 * PAT) == PAT; }
};
template <> struct RowSched<0> {
    static constexpr int P = 12, TILE = 36, BIG = 60;
    static constexpr bool xstep(int q) { return q % 3 != 0; }
    static constexpr bool rstep(int q) { return q % 3 == 1; }
};
template <int PAT> constexpr int sched_relx(int q) { int n = 0; for (int i = 0; i <= q; ++i) n += RowSched<PAT>::xstep(i) ? 1 : 0; return n; }
template <int PAT> constexpr int sched_relr(int q) { int n = 0; for (int i = 0; i <= q; ++i) n += RowSched<PAT>::rstep(i) ? 1 : 0; return n; }
static_assert(sched_relx<0>(11) == 8 && sched_relr<0>(11) == 4 && sched_relx<2>(15) == 8 && sched_relr<2>(15) == 4 &&
                  sched_relx<3>(23) == 8 && sched_relr<3>(23) == 4 && sched_relx<4>(31) == 8 && sched_relr<4>(31) == 4 &&
                  sched_relx<6>(47) == 8 && sched_relr<6>(47) == 4,
              "a period must rotate both 4-row windows a whole number of times");
constexpr int R32_MAXTILE = 96;
static inline int sched_tile(int pat, bool big) {
    switch (pat) {
        case 0: return big ? RowSched<0>::BIG : RowSched<0>::TILE;
        case 2: return big ? RowSched<2>::BIG : RowSched<2>::TILE;
        case 3: return big ? RowSched<3>::BIG : RowSched<3>::TILE;
        case 4: return big ? RowSched<4>::BIG : RowSched<4>::TILE;
        default: return big ? RowSched<6>::BIG : RowSched<6>::TILE;
    }
}

// HWC (uint8 frames only): interleaved output pixels, `layout` 1 = RGB, 2 = BGR; a thread stores its channel's two bytes
template <typename TI, typename TO, int PAT, bool HWC>
__global__ void __launch_bounds__(R32_T, 5) bicubic_add_clamp_r32_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                                         const __grid_constant__ CUtensorMap tmap_r,
                                                                         const BicubicTileGeom g, int H, int W, int rH, int rW,
                                                                         TO *__restrict__ out, int oH, int oW, int layout, int TILE) {
    pdl_trigger();
    pdl_wait();          // the residual image is written by the previous kernel of the stream
    constexpr int P = RowSched<PAT>::P;             // TILE (rows per CTA, a multiple of P, <= R32_MAXTILE) is chosen per launch
    __shared__ __align__(16) float yw[2][R32_MAXTILE][4];
    __shared__ float tapw[4][4][BS_W / 2];       // [x left, x right, residual left, residual right][tap][pair]
    __shared__ int tapi[4][BS_W / 2];            // source column of the second tap (floor of the source coordinate)
    __shared__ __align__(8) uint64_t bar;
    extern __shared__ uint8_t tile_dyn[];
    uint8_t *tile_raw = tile_dyn + ((128u - (ptx::smem_u32(tile_dyn) & 127u)) & 127u);     // TMA destinations are 128-byte aligned
    const int t = threadIdx.x;
    const int ox0 = blockIdx.x * BS_W, oy0 = blockIdx.y * TILE, b = blockIdx.z;
    const uint32_t x_bytes = 3u * g.xr * g.xc * sizeof(TI);
    const uint32_t x_bytes_al = (x_bytes + 127u) & ~127u;
    constexpr int XA = 16 / (int)sizeof(TI);       // the innermost box coordinate must start on a 16-byte boundary
    const int xr0 = src_floor_s(oy0, g.sxh, H) - 1, xc0 = (src_floor_s(ox0, g.sxw, W) - 1) & ~(XA - 1);
    const int rr0 = src_floor_s(oy0, g.srh, rH) - 1, rc0 = (src_floor_s(ox0, g.srw, rW) - 1) & ~3;
    if (t == 0) {
        const uint32_t bar_a = ptx::smem_u32(&bar);
        ptx::mbar_init(bar_a, 1);
        ptx::fence_barrier_init();
        ptx::mbar_expect_tx(bar_a, x_bytes + 3u * g.rr * g.rc * 4u);
        ptx::tma_load_4d(ptx::smem_u32(tile_raw), &tmap_x, bar_a, xc0, xr0, 0, b);
        ptx::tma_load_4d(ptx::smem_u32(tile_raw) + x_bytes_al, &tmap_r, bar_a, rc0, rr0, 0, b);
    }
    const float A = -0.75f;
    if (t < 2 * TILE) {                             // vertical filters of the block's rows, both sources (2 * TILE <= 192 threads)
        const int s = t / TILE, r = t % TILE;
        const int in_size = s ? rH : H;
        const float scale = s ? g.srh : g.sxh;
        const float src = fmaf(scale, (float)min(oy0 + r, oH - 1) + 0.5f, -0.5f);
        const int i0 = min((int)floorf(src), in_size - 1);
        const float tt = fminf(fmaxf(src - (float)i0, 0.f), 1.f), u = 1.f - tt;
        yw[s][r][0] = cubic2(tt + 1.f, A); yw[s][r][1] = cubic1(tt, A); yw[s][r][2] = cubic1(u, A); yw[s][r][3] = cubic2(u + 1.f, A);
    }
    for (int i = t; i < 4 * (BS_W / 2); i += R32_T) {       // horizontal filters: (source, left / right column) x pair
        const int which = i / (BS_W / 2), pair = i % (BS_W / 2);
        const int in_size = (which & 2) ? rW : W;
        const float scale = (which & 2) ? g.srw : g.sxw;
        const float src = fmaf(scale, (float)(ox0 + 2 * pair + (which & 1)) + 0.5f, -0.5f);
        const int i0 = min((int)floorf(src), in_size - 1);
        const float tt = fminf(fmaxf(src - (float)i0, 0.f), 1.f), u = 1.f - tt;
        tapi[which][pair] = i0;
        tapw[which][0][pair] = cubic2(tt + 1.f, A); tapw[which][1][pair] = cubic1(tt, A);
        tapw[which][2][pair] = cubic1(u, A); tapw[which][3][pair] = cubic2(u + 1.f, A);
    }
    __syncthreads();
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    replicate_pad(reinterpret_cast<TI *>(tile_raw), xr0, xc0, g.xr, g.xc, H, W);
    replicate_pad(reinterpret_cast<float *>(tile_raw + x_bytes_al), rr0, rc0, g.rr, g.rc, rH, rW);
    const int ch = t / (BS_W / 2), pair = t % (BS_W / 2), ox = ox0 + 2 * pair;
    if (ox >= oW) return;           // oW is even: a pair is inside or outside as a whole

    // ---- horizontal filters of the pair: six (x) / five (residual) consecutive elements, zero-padded weights
    const uint32_t xpitch = g.xc * sizeof(TI), rpitch = g.rc * 4u;
    ptx::f32x2 wx6[6], wr5[5];
    uint32_t px, pr;                // shared addresses of the pair's first element in tile row 0 of the thread's channel
    {
        float wA[4], wB[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { wA[k] = tapw[0][k][pair]; wB[k] = tapw[1][k][pair]; }
        const int iA = tapi[0][pair], iB = tapi[1][pair];
        const int p0 = (iA - 1) & ~1;                           // even source column (the tile starts on an even column too)
        const int offA = iA - 1 - p0, offB = iB - 1 - p0;         // 0..1, 0..2
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const float a = offA ? tap_or_zero(wA, j - 1) : tap_or_zero(wA, j);
            const float bb = offB == 0 ? tap_or_zero(wB, j) : offB == 1 ? tap_or_zero(wB, j - 1) : tap_or_zero(wB, j - 2);
            wx6[j] = ptx::pk2(a, bb);
        }
        px = ptx::smem_u32(tile_raw) + (uint32_t)(p0 - xc0) * (uint32_t)sizeof(TI) + (uint32_t)ch * g.xr * xpitch;
#pragma unroll
        for (int k = 0; k < 4; ++k) { wA[k] = tapw[2][k][pair]; wB[k] = tapw[3][k][pair]; }
        const int jA = tapi[2][pair], d = tapi[3][pair] - jA;       // d = 0 or 1
#pragma unroll
        for (int j = 0; j < 5; ++j) wr5[j] = ptx::pk2(tap_or_zero(wA, j), d ? tap_or_zero(wB, j - 1) : tap_or_zero(wB, j));
        pr = ptx::smem_u32(tile_raw) + x_bytes_al + (uint32_t)(jA - 1 - rc0) * 4u + (uint32_t)ch * g.rr * rpitch;
    }
    const long oplane = (long)oH * oW;
    // planar: the channel's plane; interleaved: the channel's byte inside the pixel (reversed for BGR), three bytes per pixel
    TO *o = HWC ? out + (((long)b * oH + oy0) * oW + ox) * 3 + (layout == 2 ? 2 - ch : ch) : out + ((long)b * 3 + ch) * oplane + (long)oy0 * oW + ox;
    const int orow = HWC ? 3 * oW : oW;
    const int nrows = min(TILE, oH - oy0);

    // window rows live in fixed registers: tile row n of a source sits in slot n & 3
    ptx::f32x2 wx[4], wr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        wx[i] = hsum6<TI>(px + i * xpitch, wx6);
        wr[i] = hsum5(pr + i * rpitch, wr5);
    }
    px += 4u * xpitch;
    pr += 4u * rpitch;
    const float *ywx = &yw[0][0][0], *ywr = &yw[1][0][0];
#pragma unroll 1
    for (int r0 = 0; r0 < nrows; r0 += P) {
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const int relx = sched_relx<PAT>(q);    // first tile row of the x window (a whole number of rotations per period: slots unchanged)
            const int relr = sched_relr<PAT>(q);    // first tile row of the residual window
            if (RowSched<PAT>::xstep(q)) {
                wx[(relx + 3) & 3] = hsum6<TI>(px, wx6);
                px += xpitch;
            }
            if (RowSched<PAT>::rstep(q)) {
                wr[(relr + 3) & 3] = hsum5(pr, wr5);
                pr += rpitch;
            }
            const float4 a = *reinterpret_cast<const float4 *>(ywx + (r0 + q) * 4);
            const float4 gg = *reinterpret_cast<const float4 *>(ywr + (r0 + q) * 4);
            ptx::f32x2 v = ptx::mul2(wx[relx & 3], ptx::pk2(a.x, a.x));
            v = ptx::fma2(wx[(relx + 1) & 3], ptx::pk2(a.y, a.y), v);
            v = ptx::fma2(wx[(relx + 2) & 3], ptx::pk2(a.z, a.z), v);
            v = ptx::fma2(wx[(relx + 3) & 3], ptx::pk2(a.w, a.w), v);
            ptx::f32x2 u = ptx::mul2(wr[relr & 3], ptx::pk2(gg.x, gg.x));
            u = ptx::fma2(wr[(relr + 1) & 3], ptx::pk2(gg.y, gg.y), u);
            u = ptx::fma2(wr[(relr + 2) & 3], ptx::pk2(gg.z, gg.z), u);
            u = ptx::fma2(wr[(relr + 3) & 3], ptx::pk2(gg.w, gg.w), u);
            v = ptx::add2(v, u);
            float lo, hi;
            ptx::up2(v, lo, hi);
            if (HWC) { o[0] = unit_to_u8<TO>(fminf(fmaxf(lo, 0.f), 1.f)); o[3] = unit_to_u8<TO>(fminf(fmaxf(hi, 0.f), 1.f)); }
            else clamp_store_pair(o, lo, hi);       // outH % P == 0 (host): whole periods of rows
            o += orow;
        }
    }
}

// ---- the same, streaming ---------------------------------------------------------------------------------------------------
// The tile kernel pays its set-up (column filters, vertical table, four window rows per source, the wait for the tile) once per
// 36 output rows: 30 % of its instructions.  Here a CTA walks DOWN a 128-column strip for up to 16 blocks of 12 output rows; the
// window registers simply carry over, and the source rows arrive through a ring of three small stages (8 x rows + 4 residual rows
// = the new rows of one block), loaded by TMA two blocks ahead.  full[] barriers carry the TMA bytes, empty[] barriers count one
// arrival per warp; warp 0 refills the stage released one block earlier.  Load 0 is the window's initial four rows (an 8-row x
// box whose upper half is not used, so that every box has the same shape).  The grid is sized to be co-resident: strips x
// segments <= 5 CTAs per SM.
constexpr int R32S_NST = 3, R32S_MAXBLK = 16;

template <typename TI, typename TO>
__global__ void __launch_bounds__(R32_T, 5) bicubic_add_clamp_r32s_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                                          const __grid_constant__ CUtensorMap tmap_r, int xc, int rc,
                                                                          int H, int W, int rH, int rW, TO *__restrict__ out, int oH,
                                                                          int oW, int blocks_per_seg) {
    pdl_trigger();
    pdl_wait();          // the residual image is written by the previous kernel of the stream
    __shared__ __align__(16) float yw[2][R32S_MAXBLK * 12][4];
    __shared__ float tapw[4][4][BS_W / 2];       // [x left, x right, residual left, residual right][tap][pair]
    __shared__ int tapi[4][BS_W / 2];            // source column of the second tap (floor of the source coordinate)
    __shared__ __align__(8) uint64_t full[R32S_NST], empty[R32S_NST];
    extern __shared__ uint8_t tile_dyn[];
    uint8_t *tile_raw = tile_dyn + ((128u - (ptx::smem_u32(tile_dyn) & 127u)) & 127u);     // TMA destinations are 128-byte aligned
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int ox0 = blockIdx.x * BS_W, b = blockIdx.z;
    const int kb0 = blockIdx.y * blocks_per_seg, nblk = min(blocks_per_seg, oH / 12 - kb0), oy0 = kb0 * 12;
    const int X0 = 8 * kb0 - 6, R0 = 4 * kb0 - 2;   // first source rows of load 0: the window of output row oy0 is x rows 8 kb0 - 2 .. + 1
    constexpr int XA = 16 / (int)sizeof(TI);       // the innermost box coordinate must start on a 16-byte boundary
    const int xc0 = (src_floor(ox0, W, oW) - 1) & ~(XA - 1), rc0 = (src_floor(ox0, rW, oW) - 1) & ~3;
    const uint32_t xpitch = xc * sizeof(TI), rpitch = rc * 4u;
    const uint32_t x_bytes = 3u * 8u * xpitch, r_bytes = 3u * 4u * rpitch;
    const uint32_t xst = (x_bytes + 127u) & ~127u, stage_bytes = xst + ((r_bytes + 127u) & ~127u);
    const uint32_t ring = ptx::smem_u32(tile_raw), full_a = ptx::smem_u32(&full[0]), empty_a = ptx::smem_u32(&empty[0]);
    const int total = nblk + 1;
    auto issue = [&](int L) {
        const uint32_t st = (uint32_t)(L % R32S_NST), bar_a = full_a + 8u * st;
        ptx::mbar_expect_tx(bar_a, x_bytes + r_bytes);
        ptx::tma_load_4d(ring + st * stage_bytes, &tmap_x, bar_a, xc0, X0 + 8 * L, 0, b);
        ptx::tma_load_4d(ring + st * stage_bytes + xst, &tmap_r, bar_a, rc0, R0 + 4 * L, 0, b);
    };
    if (t == 0) {
        for (int i = 0; i < R32S_NST; ++i) { ptx::mbar_init(full_a + 8u * i, 1); ptx::mbar_init(empty_a + 8u * i, R32_T / 32); }
        ptx::fence_barrier_init();
        for (int L = 0; L < R32S_NST && L < total; ++L) issue(L);
    }
    const float A = -0.75f;
    for (int i = t; i < 2 * 12 * nblk; i += R32_T) {      // vertical filters of the segment's rows, both sources
        const int s = i / (12 * nblk), r = i % (12 * nblk);
        const int in_size = s ? rH : H;
        const float scale = (float)in_size / (float)oH;
        const float src = fmaf(scale, (float)(oy0 + r) + 0.5f, -0.5f);
        const int i0 = min((int)floorf(src), in_size - 1);
        const float tt = fminf(fmaxf(src - (float)i0, 0.f), 1.f), u = 1.f - tt;
        yw[s][r][0] = cubic2(tt + 1.f, A); yw[s][r][1] = cubic1(tt, A); yw[s][r][2] = cubic1(u, A); yw[s][r][3] = cubic2(u + 1.f, A);
    }
    for (int i = t; i < 4 * (BS_W / 2); i += R32_T) {       // horizontal filters: (source, left / right column) x pair
        const int which = i / (BS_W / 2), pair = i % (BS_W / 2);
        const int in_size = (which & 2) ? rW : W;
        const float scale = (float)in_size / (float)oW;
        const float src = fmaf(scale, (float)(ox0 + 2 * pair + (which & 1)) + 0.5f, -0.5f);
        const int i0 = min((int)floorf(src), in_size - 1);
        const float tt = fminf(fmaxf(src - (float)i0, 0.f), 1.f), u = 1.f - tt;
        tapi[which][pair] = i0;
        tapw[which][0][pair] = cubic2(tt + 1.f, A); tapw[which][1][pair] = cubic1(tt, A);
        tapw[which][2][pair] = cubic1(u, A); tapw[which][3][pair] = cubic2(u + 1.f, A);
    }
    __syncthreads();
    const int ch = t / (BS_W / 2), pair = t % (BS_W / 2), ox = ox0 + 2 * pair;
    const bool active = ox < oW;    // oW is even: a pair is inside or outside as a whole; idle threads keep the barrier counts

    // ---- horizontal filters of the pair: six (x) / five (residual) consecutive elements, zero-padded weights
    ptx::f32x2 wx6[6], wr5[5];
    uint32_t pxo, pro;              // byte offsets of the pair's first element inside a stage, in tile row 0 of the thread's channel
    {
        float wA[4], wB[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { wA[k] = tapw[0][k][pair]; wB[k] = tapw[1][k][pair]; }
        const int iA = tapi[0][pair], iB = tapi[1][pair];
        const int p0 = (iA - 1) & ~1;                           // even source column (the tile starts on an even column too)
        const int offA = iA - 1 - p0, offB = iB - 1 - p0;         // 0..1, 0..2
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const float a = offA ? tap_or_zero(wA, j - 1) : tap_or_zero(wA, j);
            const float bb = offB == 0 ? tap_or_zero(wB, j) : offB == 1 ? tap_or_zero(wB, j - 1) : tap_or_zero(wB, j - 2);
            wx6[j] = ptx::pk2(a, bb);
        }
        pxo = (active ? (uint32_t)(p0 - xc0) * (uint32_t)sizeof(TI) : 0u) + (uint32_t)ch * 8u * xpitch;
#pragma unroll
        for (int k = 0; k < 4; ++k) { wA[k] = tapw[2][k][pair]; wB[k] = tapw[3][k][pair]; }
        const int jA = tapi[2][pair], d = tapi[3][pair] - jA;       // d = 0 or 1
#pragma unroll
        for (int j = 0; j < 5; ++j) wr5[j] = ptx::pk2(tap_or_zero(wA, j), d ? tap_or_zero(wB, j - 1) : tap_or_zero(wB, j));
        pro = xst + (active ? (uint32_t)(jA - 1 - rc0) * 4u : 0u) + (uint32_t)ch * 4u * rpitch;
    }
    const long oplane = (long)oH * oW;
    TO *o = out + ((long)b * 3 + ch) * oplane + (long)oy0 * oW + (active ? ox : 0);
    // loads whose boxes reach outside the image (uniform over the CTA): every one of a strip at the left / right edge, else the
    // first ones of the top segment and the last ones of the bottom segment
    int kpad_lo = 0, kpad_hi = total;
    if (xc0 < 0 || xc0 + xc > W || rc0 < 0 || rc0 + rc > rW) kpad_lo = total;
    else {
        while (kpad_lo < total && (X0 + 8 * kpad_lo < 0 || R0 + 4 * kpad_lo < 0)) ++kpad_lo;
        while (kpad_hi > kpad_lo && (X0 + 8 * (kpad_hi - 1) + 8 > H || R0 + 4 * (kpad_hi - 1) + 4 > rH)) --kpad_hi;
    }

    // window rows live in fixed registers: row n of a source (counted from the window's first row at oy0) sits in slot n & 3
    ptx::f32x2 wx[4], wr[4];
    const float *ywx = &yw[0][0][0], *ywr = &yw[1][0][0];
#pragma unroll 1
    for (int k = 0; k < total; ++k) {
        const uint32_t st = (uint32_t)(k % R32S_NST);
        if (warp == 0 && k >= 1 && k - 1 + R32S_NST < total) {      // refill the stage every warp released one block ago
            ptx::mbar_wait(empty_a + 8u * (uint32_t)((k - 1) % R32S_NST), (uint32_t)((k - 1) / R32S_NST) & 1u);
            if (lane == 0) issue(k - 1 + R32S_NST);
        }
        ptx::mbar_wait(full_a + 8u * st, (uint32_t)(k / R32S_NST) & 1u);
        const bool pad = k < kpad_lo || k >= kpad_hi;
        if (pad) {          // the TMA zero-fills outside the image; bicubic taps clamp to the border
            replicate_pad(reinterpret_cast<TI *>(tile_raw + st * stage_bytes), X0 + 8 * k, xc0, 8, xc, H, W);
            replicate_pad(reinterpret_cast<float *>(tile_raw + st * stage_bytes + xst), R0 + 4 * k, rc0, 4, rc, rH, rW);
        }
        uint32_t px = ring + st * stage_bytes + pxo, pr = ring + st * stage_bytes + pro;
        if (k == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                wx[i] = hsum6<TI>(px + (4 + i) * xpitch, wx6);
                wr[i] = hsum5(pr + i * rpitch, wr5);
            }
        } else {
            const int r0 = (k - 1) * 12;
#pragma unroll
            for (int q = 0; q < 12; ++q) {
                const int kk = q / 3, j = q % 3;
                const int relx = 2 * kk + j;            // first row of the x window (8 rows per block: slots unchanged)
                const int relr = kk + (j >= 1 ? 1 : 0); // first row of the residual window
                if (j != 0) {                           // x steps one source row on two output rows of three
                    wx[(relx + 3) & 3] = hsum6<TI>(px, wx6);
                    px += xpitch;
                }
                if (j == 1) {                           // the residual steps on one of three
                    wr[(relr + 3) & 3] = hsum5(pr, wr5);
                    pr += rpitch;
                }
                const float4 a = *reinterpret_cast<const float4 *>(ywx + (r0 + q) * 4);
                const float4 gg = *reinterpret_cast<const float4 *>(ywr + (r0 + q) * 4);
                ptx::f32x2 v = ptx::mul2(wx[relx & 3], ptx::pk2(a.x, a.x));
                v = ptx::fma2(wx[(relx + 1) & 3], ptx::pk2(a.y, a.y), v);
                v = ptx::fma2(wx[(relx + 2) & 3], ptx::pk2(a.z, a.z), v);
                v = ptx::fma2(wx[(relx + 3) & 3], ptx::pk2(a.w, a.w), v);
                ptx::f32x2 u = ptx::mul2(wr[relr & 3], ptx::pk2(gg.x, gg.x));
                u = ptx::fma2(wr[(relr + 1) & 3], ptx::pk2(gg.y, gg.y), u);
                u = ptx::fma2(wr[(relr + 2) & 3], ptx::pk2(gg.z, gg.z), u);
                u = ptx::fma2(wr[(relr + 3) & 3], ptx::pk2(gg.w, gg.w), u);
                v = ptx::add2(v, u);
                float lo, hi;
                ptx::up2(v, lo, hi);
                lo = fminf(fmaxf(lo, 0.f), 1.f); hi = fminf(fmaxf(hi, 0.f), 1.f);
                if (active) store_pair_unit(o, lo, hi);
                o += oW;
            }
        }
        __syncwarp();
        if (pad) ptx::fence_proxy_async();      // the stage was rewritten through the generic proxy; the next writer is the TMA
        if (lane == 0) ptx::mbar_arrive(empty_a + 8u * st);
    }
}

// The tile kernel's row schedule, checked with the device's own fp32 coordinate arithmetic (ATen's): the source row of output row oy
// must be the exact-arithmetic floor on every row — schedule 0: floor((4 oy - 1) / 6) for x (3:2) and floor((oy - 1) / 3) for the
// residual (3:1); schedule n: floor((2 oy + 1 - n) / 2n) for x (n:1) and floor((2 oy + 1 - 2n) / 4n) for the residual (2n:1).  Where
// the exact coordinate is an integer (every third row at 3:1) the rounding of scale = (float)in / out decides which side the floor falls.
// Returns the schedule, or -1 (the pair kernel handles everything else).
static int floor_div(int a, int b) { return (a >= 0 ? a : a - b + 1) / b; }
static int row_pattern(int H, int rH, int oH) {
    int pat = -1;
    if ((long)H * 3 == (long)oH * 2 && (long)rH * 3 == (long)oH) pat = 0;
    else if (H > 0 && oH % H == 0 && (long)rH * 2 * (oH / H) == (long)oH) pat = oH / H;
    if (!(pat == 0 || pat == 2 || pat == 3 || pat == 4 || pat == 6) || oH % (pat ? 8 * pat : 12) != 0) return -1;
    static thread_local int key[3] = {0, 0, 0}, val = -1;
    if (key[0] == H && key[1] == rH && key[2] == oH) return val;
    const float sx = (float)H / (float)oH, sr = (float)rH / (float)oH;
    bool ok = true;
    for (int oy = 0; oy < oH && ok; ++oy) {
        int ix = (int)floorf(fmaf(sx, (float)oy + 0.5f, -0.5f));
        int ir = (int)floorf(fmaf(sr, (float)oy + 0.5f, -0.5f));
        if (ix > H - 1) ix = H - 1;
        if (ir > rH - 1) ir = rH - 1;
        const int ex = pat == 0 ? floor_div(4 * oy - 1, 6) : floor_div(2 * oy + 1 - pat, 2 * pat);
        const int er = pat == 0 ? floor_div(oy - 1, 3) : floor_div(2 * oy + 1 - 2 * pat, 4 * pat);
        ok = ix == ex && ir == er;
    }
    key[0] = H; key[1] = rH; key[2] = oH; val = ok ? pat : -1;
    return val;
}

// triangle-filter taps for one output index: [lo, lo+n) and the normalisation 1/sum
__device__ __forceinline__ void aa_range(int dst, int in_size, int out_size, int &lo, int &n, float &center, float &inv,
                                         float &norm) {
    const float scale = (float)in_size / (float)out_size;
    const float support = scale >= 1.f ? scale : 1.f;
    center = scale * ((float)dst + 0.5f);
    inv = scale >= 1.f ? 1.f / scale : 1.f;
    lo = max((int)(center - support + 0.5f), 0);
    n = min((int)(center + support + 0.5f), in_size) - lo;
    float tot = 0.f;
    for (int j = 0; j < n; ++j) tot += fmaxf(0.f, 1.f - fabsf(((float)(j + lo) - center + 0.5f) * inv));
    norm = tot != 0.f ? 1.f / tot : 0.f;
}

template <typename T, typename TO>
__global__ void __launch_bounds__(256) resize_aa_kernel(const T *__restrict__ in, TO *__restrict__ out, int H, int W, int oH,
                                                        int oW, int clamp, int layout) {
    const int ox = blockIdx.x * 64 + (threadIdx.x & 63);
    const int oy = blockIdx.y * 4 + (threadIdx.x >> 6);
    const long plane = blockIdx.z;
    if (ox >= oW || oy >= oH) return;
    int ylo, yn, xlo, xn;
    float yc, yi, ynorm, xc, xi, xnorm;
    aa_range(oy, H, oH, ylo, yn, yc, yi, ynorm);
    aa_range(ox, W, oW, xlo, xn, xc, xi, xnorm);
    const T *p = in + plane * H * W;
    float acc = 0.f;
    for (int a = 0; a < yn; ++a) {
        const float wy = fmaxf(0.f, 1.f - fabsf(((float)(a + ylo) - yc + 0.5f) * yi)) * ynorm;
        float t = 0.f;
        for (int c = 0; c < xn; ++c) {
            const float wx = fmaxf(0.f, 1.f - fabsf(((float)(c + xlo) - xc + 0.5f) * xi)) * xnorm;
            t = fmaf(wx, to_f(p[(long)(a + ylo) * W + c + xlo]), t);
        }
        acc = fmaf(wy, t, acc);
    }
    if (clamp) acc = fminf(fmaxf(acc, 0.f), 1.f);
    if (layout == 0) {
        out[(plane * oH + oy) * oW + ox] = from_f<TO>(acc);
    } else {
        const long b = plane / 3;
        const int c = (int)(plane - 3 * b);
        out[((b * oH + oy) * oW + ox) * 3 + (layout == 2 ? 2 - c : c)] = from_f<TO>(acc);
    }
}

// interleaved uint8 frames (B,H,W,3) -> planar RGB (B,3,H,W): four pixels per thread, 12 bytes in, 3 x 4 bytes out
__global__ void __launch_bounds__(256) frames_to_planar_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, long plane,
                                                               int swap) {
    pdl_trigger();
    pdl_wait();
    const long b = blockIdx.y;
    const long q = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (q >= plane) return;
    const uint8_t *src = in + (b * plane + q) * 3;
    uint8_t *dst = out + b * 3 * plane + q;
    const int c0 = swap ? 2 : 0, c2 = 2 - c0;
    if (q + 4 <= plane && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)plane) & 3) == 0) {
        const uint32_t w0 = reinterpret_cast<const uint32_t *>(src)[0], w1 = reinterpret_cast<const uint32_t *>(src)[1],
                       w2 = reinterpret_cast<const uint32_t *>(src)[2];
        // bytes: w0 = p0c0 p0c1 p0c2 p1c0 | w1 = p1c1 p1c2 p2c0 p2c1 | w2 = p2c2 p3c0 p3c1 p3c2
        const uint32_t ch0 = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);      // p0c0 p1c0 p2c0 p3c0
        const uint32_t ch1 = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);      // p0c1 p1c1 p2c1 p3c1
        const uint32_t ch2 = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);      // p0c2 p1c2 p2c2 p3c2
        reinterpret_cast<uint32_t *>(dst + c0 * plane)[0] = ch0;
        reinterpret_cast<uint32_t *>(dst + plane)[0] = ch1;
        reinterpret_cast<uint32_t *>(dst + c2 * plane)[0] = ch2;
    } else {
        for (int i = 0; i < 4 && q + i < plane; ++i) {
            dst[c0 * plane + i] = src[3 * i];
            dst[plane + i] = src[3 * i + 1];
            dst[c2 * plane + i] = src[3 * i + 2];
        }
    }
}

}  // namespace tu

using namespace tu;

thread_local int tu::g_bicubic_tile = 0;      // debug key "bicubic_tile": rows per CTA of the row-schedule kernel: 0 = by grid size (default), 1 = the smaller tile, 2 = the larger
thread_local int tu::g_bicubic_pair = 2;      // debug key "bicubic_pair": 2 = two output columns per thread + the unrolled fixed-row-pattern kernel where it applies (default), 3 = the same with the streaming form of that kernel (measured slower), 1 = the pair kernel only, 0 = the one-column strip kernel

// (W, H, 3, B) view of an NCHW image for the tile loads; box = (cols, rows, 3, 1)
static bool encode_image_map(CUtensorMap *tm, const void *ptr, int elem_bytes, int B, int H, int W, int box_rows, int box_cols) {
    TcEncodeFn enc = tc_encode_fn();
    if (!enc || box_rows > 256 || box_cols > 256) return false;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)W * elem_bytes, (cuuint64_t)H * W * elem_bytes, (cuuint64_t)3 * H * W * elem_bytes};
    cuuint32_t box[4] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 3, 1}, es[4] = {1, 1, 1, 1};
    return enc(tm, elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)ptr, dims, strides,
               box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

extern "C" int tu_bicubic_add_clamp(const void *x, int in_dtype, int H, int W, const float *res, int rH, int rW, void *out,
                                    int out_dtype, int B, int outH, int outW, int clamp, void *stream) {
    TU_CHECK_ARG(x && out && B > 0 && H > 0 && W > 0 && outH > 0 && outW > 0, "bicubic_add_clamp: bad argument");
    const int layout = dtype_layout(out_dtype);
    out_dtype = dtype_base(out_dtype);
    TU_CHECK_ARG((in_dtype == TU_F32 || in_dtype == TU_BF16 || in_dtype == TU_U8) && (out_dtype == TU_F32 || out_dtype == TU_BF16 || out_dtype == TU_U8),
                 "bicubic_add_clamp: bad dtype");
    TU_CHECK_ARG(layout == 0 || (out_dtype == TU_U8 && layout <= 2), "bicubic_add_clamp: interleaved layouts are for uint8 frames");
    cudaStream_t st = (cudaStream_t)stream;
    const int eb = (int)dtype_size(in_dtype), ob = (int)dtype_size(out_dtype);
    BicubicTileGeom g;
    g.sxh = (float)H / (float)outH; g.sxw = (float)W / (float)outW;
    g.srh = res ? (float)rH / (float)outH : 0.f; g.srw = res ? (float)rW / (float)outW : 0.f;
    size_t tile_bytes = 0;
    CUtensorMap tx, tr;
    // source footprint of a bh x 128 output tile (+4 taps, +2 for fp32 rounding of the coordinates), cols padded to 16 bytes
    auto plan = [&](int bh) -> bool {
        g.xr = (int)((long)(bh - 1) * H / outH) + 6;
        g.xc = ((int)((long)(BS_W - 1) * W / outW) + 8 + (16 / eb - 1) + 15) & ~15;        // + alignment slack of the first column
        g.rr = res ? (int)((long)(bh - 1) * rH / outH) + 6 : 0;
        g.rc = res ? ((int)((long)(BS_W - 1) * rW / outW) + 8 + 3 + 3) & ~3 : 0;
        const size_t x_bytes = ((size_t)3 * g.xr * g.xc * eb + 127) & ~(size_t)127;
        tile_bytes = x_bytes + (size_t)3 * g.rr * g.rc * 4;
        memset(&tx, 0, sizeof(tx));
        memset(&tr, 0, sizeof(tr));
        bool ok = tile_bytes <= 96 * 1024 && ((size_t)W * eb) % 16 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                  (!res || (((size_t)rW * 4) % 16 == 0 && (reinterpret_cast<uintptr_t>(res) & 15) == 0));
        return ok && encode_image_map(&tx, x, eb, B, H, W, g.xr, g.xc) && (!res || encode_image_map(&tr, res, 4, B, rH, rW, g.rr, g.rc));
    };
    // two output columns per thread: TMA tiles, even output width, pair stores aligned
    const bool pair = g_bicubic_pair && (outW % 2) == 0 && (reinterpret_cast<uintptr_t>(out) % (2 * ob)) == 0 && plan(PAIR_H);
    // periodic row schedules (x1.5 outputs such as 720p -> 1080p, x3 outputs such as 720p -> 4K): the unrolled circular-window kernel
    const int pat = pair && g_bicubic_pair >= 2 && res && clamp && (in_dtype == TU_BF16 || in_dtype == TU_U8) &&
                            W <= outW && rW <= outW
                        ? row_pattern(H, rH, outH)
                        : -1;
    // rows per CTA: the larger tile as soon as it still gives every SM a CTA (measured at 720p -> 1080p, profiles/r3_ab_bicubic_tile_rows.log:
    // 1 / 2 / 4 / 8 frames = 16.3 -> 14.8, 23.1 -> 21.8, 34.1 -> 34.7, 63.3 -> 59.2 us), else the smaller one
    int tile_rows = 0;
    bool r32 = false;
    if (pat >= 0) {
        const long ctas_big = (long)ceil_div(outW, BS_W) * ceil_div(outH, sched_tile(pat, true)) * B;
        tile_rows = sched_tile(pat, g_bicubic_tile ? g_bicubic_tile == 2 : ctas_big >= (long)device_sm_count());
        if (!(r32 = plan(tile_rows))) plan(PAIR_H);            // or back to the pair kernel's plan
    }
    static_assert(RowSched<0>::BIG <= R32_MAXTILE && RowSched<2>::BIG <= R32_MAXTILE && RowSched<3>::BIG <= R32_MAXTILE &&
                      RowSched<4>::BIG <= R32_MAXTILE && RowSched<6>::BIG <= R32_MAXTILE && 2 * R32_MAXTILE <= R32_T,
                  "the vertical table is filled in one pass of the CTA's threads");
    const bool tma = pair || plan(BS_H);
    dim3 grid(ceil_div(outW, BS_W), ceil_div(outH, pair ? PAIR_H : BS_H), B);
#define TU_BIC_PAIR(TI, TO)                                                                                                     \
    do {                                                                                                                        \
        static PerDeviceFlag attr_done;                                                                                          \
        if (!attr_done.is_set()) {                                                                                                       \
            cudaError_t e = cudaFuncSetAttribute(bicubic_add_clamp_pair_kernel<TI, TO, true>,                                  \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);                      \
            if (e == cudaSuccess)                                                                                               \
                e = cudaFuncSetAttribute(bicubic_add_clamp_pair_kernel<TI, TO, false>,                                         \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);                              \
            if (e != cudaSuccess) return cuda_fail(e, "bicubic smem attribute");                                                \
            attr_done.set();                                                                                                   \
        }                                                                                                                       \
        if (res)                                                                                                                \
            launch_pdl(bicubic_add_clamp_pair_kernel<TI, TO, true>, grid, dim3(BS_W / 2), tile_bytes + 128, st, tx, tr, g, H, W, rH, rW, \
                       (TO *)out, outH, outW, clamp, layout);                                                                  \
        else                                                                                                                    \
            launch_pdl(bicubic_add_clamp_pair_kernel<TI, TO, false>, grid, dim3(BS_W / 2), tile_bytes + 128, st, tx, tr, g, H, W, rH, rW, \
                       (TO *)out, outH, outW, clamp, layout);                                                                  \
    } while (0)
#define TU_BIC(TI, TO)                                                                                                          \
    do {                                                                                                                        \
        if (tma) {                                                                                                              \
            static PerDeviceFlag attr_done;                                                                                      \
            if (!attr_done.is_set()) {                                                                                                   \
                cudaError_t e = cudaFuncSetAttribute(bicubic_add_clamp_strip_kernel<TI, TO, 1>,                              \
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);                  \
                if (e == cudaSuccess)                                                                                           \
                    e = cudaFuncSetAttribute(bicubic_add_clamp_strip_kernel<TI, TO, 2>,                                         \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);                          \
                if (e != cudaSuccess) return cuda_fail(e, "bicubic smem attribute");                                            \
                attr_done.set();                                                                                               \
            }                                                                                                                   \
            if (res)                                                                                                            \
                launch_pdl(bicubic_add_clamp_strip_kernel<TI, TO, 1>, grid, dim3(BS_W), tile_bytes + 128, st, tx, tr, g, (const TI *)x, H, W, \
                           res, rH, rW, (TO *)out, outH, outW, clamp, layout);                                              \
            else                                                                                                                \
                launch_pdl(bicubic_add_clamp_strip_kernel<TI, TO, 2>, grid, dim3(BS_W), tile_bytes + 128, st, tx, tr, g, (const TI *)x, H, W, \
                           res, rH, rW, (TO *)out, outH, outW, clamp, layout);                                              \
        } else {                                                                                                                \
            bicubic_add_clamp_strip_kernel<TI, TO, 0><<<grid, BS_W, 0, st>>>(tx, tr, g, (const TI *)x, H, W, res, rH, rW,     \
                                                                                  (TO *)out, outH, outW, clamp, layout);        \
        }                                                                                                                       \
    } while (0)
#define TU_BIC2(TI, TO)          \
    do {                         \
        if (pair) TU_BIC_PAIR(TI, TO); \
        else TU_BIC(TI, TO);     \
    } while (0)
#define TU_BIC_R32_P(TI, TO, PAT, HWC)                                                                                          \
    do {                                                                                                                        \
        static PerDeviceFlag attr_done;                                                                                          \
        if (!attr_done.is_set()) {                                                                                               \
            cudaError_t e = cudaFuncSetAttribute(bicubic_add_clamp_r32_kernel<TI, TO, PAT, HWC>,                                \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);                       \
            if (e != cudaSuccess) return cuda_fail(e, "bicubic smem attribute");                                                \
            attr_done.set();                                                                                                    \
        }                                                                                                                       \
        launch_pdl(bicubic_add_clamp_r32_kernel<TI, TO, PAT, HWC>, dim3(grid.x, ceil_div(outH, tile_rows), B), dim3(R32_T),     \
                   tile_bytes + 128, st, tx, tr, g, H, W, rH, rW, (TO *)out, outH, outW, layout, tile_rows);                    \
    } while (0)
#define TU_BIC_R32_L(TI, TO, HWC)                                                                                               \
    do {                                                                                                                        \
        if (pat == 0) TU_BIC_R32_P(TI, TO, 0, HWC);                                                                             \
        else if (pat == 2) TU_BIC_R32_P(TI, TO, 2, HWC);                                                                        \
        else if (pat == 3) TU_BIC_R32_P(TI, TO, 3, HWC);                                                                        \
        else if (pat == 4) TU_BIC_R32_P(TI, TO, 4, HWC);                                                                        \
        else TU_BIC_R32_P(TI, TO, 6, HWC);                                                                                      \
    } while (0)
#define TU_BIC_R32(TI, TO) TU_BIC_R32_L(TI, TO, false)
    if (r32 && pat == 0 && layout == 0 && g_bicubic_pair >= 3) {
        // streaming variant: strips x segments co-resident (5 CTAs per SM), <= 16 blocks of 12 rows per segment
        const int nblk = outH / 12, strips = (int)grid.x * B;
        const size_t stage = (((size_t)3 * 8 * g.xc * eb + 127) & ~(size_t)127) + (((size_t)3 * 4 * g.rc * 4 + 127) & ~(size_t)127);
        const size_t ring_bytes = R32S_NST * stage + 128;
        // CTAs per SM: 5 by registers (64 x 192 threads), fewer if the ring + 12 KB of static tables + 1 KB reserved do not fit 227 KB
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(5, (size_t)227 * 1024 / (ring_bytes + 13 * 1024)));
        int nseg = std::max(1, std::min(nblk, device_sm_count() * per_sm / std::max(strips, 1)));
        int bps = std::min(std::max(ceil_div(nblk, nseg), std::min(3, nblk)), R32S_MAXBLK);      // at least 36 rows per CTA
        nseg = ceil_div(nblk, bps);
        CUtensorMap sx, sr;
        memset(&sx, 0, sizeof(sx));
        memset(&sr, 0, sizeof(sr));
        if (ring_bytes <= 96 * 1024 && encode_image_map(&sx, x, eb, B, H, W, 8, g.xc) && encode_image_map(&sr, res, 4, B, rH, rW, 4, g.rc)) {
            dim3 sgrid(grid.x, nseg, B);
#define TU_BIC_R32S(TI, TO)                                                                                                     \
    do {                                                                                                                        \
        static PerDeviceFlag attr_done;                                                                                          \
        if (!attr_done.is_set()) {                                                                                               \
            cudaError_t e = cudaFuncSetAttribute(bicubic_add_clamp_r32s_kernel<TI, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                 96 * 1024);                                                                    \
            if (e != cudaSuccess) return cuda_fail(e, "bicubic smem attribute");                                                \
            attr_done.set();                                                                                                    \
        }                                                                                                                       \
        launch_pdl(bicubic_add_clamp_r32s_kernel<TI, TO>, sgrid, dim3(R32_T), ring_bytes, st, sx, sr, g.xc, g.rc, H, W, rH, rW,   \
                   (TO *)out, outH, outW, bps);                                                                                 \
    } while (0)
            if (in_dtype == TU_BF16 && out_dtype == TU_BF16) TU_BIC_R32S(bf16, bf16);
            else if (in_dtype == TU_BF16 && out_dtype == TU_F32) TU_BIC_R32S(bf16, float);
            else if (in_dtype == TU_BF16) TU_BIC_R32S(bf16, uint8_t);
            else if (out_dtype == TU_U8) TU_BIC_R32S(uint8_t, uint8_t);
            else if (out_dtype == TU_BF16) TU_BIC_R32S(uint8_t, bf16);
            else TU_BIC_R32S(uint8_t, float);
#undef TU_BIC_R32S
            TU_CHECK_LAUNCH("bicubic_add_clamp");
            return TU_OK;
        }
    }
    if (r32 && layout != 0) {           // interleaved uint8 frames (layouts are accepted for uint8 output only)
        if (in_dtype == TU_BF16) TU_BIC_R32_L(bf16, uint8_t, true);
        else TU_BIC_R32_L(uint8_t, uint8_t, true);
    } else if (r32) {
        if (in_dtype == TU_BF16 && out_dtype == TU_BF16) TU_BIC_R32(bf16, bf16);
        else if (in_dtype == TU_BF16 && out_dtype == TU_F32) TU_BIC_R32(bf16, float);
        else if (in_dtype == TU_BF16) TU_BIC_R32(bf16, uint8_t);
        else if (out_dtype == TU_U8) TU_BIC_R32(uint8_t, uint8_t);
        else if (out_dtype == TU_BF16) TU_BIC_R32(uint8_t, bf16);
        else TU_BIC_R32(uint8_t, float);
    } else
    if (in_dtype == TU_F32 && out_dtype == TU_F32) TU_BIC2(float, float);
    else if (in_dtype == TU_F32 && out_dtype == TU_BF16) TU_BIC2(float, bf16);
    else if (in_dtype == TU_BF16 && out_dtype == TU_F32) TU_BIC2(bf16, float);
    else if (in_dtype == TU_BF16 && out_dtype == TU_BF16) TU_BIC2(bf16, bf16);
    else if (in_dtype == TU_U8 && out_dtype == TU_U8) TU_BIC2(uint8_t, uint8_t);
    else if (in_dtype == TU_U8 && out_dtype == TU_F32) TU_BIC2(uint8_t, float);
    else if (in_dtype == TU_U8 && out_dtype == TU_BF16) TU_BIC2(uint8_t, bf16);
    else if (in_dtype == TU_F32 && out_dtype == TU_U8) TU_BIC2(float, uint8_t);
    else TU_BIC2(bf16, uint8_t);
#undef TU_BIC2
#undef TU_BIC_R32
#undef TU_BIC_R32_P
#undef TU_BIC_R32_L
#undef TU_BIC_PAIR
#undef TU_BIC
    TU_CHECK_LAUNCH("bicubic_add_clamp");
    return TU_OK;
}

extern "C" int tu_bicubic_row_schedule(int H, int rH, int outH) { return H > 0 && rH > 0 && outH > 0 ? row_pattern(H, rH, outH) : -1; }

extern "C" int tu_resize_bilinear_aa_to(const void *in, int in_dtype, void *out, int out_dtype, int B, int H, int W, int outH, int outW,
                                        int clamp, void *stream) {
    TU_CHECK_ARG(in && out && B > 0 && H > 0 && W > 0 && outH > 0 && outW > 0, "resize_bilinear_aa: bad argument");
    const int layout = dtype_layout(out_dtype);
    out_dtype = dtype_base(out_dtype);
    TU_CHECK_ARG(layout == 0 || (out_dtype == TU_U8 && layout <= 2), "resize_bilinear_aa: interleaved layouts are for uint8 frames");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(ceil_div(outW, 64), ceil_div(outH, 4), B * 3);
#define TU_RS(TI, TO) resize_aa_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI *)in, (TO *)out, H, W, outH, outW, clamp, layout)
    if (in_dtype == TU_F32 && out_dtype == TU_F32) TU_RS(float, float);
    else if (in_dtype == TU_BF16 && out_dtype == TU_BF16) TU_RS(bf16, bf16);
    else if (in_dtype == TU_F32 && out_dtype == TU_BF16) TU_RS(float, bf16);
    else if (in_dtype == TU_BF16 && out_dtype == TU_F32) TU_RS(bf16, float);
    else if (in_dtype == TU_F32 && out_dtype == TU_U8) TU_RS(float, uint8_t);
    else if (in_dtype == TU_BF16 && out_dtype == TU_U8) TU_RS(bf16, uint8_t);
    else TU_CHECK_ARG(false, "resize_bilinear_aa: bad dtype");
#undef TU_RS
    TU_CHECK_LAUNCH("resize_bilinear_aa");
    return TU_OK;
}

extern "C" int tu_resize_bilinear_aa(const void *in, int dtype, void *out, int B, int H, int W, int outH, int outW,
                                     int clamp, void *stream) {
    TU_CHECK_ARG(dtype == TU_F32 || dtype == TU_BF16, "resize_bilinear_aa: bad dtype");
    return tu_resize_bilinear_aa_to(in, dtype, out, dtype, B, H, W, outH, outW, clamp, stream);
}

extern "C" int tu_frames_to_planar(const void *in, int layout, void *out, int B, int H, int W, void *stream) {
    TU_CHECK_ARG(in && out && B > 0 && H > 0 && W > 0, "frames_to_planar: bad argument");
    TU_CHECK_ARG(layout == TU_LAYOUT_HWC || layout == TU_LAYOUT_HWC_BGR, "frames_to_planar: layout must be TU_LAYOUT_HWC or TU_LAYOUT_HWC_BGR");
    const long plane = (long)H * W;
    dim3 grid((unsigned)((plane + 1023) / 1024), B);
    launch_pdl(frames_to_planar_kernel, grid, dim3(256), 0, (cudaStream_t)stream, (const uint8_t *)in, (uint8_t *)out, plane,
               layout == TU_LAYOUT_HWC_BGR ? 1 : 0);
    TU_CHECK_LAUNCH("frames_to_planar");
    return TU_OK;
}
