// Memory-bound resampling kernels.
//  * fused  out = clamp( bicubic(x) + bicubic(residual) )  — replaces two F.interpolate(mode='bicubic',
//    align_corners=False) calls, the add and the clamp (WindowTransformer/model.py:241,301,304-305;
//    ResidualTransformer/model.py:125,160,163-164) with one pass: x and residual read once, out written once.
//    Tap arithmetic follows ATen bit-for-bit in fp32 (ATen/native/UpSample.h:259-312,400-448):
//    scale = (float)in/out, src = fma(scale, dst+0.5, -0.5), i = floor(src), t = src - i,
//    Keys cubic A = -0.75, taps clamped to [0, in-1].
//  * antialiased bilinear resize (torchvision Resize on a tensor = ATen _upsample_bilinear2d_aa), used by
//    FastTransformer when the integer factor overshoots res_out (FastTransformer/model.py:323-325).
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

template <typename TS>
__device__ __forceinline__ float cubic_sample(const TS *__restrict__ plane, int W, const Cubic &cy, const Cubic &cx) {
    float out = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const TS *row = plane + (long)cy.idx[i] * W;
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) t += to_f(row[cx.idx[j]]) * cx.w[j];
        out += t * cy.w[i];
    }
    return out;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) bicubic_add_clamp_kernel(const TI *__restrict__ x, int H, int W,
                                                                const float *__restrict__ res, int rH, int rW,
                                                                TO *__restrict__ out, int oH, int oW, int clamp) {
    const int ox = blockIdx.x * 64 + (threadIdx.x & 63);
    const int oy = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int b = blockIdx.z;
    if (ox >= oW || oy >= oH) return;
    const Cubic cy = cubic_taps(oy, H, oH), cx = cubic_taps(ox, W, oW);
    Cubic ry, rx;
    if (res) { ry = cubic_taps(oy, rH, oH); rx = cubic_taps(ox, rW, oW); }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = cubic_sample<TI>(x + ((long)b * 3 + c) * H * W, W, cy, cx);
        if (res) v += cubic_sample<float>(res + ((long)b * 3 + c) * rH * rW, rW, ry, rx);
        if (clamp) v = fminf(fmaxf(v, 0.f), 1.f);
        out[(((long)b * 3 + c) * oH + oy) * oW + ox] = from_f<TO>(v);
    }
}

// Tiled separable version for up-scaling (at most TR source rows feed the 16 output rows of a tile):
// per tile, every needed source row is first resampled horizontally into shared memory (4 taps), then the 16
// output rows take their 4 vertical taps from shared memory.  Same operation order as ATen (horizontal sum
// inside each source row, then the vertical sum), ~6x fewer instructions per output than the direct kernel.
constexpr int BT_W = 128, BT_H = 16, BT_R = 24;

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) bicubic_add_clamp_tiled_kernel(const TI *__restrict__ x, int H, int W,
                                                                      const float *__restrict__ res, int rH, int rW,
                                                                      TO *__restrict__ out, int oH, int oW, int clamp) {
    __shared__ float hx[BT_R][BT_W];
    __shared__ float hr[BT_R][BT_W];
    __shared__ int yidx[2][BT_H][4];
    __shared__ float yw[2][BT_H][4];
    const int t = threadIdx.x, cx = t & (BT_W - 1), half = t >> 7;
    const int ox0 = blockIdx.x * BT_W, oy0 = blockIdx.y * BT_H, b = blockIdx.z;
    const int ox = ox0 + cx;
    const Cubic tx = cubic_taps(min(ox, oW - 1), W, oW);
    Cubic tr = tx;
    if (res) tr = cubic_taps(min(ox, oW - 1), rW, oW);
    if (t < 2 * BT_H) {
        const int s = t / BT_H, r = t % BT_H;
        const Cubic c = cubic_taps(min(oy0 + r, oH - 1), s ? rH : H, oH);
#pragma unroll
        for (int i = 0; i < 4; ++i) { yidx[s][r][i] = c.idx[i]; yw[s][r][i] = c.w[i]; }
    }
    __syncthreads();
    const int lo_x = yidx[0][0][0], n_x = yidx[0][BT_H - 1][3] - lo_x + 1;
    const int lo_r = yidx[1][0][0], n_r = res ? yidx[1][BT_H - 1][3] - lo_r + 1 : 0;
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
        const TI *px = x + (((long)b * 3 + c) * H + lo_x) * W;
        for (int s = half; s < n_x; s += 2) {
            const TI *row = px + (long)s * W;
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) v += to_f(row[tx.idx[j]]) * tx.w[j];
            hx[s][cx] = v;
        }
        if (res) {
            const float *pr = res + (((long)b * 3 + c) * rH + lo_r) * rW;
            for (int s = half; s < n_r; s += 2) {
                const float *row = pr + (long)s * rW;
                float v = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) v += row[tr.idx[j]] * tr.w[j];
                hr[s][cx] = v;
            }
        }
        __syncthreads();
        if (ox < oW) {
#pragma unroll
            for (int k = 0; k < BT_H / 2; ++k) {
                const int r = half * (BT_H / 2) + k, oy = oy0 + r;
                if (oy >= oH) break;
                float v = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) v += hx[yidx[0][r][i] - lo_x][cx] * yw[0][r][i];
                if (res) {
                    float u = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) u += hr[yidx[1][r][i] - lo_r][cx] * yw[1][r][i];
                    v += u;
                }
                if (clamp) v = fminf(fmaxf(v, 0.f), 1.f);
                out[(((long)b * 3 + c) * oH + oy) * oW + ox] = from_f<TO>(v);
            }
        }
        __syncthreads();
    }
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

template <typename T>
__global__ void __launch_bounds__(256) resize_aa_kernel(const T *__restrict__ in, T *__restrict__ out, int H, int W, int oH,
                                                        int oW, int clamp) {
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
    out[(plane * oH + oy) * oW + ox] = from_f<T>(acc);
}

}  // namespace tu

using namespace tu;

extern "C" int tu_bicubic_add_clamp(const void *x, int in_dtype, int H, int W, const float *res, int rH, int rW, void *out,
                                    int out_dtype, int B, int outH, int outW, int clamp, void *stream) {
    TU_CHECK_ARG(x && out && B > 0 && H > 0 && W > 0 && outH > 0 && outW > 0, "bicubic_add_clamp: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(ceil_div(outW, 64), ceil_div(outH, 4), B);
    // tiled kernel when the 16 output rows of a tile never need more than BT_R source rows (any up-scaling)
    const bool tiled = (long)BT_H * H <= (long)(BT_R - 5) * outH && (!res || (long)BT_H * rH <= (long)(BT_R - 5) * outH);
    dim3 tgrid(ceil_div(outW, BT_W), ceil_div(outH, BT_H), B);
#define TU_BIC(TI, TO)                                                                                                          \
    if (tiled)                                                                                                                  \
        bicubic_add_clamp_tiled_kernel<TI, TO><<<tgrid, 256, 0, st>>>((const TI *)x, H, W, res, rH, rW, (TO *)out, outH, outW, clamp); \
    else                                                                                                                        \
        bicubic_add_clamp_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI *)x, H, W, res, rH, rW, (TO *)out, outH, outW, clamp)
    if (in_dtype == TU_F32 && out_dtype == TU_F32) { TU_BIC(float, float); }
    else if (in_dtype == TU_F32 && out_dtype == TU_BF16) { TU_BIC(float, bf16); }
    else if (in_dtype == TU_BF16 && out_dtype == TU_F32) { TU_BIC(bf16, float); }
    else if (in_dtype == TU_BF16 && out_dtype == TU_BF16) { TU_BIC(bf16, bf16); }
    else TU_CHECK_ARG(false, "bicubic_add_clamp: bad dtype");
#undef TU_BIC
    TU_CHECK_LAUNCH("bicubic_add_clamp");
    return TU_OK;
}

extern "C" int tu_resize_bilinear_aa(const void *in, int dtype, void *out, int B, int H, int W, int outH, int outW,
                                     int clamp, void *stream) {
    TU_CHECK_ARG(in && out && B > 0 && H > 0 && W > 0 && outH > 0 && outW > 0, "resize_bilinear_aa: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(ceil_div(outW, 64), ceil_div(outH, 4), B * 3);
    if (dtype == TU_F32)
        resize_aa_kernel<float><<<grid, 256, 0, st>>>((const float *)in, (float *)out, H, W, outH, outW, clamp);
    else if (dtype == TU_BF16)
        resize_aa_kernel<bf16><<<grid, 256, 0, st>>>((const bf16 *)in, (bf16 *)out, H, W, outH, outW, clamp);
    else
        TU_CHECK_ARG(false, "resize_bilinear_aa: bad dtype");
    TU_CHECK_LAUNCH("resize_bilinear_aa");
    return TU_OK;
}
