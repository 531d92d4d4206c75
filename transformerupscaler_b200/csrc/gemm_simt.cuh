// CUDA-core (FFMA, fp32 accumulate) tiled GEMM with functor-defined A gather and epilogue.
//   C[m][n] = sum_k A(m,k) * W(n,k)
// This is the exact-fp32 path of the engine (the 1e-4 parity bar needs true FFMA; single-pass TF32 is
// not enough) and the bring-up path for bf16 storage.  Every dense contraction of the
// three models is expressed through it by choosing the A functor (3x3 conv taps, 8x8 patch gather,
// plain rows) and the epilogue functor (bias / ReLU / GELU / residual add / PixelShuffle store /
// window-ordered token store / patch scatter + skip add).
#pragma once
#include "tu_common.cuh"

namespace tu {

// Weight addressing: element (n,k) lives at p + (n/64)*chunk_stride + (k/KC)*tap_stride + (n%64)*KC + (k%KC).
// Plain row-major (N,K): KC = K, chunk_stride = 64*K.  Conv [chunk][tap][co][ci]: KC = 64,
// tap_stride = 64*64, chunk_stride = 9*64*64.
template <typename T> struct WDesc {
    const T *p;
    int KC;
    long tap_stride;
    long chunk_stride;
};

constexpr int GEMM_BM = 128, GEMM_BN = 64, GEMM_BK = 16, GEMM_THREADS = 256;

template <typename T, typename ALoad, typename Epi>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_simt_kernel(ALoad a, WDesc<T> w, int M, int N, int K, Epi epi) {
    __shared__ __align__(16) float As[GEMM_BK][GEMM_BM + 4];
    __shared__ __align__(16) float Ws[GEMM_BK][GEMM_BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * GEMM_BM;
    const int n0 = blockIdx.y * GEMM_BN;
    const int tx = tid & 15, ty = tid >> 4;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // loader mapping: 4 threads cover the 16 k of one row (one float4 each)
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const T *wchunk = w.p + (long)(n0 / 64) * w.chunk_stride;

    for (int k0 = 0; k0 < K; k0 += GEMM_BK) {
        float4 a0 = a.load4(m0 + lrow, k0 + lk);
        float4 a1 = a.load4(m0 + lrow + 64, k0 + lk);
        const T *wt = wchunk + (long)(k0 / w.KC) * w.tap_stride + (k0 % w.KC);
        float4 w0 = load4(wt + (long)lrow * w.KC + lk);
        __syncthreads();   // previous tile fully consumed
        As[lk + 0][lrow] = a0.x; As[lk + 1][lrow] = a0.y; As[lk + 2][lrow] = a0.z; As[lk + 3][lrow] = a0.w;
        As[lk + 0][lrow + 64] = a1.x; As[lk + 1][lrow + 64] = a1.y; As[lk + 2][lrow + 64] = a1.z; As[lk + 3][lrow + 64] = a1.w;
        Ws[lk + 0][lrow] = w0.x; Ws[lk + 1][lrow] = w0.y; Ws[lk + 2][lrow] = w0.z; Ws[lk + 3][lrow] = w0.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GEMM_BK; ++kk) {
            float4 av0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8]);
            float4 av1 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8 + 4]);
            float4 wv = *reinterpret_cast<const float4 *>(&Ws[kk][tx * 4]);
            const float av[8] = {av0.x, av0.y, av0.z, av0.w, av1.x, av1.y, av1.z, av1.w};
            const float wn[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wn[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m0 + ty * 8 + i;
        if (m < M) epi(m, n0 + tx * 4, acc[i]);
    }
}

template <typename T, typename ALoad, typename Epi>
static inline int launch_gemm_simt(ALoad a, WDesc<T> w, int M, int N, int K, Epi epi, cudaStream_t st, const char *what) {
    if (M <= 0) return TU_OK;
    if (N % GEMM_BN != 0 || K % GEMM_BK != 0 || w.KC % GEMM_BK != 0) {
        set_error(std::string("tu: gemm shape not tileable in ") + what);
        return TU_ERR_ARG;
    }
    dim3 grid(ceil_div(M, GEMM_BM), N / GEMM_BN);
    gemm_simt_kernel<T, ALoad, Epi><<<grid, GEMM_THREADS, 0, st>>>(a, w, M, N, K, epi);
    TU_CHECK_LAUNCH(what);
    return TU_OK;
}

// ------------------------------------------------------------------ A functors
// plain row-major rows of T (or float) with leading dimension ld
template <typename TA> struct ARows {
    const TA *p;
    int M;
    long ld;
    __device__ __forceinline__ float4 load4(int m, int k) const {
        if (m >= M) return make_float4(0.f, 0.f, 0.f, 0.f);
        return tu::load4(p + (long)m * ld + k);
    }
};

// 3x3 / pad 1 / stride S taps over NHWC with 64 channels: m = output pixel, k = tap*64 + ci
template <typename T> struct AConv3x3 {
    const T *p;
    int M, H, W, Ho, Wo, stride;
    __device__ __forceinline__ float4 load4(int m, int k) const {
        if (m >= M) return make_float4(0.f, 0.f, 0.f, 0.f);
        int tap = k >> 6, ci = k & 63;
        int x = m % Wo, t = m / Wo;
        int y = t % Ho, b = t / Ho;
        int iy = y * stride + tap / 3 - 1, ix = x * stride + tap % 3 - 1;
        if (iy < 0 || iy >= H || ix < 0 || ix >= W) return make_float4(0.f, 0.f, 0.f, 0.f);
        return tu::load4(p + (((long)b * H + iy) * W + ix) * 64 + ci);
    }
};

// 8x8 patch gather over NHWC(64): m = (b, ty, tx) real tokens, k = (ky*8 + kx)*64 + ci.
// reflect != 0: rows/cols beyond H/W mirror as F.pad(mode='reflect') on the bottom/right.
template <typename T> struct APatch {
    const T *p;
    int M, H, W, Ht, Wt, reflect;
    __device__ __forceinline__ float4 load4(int m, int k) const {
        if (m >= M) return make_float4(0.f, 0.f, 0.f, 0.f);
        int ci = k & 63, kx = (k >> 6) & 7, ky = k >> 9;
        int tx = m % Wt, t = m / Wt;
        int ty = t % Ht, b = t / Ht;
        int y = ty * 8 + ky, x = tx * 8 + kx;
        if (reflect) {
            if (y >= H) y = 2 * (H - 1) - y;
            if (x >= W) x = 2 * (W - 1) - x;
        }
        return tu::load4(p + (((long)b * H + y) * W + x) * 64 + ci);
    }
};

// window-ordered row of token (b,ty,tx) in a (B, nWy, nWx, 64, dim) stream
__device__ __forceinline__ long window_row(int b, int ty, int tx, int nWy, int nWx) {
    return (((long)b * nWy + (ty >> 3)) * nWx + (tx >> 3)) * 64 + (ty & 7) * 8 + (tx & 7);
}

// fp32 token rows addressed by real-token index m = (b,ty,tx); window != 0 -> window order
struct ATokens {
    const float *p;
    int M, Ht, Wt, dim, window, nWy, nWx;
    __device__ __forceinline__ float4 load4(int m, int k) const {
        if (m >= M) return make_float4(0.f, 0.f, 0.f, 0.f);
        long row = m;
        if (window) {
            int tx = m % Wt, t = m / Wt;
            row = window_row(t / Ht, t % Ht, tx, nWy, nWx);
        }
        return tu::load4(p + row * dim + k);
    }
};

// ------------------------------------------------------------------ epilogue functors
// out[m][n] = act(acc + bias[n]) as T, row-major with leading dimension ld
template <typename T, int ACT /*0 none, 1 relu, 2 gelu*/> struct EpiStore {
    T *out;
    const float *bias;
    long ld;
    __device__ __forceinline__ void operator()(int m, int n, const float *v) const {
        float r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float t = v[j] + (bias ? bias[n + j] : 0.f);
            if (ACT == 1) t = fmaxf(t, 0.f);
            if (ACT == 2) t = gelu_erf(t);
            r[j] = t;
        }
        store4(out + (long)m * ld + n, make_float4(r[0], r[1], r[2], r[3]));
    }
};

// x[m][n] += acc + bias[n]   (fp32 residual stream)
struct EpiResidual {
    float *x;
    const float *bias;
    long ld;
    __device__ __forceinline__ void operator()(int m, int n, const float *v) const {
        float4 o = load4(x + (long)m * ld + n);
        o.x += v[0] + bias[n];
        o.y += v[1] + bias[n + 1];
        o.z += v[2] + bias[n + 2];
        o.w += v[3] + bias[n + 3];
        store4(x + (long)m * ld + n, o);
    }
};

// conv output: pixel m of (B,Ho,Wo), 64-channel chunk n/64; optional PixelShuffle(r) placement
template <typename T> struct EpiConv {
    T *out;
    const float *bias;
    int Ho, Wo, relu, r;   // r == 0: plain NHWC with C = nchunk*64 (nchunk given by ldc)
    int ldc;
    __device__ __forceinline__ void operator()(int m, int n, const float *v) const {
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float t = v[j] + (bias ? bias[n + j] : 0.f);
            o[j] = relu ? fmaxf(t, 0.f) : t;
        }
        long off;
        if (r == 0) {
            off = (long)m * ldc + n;
        } else {
            int ph = n >> 6, c = n & 63;
            int x = m % Wo, t = m / Wo;
            int y = t % Ho, b = t / Ho;
            off = ((((long)b * Ho + y) * r + ph / r) * ((long)Wo * r) + (long)x * r + ph % r) * 64 + c;
        }
        store4(out + off, make_float4(o[0], o[1], o[2], o[3]));
    }
};

// patch-embed epilogue: fp32 tokens, window-ordered (pad tokens stay zero) or row-major + pos_embed
struct EpiEmbed {
    float *tok;
    const float *bias, *pos;
    int Ht, Wt, dim, window, nWy, nWx;
    __device__ __forceinline__ void operator()(int m, int n, const float *v) const {
        int tx = m % Wt, t = m / Wt;
        int ty = t % Ht, b = t / Ht;
        long row = window ? window_row(b, ty, tx, nWy, nWx) : (long)m;
        float4 o = make_float4(v[0] + bias[n], v[1] + bias[n + 1], v[2] + bias[n + 2], v[3] + bias[n + 3]);
        if (pos) {
            float4 pe = load4(pos + (long)(ty * Wt + tx) * dim + n);
            o.x += pe.x; o.y += pe.y; o.z += pe.z; o.w += pe.w;
        }
        store4(tok + row * dim + n, o);
    }
};

// patch-unembed epilogue: n = (ky*8+kx)*64 + c -> pixel (8ty+ky, 8tx+kx), cropped to (Hc,Wc), + bias + skip
template <typename T> struct EpiUnembed {
    T *out;
    const float *bias;
    const T *skip;
    int Ht, Wt, Hc, Wc, skipH, skipW;
    __device__ __forceinline__ void operator()(int m, int n, const float *v) const {
        int c = n & 63, kx = (n >> 6) & 7, ky = n >> 9;
        int tx = m % Wt, t = m / Wt;
        int ty = t % Ht, b = t / Ht;
        int y = ty * 8 + ky, x = tx * 8 + kx;
        if (y >= Hc || x >= Wc) return;
        float4 s = load4(skip + (((long)b * skipH + y) * skipW + x) * 64 + c);
        float4 o = make_float4(v[0] + bias[c] + s.x, v[1] + bias[c + 1] + s.y, v[2] + bias[c + 2] + s.z,
                               v[3] + bias[c + 3] + s.w);
        store4(out + (((long)b * Hc + y) * Wc + x) * 64 + c, o);
    }
};

}  // namespace tu
