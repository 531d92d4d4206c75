// Pre-LN transformer block pieces on CUDA cores (exact fp32 math): LayerNorm, 8x8 window attention with
// dense relative-position bias, flash-style global attention, and the block driver that strings them
// together with the SIMT GEMMs.
// Reference: WindowAttention / WindowTransformerBlock (WindowTransformer/model.py:63-170,
// FastTransformer/model.py:65-172) and TransformerBlock with nn.MultiheadAttention
// (ResidualTransformer/model.py:22-50).
#include "gemm_simt.cuh"
#include "tc/tc_api.cuh"

namespace tu {

// ------------------------------------------------------------------ LayerNorm (eps 1e-5, biased variance)
// fp32 stream x (M, D) -> T (M, D).  One warp per token; D = 32 * PER.
template <typename T, int PER>
__global__ void __launch_bounds__(256) layernorm_kernel(const float *__restrict__ x, const float *__restrict__ g,
                                                        const float *__restrict__ b, T *__restrict__ out, int M) {
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= M) return;
    constexpr int D = 32 * PER;
    const float *row = x + (long)warp * D;
    float v[PER];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { v[i] = row[lane + 32 * i]; s += v[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { float d = v[i] - mean; q = fmaf(d, d, q); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q * (1.0f / D) + 1e-5f);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        int c = lane + 32 * i;
        out[(long)warp * D + c] = from_f<T>((v[i] - mean) * rstd * g[c] + b[c]);
    }
}

// ------------------------------------------------------------------ window attention (64 tokens, head_dim 16)
// qkv (nWin*64, 3*dim) T with q pre-scaled; out (nWin*64, dim) T.  One CTA per (window, head), one
// query per thread; K, V of the head in shared memory (broadcast reads); softmax over 64 keys in registers.
// bias[h][i][j] dense (heads, query, key) fp32.
template <typename T>
__global__ void __launch_bounds__(64) window_attn_kernel(const T *__restrict__ qkv, const float *__restrict__ bias_t,
                                                         T *__restrict__ out, int dim, int heads) {
    __shared__ float ks[64][16];
    __shared__ float vs[64][16];
    const int win = blockIdx.x, h = blockIdx.y, i = threadIdx.x;
    const T *base = qkv + ((long)win * 64 + i) * (3 * dim) + h * 16;
    float q[16];
#pragma unroll
    for (int d = 0; d < 16; d += 4) {
        float4 a = load4(base + d), k4 = load4(base + dim + d), v4 = load4(base + 2 * dim + d);
        q[d] = a.x; q[d + 1] = a.y; q[d + 2] = a.z; q[d + 3] = a.w;
        ks[i][d] = k4.x; ks[i][d + 1] = k4.y; ks[i][d + 2] = k4.z; ks[i][d + 3] = k4.w;
        vs[i][d] = v4.x; vs[i][d + 1] = v4.y; vs[i][d + 2] = v4.z; vs[i][d + 3] = v4.w;
    }
    __syncthreads();
    const float *bp = bias_t + ((long)h * 64 + i) * 64;
    float s[64];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < 16; ++d) a = fmaf(q[d], ks[j][d], a);
        a += bp[j];
        s[j] = a;
        mx = fmaxf(mx, a);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
    const float inv = 1.0f / sum;
    float o[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        float p = s[j] * inv;
#pragma unroll
        for (int d = 0; d < 16; ++d) o[d] = fmaf(p, vs[j][d], o[d]);
    }
    T *op = out + ((long)win * 64 + i) * dim + h * 16;
#pragma unroll
    for (int d = 0; d < 16; d += 4) store4(op + d, make_float4(o[d], o[d + 1], o[d + 2], o[d + 3]));
}

// ------------------------------------------------------------------ window attention on mma.sync (bf16)
// Same contract as window_attn_kernel for T = bf16.  CTA = (window, head), 4 warps x 16 query rows.
// S = Q K^T via 8 m16n8k16 MMAs per warp (K fragments straight from global), + bias, softmax in the accumulator
// fragments (quad shuffles), P re-used in place as the A operand of the 8 P V MMAs (V via ldmatrix.trans from
// shared memory), normalisation by the row sum in fp32 at the end.  64x64x16 problems are far too small for a
// 128-row tcgen05 tile, so the warp-level tensor path is the right tool for this 0.8 % of the FLOPs.
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// D = A B: the C operand is the constant zero, so the destination needs no initialisation (saves four moves per MMA)
__device__ __forceinline__ void mma_bf16_16816_z(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}

__global__ void __launch_bounds__(128) window_attn_mma_kernel(const bf16 *__restrict__ qkv, const float *__restrict__ bias,
                                                              bf16 *__restrict__ out, int dim) {
    __shared__ __align__(16) bf16 vs[64][24];     // V of this head, row pitch 48 B (conflict-free ldmatrix)
    const int win = blockIdx.x, h = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;
    const long ld = 3L * dim;
    const bf16 *base = qkv + (long)win * 64 * ld + h * 16;
    {   // stage V: 64 rows x 32 B
        const int r = threadIdx.x >> 1, hf = threadIdx.x & 1;
        *reinterpret_cast<uint4 *>(&vs[r][hf * 8]) = *reinterpret_cast<const uint4 *>(base + r * ld + 2 * dim + hf * 8);
    }
    const int r0 = warp * 16 + g;                 // this thread's query rows: r0 and r0 + 8
    uint32_t qa[4];
    qa[0] = *reinterpret_cast<const uint32_t *>(base + (long)r0 * ld + tq * 2);
    qa[1] = *reinterpret_cast<const uint32_t *>(base + (long)(r0 + 8) * ld + tq * 2);
    qa[2] = *reinterpret_cast<const uint32_t *>(base + (long)r0 * ld + tq * 2 + 8);
    qa[3] = *reinterpret_cast<const uint32_t *>(base + (long)(r0 + 8) * ld + tq * 2 + 8);
    float s[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        const bf16 *kp = base + (long)(n * 8 + g) * ld + dim + tq * 2;      // key j = n*8 + g, d = tq*2 (+8)
        const uint32_t b0 = *reinterpret_cast<const uint32_t *>(kp), b1 = *reinterpret_cast<const uint32_t *>(kp + 8);
        s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
        mma_bf16_16816(s[n], qa, b0, b1);
    }
    const float *bp0 = bias + ((long)h * 64 + r0) * 64 + tq * 2, *bp1 = bp0 + 8 * 64;
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        const float2 ba = *reinterpret_cast<const float2 *>(bp0 + n * 8), bb = *reinterpret_cast<const float2 *>(bp1 + n * 8);
        s[n][0] += ba.x; s[n][1] += ba.y; s[n][2] += bb.x; s[n][3] += bb.y;
        m0 = fmaxf(m0, fmaxf(s[n][0], s[n][1]));
        m1 = fmaxf(m1, fmaxf(s[n][2], s[n][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
    const float L2E = 1.4426950408889634f;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        s[n][0] = exp2f((s[n][0] - m0) * L2E); s[n][1] = exp2f((s[n][1] - m0) * L2E);
        s[n][2] = exp2f((s[n][2] - m1) * L2E); s[n][3] = exp2f((s[n][3] - m1) * L2E);
        l0 += s[n][0] + s[n][1];
        l1 += s[n][2] + s[n][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    __syncthreads();                               // V staged
    float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int t = 0; t < 4; ++t) {                  // 16 keys per step
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * t][0], s[2 * t][1]);
        pa[1] = pack_bf16x2(s[2 * t][2], s[2 * t][3]);
        pa[2] = pack_bf16x2(s[2 * t + 1][0], s[2 * t + 1][1]);
        pa[3] = pack_bf16x2(s[2 * t + 1][2], s[2 * t + 1][3]);
        // four 8x8 matrices: (keys 0-7 | 8-15) x (d 0-7 | 8-15), transposed on load -> B fragments
        uint32_t v0, v1, v2, v3;
        const uint32_t addr = (uint32_t)__cvta_generic_to_shared(&vs[t * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][(lane >> 4) * 8]);
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                     : "r"(addr));
        mma_bf16_16816(o[0], pa, v0, v1);
        mma_bf16_16816(o[1], pa, v2, v3);
    }
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    bf16 *op0 = out + ((long)win * 64 + r0) * dim + h * 16 + tq * 2, *op1 = op0 + 8L * dim;
    *reinterpret_cast<uint32_t *>(op0) = pack_bf16x2(o[0][0] * i0, o[0][1] * i0);
    *reinterpret_cast<uint32_t *>(op0 + 8) = pack_bf16x2(o[1][0] * i0, o[1][1] * i0);
    *reinterpret_cast<uint32_t *>(op1) = pack_bf16x2(o[0][2] * i1, o[0][3] * i1);
    *reinterpret_cast<uint32_t *>(op1 + 8) = pack_bf16x2(o[1][2] * i1, o[1][3] * i1);
}

// ------------------------------------------------------------------ global attention on the tensor cores (bf16, head_dim 16)
// Flash-style: CTA = 128 queries of one (frame, head), warp = 32 query rows (two m16 tiles sharing every K/V fragment);
// keys/values stream through shared memory in double-buffered tiles of 64 (cp.async) with an online softmax, so the
// S x S score matrix never exists.  QK^T is one m16n8k16 step per 8 keys (K = head_dim = 16), PV four steps per tile.
// Reference: nn.MultiheadAttention(128, 8, batch_first=True) inside TransformerBlock (ResidualTransformer/model.py:31,44);
// q is pre-scaled by head_dim^-0.5 in the packed in_proj weights.
constexpr int GA_KT = 64;          // keys per tile
constexpr int GA_PITCH = 24;       // bf16 per staged row (48 B: conflict-free fragment loads and ldmatrix)

__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int sz = valid ? 16 : 0;                 // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

// packed fp32 pairs and single-instruction exp2 for the attention inner loop (sm_100: FFMA2 / FADD2 are two IEEE fp32
// operations per issue slot; ex2.approx is 2 ulp, its results are rounded to bf16 for the PV product)
__device__ __forceinline__ uint64_t ga_pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void ga_up2(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ga_fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t ga_add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ga_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// GA_MT m16 query tiles per warp, GA_WARPS warps per CTA (128 queries per CTA either way).  Measured on cfg5a (2 frames, 8 layers,
// transformer blocks in total): GA_MT = 1 1.33 ms, 2 1.21 ms, 4 1.33 ms -- one tile per warp doubles the K/V fragment loads per
// query, four tiles leave too few warps per SM.
//
// BAL (balanced launch, few frames): the CTAs of the plain launch are all resident at once and each streams the whole key range, so
// the last wave -- the CTAs that do not fit -- starts when everything else ends (2 frames: 912 CTAs on 888 slots: + 20 %).  With
// BAL the (query tile, key tile) units of the launch are laid out in one sequence and cut into gridDim.x equal contiguous ranges, one
// per resident CTA (the decomposition of tc/global_attn_tcgen05.cu); a CTA that sees only part of a query tile's keys leaves an
// online-softmax partial (m, l, o[16]) per row in `scratch`, and global_attn_mma_merge_kernel combines the partials of those rows.
struct GaBal {
    float *scratch;          // [item][part][18][GA_QPB] fp32
    long long units;         // items * ktiles
    int ktiles, qtiles, heads, max_parts;
};
__host__ __device__ __forceinline__ int ga_bal_owner(long long u, long long U, int G) { return (int)(((u + 1) * G - 1) / U); }

template <int GA_MT, int GA_WARPS, bool BAL>
__global__ void __launch_bounds__(GA_WARPS * 32) global_attn_mma_kernel(const bf16 *__restrict__ qkv, bf16 *__restrict__ out, int S, int dim,
                                                                        const GaBal bal) {
    constexpr int GA_QPB = GA_WARPS * 16 * GA_MT;      // queries per CTA
    __shared__ __align__(16) bf16 ks[2][GA_KT][GA_PITCH];
    __shared__ __align__(16) bf16 vs[2][GA_KT][GA_PITCH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;
    const long ld = 3L * dim;
    const float L2E = 1.4426950408889634f;
    const int ntiles_all = (S + GA_KT - 1) / GA_KT;
    long long u = 0, u1 = 1;
    if (BAL) { u = bal.units * blockIdx.x / gridDim.x; u1 = bal.units * (blockIdx.x + 1) / gridDim.x; }
    for (; u < u1;) {
    int b, h, qt, kt0, kt1, item = 0;
    if (BAL) {
        item = (int)(u / bal.ktiles);
        kt0 = (int)(u - (long long)item * bal.ktiles);
        kt1 = (int)min((long long)bal.ktiles, kt0 + (u1 - u));
        qt = item % bal.qtiles;
        const int bh = item / bal.qtiles;
        h = bh % bal.heads; b = bh / bal.heads;
        u += kt1 - kt0;
    } else {
        b = blockIdx.z; h = blockIdx.y; qt = blockIdx.x; kt0 = 0; kt1 = ntiles_all;
        u = u1;
    }
    const bf16 *base = qkv + (long)b * S * ld + h * 16;
    const int q0 = qt * GA_QPB + warp * 16 * GA_MT;

    auto stage = [&](int buf, int k0) {
        for (int e = threadIdx.x; e < 2 * GA_KT; e += GA_WARPS * 32) {
            const int r = e >> 1, hf = e & 1;
            const int key = k0 + r;
            const bool ok = key < S;
            const bf16 *src = base + (long)(ok ? key : 0) * ld + hf * 8;
            cp_async16(&ks[buf][r][hf * 8], src + dim, ok);
            cp_async16(&vs[buf][r][hf * 8], src + 2 * dim, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // Q fragments of the two m16 tiles (rows beyond S read row S-1; their results are never stored)
    uint32_t qa[GA_MT][4];
#pragma unroll
    for (int mt = 0; mt < GA_MT; ++mt) {
        const int r0 = min(q0 + mt * 16 + g, S - 1), r1 = min(q0 + mt * 16 + g + 8, S - 1);
        qa[mt][0] = *reinterpret_cast<const uint32_t *>(base + (long)r0 * ld + tq * 2);
        qa[mt][1] = *reinterpret_cast<const uint32_t *>(base + (long)r1 * ld + tq * 2);
        qa[mt][2] = *reinterpret_cast<const uint32_t *>(base + (long)r0 * ld + tq * 2 + 8);
        qa[mt][3] = *reinterpret_cast<const uint32_t *>(base + (long)r1 * ld + tq * 2 + 8);
    }
    // o[mt][2] is the running row sum: P times a column of ones on the tensor core (the bf16 probabilities the PV product uses,
    // summed in fp32; every column of that accumulator tile holds the same sum, so no shuffles and no packed adds)
    float o[GA_MT][3][4], m[GA_MT][2];
#pragma unroll
    for (int mt = 0; mt < GA_MT; ++mt) {
        m[mt][0] = m[mt][1] = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) o[mt][nt][0] = o[mt][nt][1] = o[mt][nt][2] = o[mt][nt][3] = 0.f;
    }
    const int ntiles = kt1;
    stage(kt0 & 1, kt0 * GA_KT);
    for (int t = kt0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) {
            stage(buf ^ 1, (t + 1) * GA_KT);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const int nk = min(GA_KT, S - t * GA_KT);
        // K fragments of the tile: key n*8+g, dims tq*2.. (+8)
        uint32_t kb[8][2];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            kb[n][0] = *reinterpret_cast<const uint32_t *>(&ks[buf][n * 8 + g][tq * 2]);
            kb[n][1] = *reinterpret_cast<const uint32_t *>(&ks[buf][n * 8 + g][tq * 2 + 8]);
        }
        uint32_t vb[4][4];
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            const uint32_t addr = (uint32_t)__cvta_generic_to_shared(&vs[buf][kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][(lane >> 4) * 8]);
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(vb[kt][0]), "=r"(vb[kt][1]), "=r"(vb[kt][2]), "=r"(vb[kt][3])
                         : "r"(addr));
        }
#pragma unroll
        for (int mt = 0; mt < GA_MT; ++mt) {
            float s[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) mma_bf16_16816_z(s[n], qa[mt], kb[n][0], kb[n][1]);
            if (nk < GA_KT) {
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const int j = n * 8 + tq * 2;
                    if (j >= nk) s[n][0] = s[n][2] = -INFINITY;
                    if (j + 1 >= nk) s[n][1] = s[n][3] = -INFINITY;
                }
            }
            float t0 = -INFINITY, t1 = -INFINITY;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                t0 = fmaxf(t0, fmaxf(s[n][0], s[n][1]));
                t1 = fmaxf(t1, fmaxf(s[n][2], s[n][3]));
            }
            t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 1)); t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 2));
            t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 1)); t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 2));
            const float n0 = fmaxf(m[mt][0], t0), n1 = fmaxf(m[mt][1], t1);
            const float c0 = ga_ex2((m[mt][0] - n0) * L2E), c1 = ga_ex2((m[mt][1] - n1) * L2E);   // exp2(-inf) = 0 on the first tile
            m[mt][0] = n0; m[mt][1] = n1;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) { o[mt][nt][0] *= c0; o[mt][nt][1] *= c0; o[mt][nt][2] *= c1; o[mt][nt][3] *= c1; }
            o[mt][2][0] *= c0; o[mt][2][2] *= c1;
            // exponent arguments on packed fp32 pairs (FFMA2), one MUFU per exponential, row sums on packed pairs
            const uint64_t l2e2 = ga_pk2(L2E, L2E), nb0 = ga_pk2(-n0 * L2E, -n0 * L2E), nb1 = ga_pk2(-n1 * L2E, -n1 * L2E);
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                float e0, e1, e2, e3;
                ga_up2(ga_fma2(ga_pk2(s[n][0], s[n][1]), l2e2, nb0), e0, e1);
                ga_up2(ga_fma2(ga_pk2(s[n][2], s[n][3]), l2e2, nb1), e2, e3);
                s[n][0] = ga_ex2(e0); s[n][1] = ga_ex2(e1);
                s[n][2] = ga_ex2(e2); s[n][3] = ga_ex2(e3);
            }
#pragma unroll
            for (int kt = 0; kt < 4; ++kt) {
                uint32_t pa[4];
                pa[0] = pack_bf16x2(s[2 * kt][0], s[2 * kt][1]);
                pa[1] = pack_bf16x2(s[2 * kt][2], s[2 * kt][3]);
                pa[2] = pack_bf16x2(s[2 * kt + 1][0], s[2 * kt + 1][1]);
                pa[3] = pack_bf16x2(s[2 * kt + 1][2], s[2 * kt + 1][3]);
                mma_bf16_16816(o[mt][0], pa, vb[kt][0], vb[kt][1]);
                mma_bf16_16816(o[mt][1], pa, vb[kt][2], vb[kt][3]);
                mma_bf16_16816(o[mt][2], pa, 0x3f803f80u, 0x3f803f80u);
            }
        }
        __syncthreads();       // everyone is done with this buffer before the next prefetch overwrites it
    }
    if (BAL && (kt0 > 0 || kt1 < ntiles_all)) {
        // partial of this query tile: part index = distance from the first CTA that owns one of the tile's units
        const int part = (int)blockIdx.x - ga_bal_owner((long long)item * bal.ktiles, bal.units, (int)gridDim.x);
        float *dst = bal.scratch + ((long long)item * bal.max_parts + part) * (18 * GA_QPB);
#pragma unroll
        for (int mt = 0; mt < GA_MT; ++mt) {
            const int r0 = warp * 16 * GA_MT + mt * 16 + g, r1 = r0 + 8;
            if (tq == 0) {
                dst[r0] = m[mt][0]; dst[r1] = m[mt][1];
                dst[GA_QPB + r0] = o[mt][2][0]; dst[GA_QPB + r1] = o[mt][2][2];
            }
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int d = nt * 8 + tq * 2;
                dst[(2 + d) * GA_QPB + r0] = o[mt][nt][0]; dst[(3 + d) * GA_QPB + r0] = o[mt][nt][1];
                dst[(2 + d) * GA_QPB + r1] = o[mt][nt][2]; dst[(3 + d) * GA_QPB + r1] = o[mt][nt][3];
            }
        }
        continue;
    }
#pragma unroll
    for (int mt = 0; mt < GA_MT; ++mt) {
        const float i0 = 1.f / o[mt][2][0], i1 = 1.f / o[mt][2][2];
        const int r0 = q0 + mt * 16 + g, r1 = r0 + 8;
        bf16 *op0 = out + ((long)b * S + r0) * dim + h * 16 + tq * 2, *op1 = out + ((long)b * S + r1) * dim + h * 16 + tq * 2;
        if (r0 < S) {
            *reinterpret_cast<uint32_t *>(op0) = pack_bf16x2(o[mt][0][0] * i0, o[mt][0][1] * i0);
            *reinterpret_cast<uint32_t *>(op0 + 8) = pack_bf16x2(o[mt][1][0] * i0, o[mt][1][1] * i0);
        }
        if (r1 < S) {
            *reinterpret_cast<uint32_t *>(op1) = pack_bf16x2(o[mt][0][2] * i1, o[mt][0][3] * i1);
            *reinterpret_cast<uint32_t *>(op1 + 8) = pack_bf16x2(o[mt][1][2] * i1, o[mt][1][3] * i1);
        }
    }
    }      // segments
}

// CTA shape of the mma.sync kernel (debug key "ga_shape", -1 = automatic): 0 = 4 warps x 32 queries (128 queries per CTA), 1 = 2 warps x 32 queries
// (64 per CTA: twice as many, smaller CTAs spread more evenly over the SMs), 2 = 4 warps x 16, 3 = 8 warps x 16

// rows of the query tiles that were split between CTAs: combine their partials and write softmax(q k^T) v
template <int QPB>
__global__ void __launch_bounds__(QPB) global_attn_mma_merge_kernel(const GaBal bal, int grid_attn, bf16 *__restrict__ out, int S, int dim) {
    const int item = blockIdx.x, r = threadIdx.x;
    const int qt = item % bal.qtiles, bh = item / bal.qtiles, h = bh % bal.heads, b = bh / bal.heads;
    const int tok = qt * QPB + r;
    const int first = ga_bal_owner((long long)item * bal.ktiles, bal.units, grid_attn);
    const int last = ga_bal_owner((long long)item * bal.ktiles + bal.ktiles - 1, bal.units, grid_attn);
    if (first == last || tok >= S) return;          // written by its only CTA
    const int n = last - first + 1;
    const float *base = bal.scratch + (long long)item * bal.max_parts * (18 * QPB) + r;
    const float L2E = 1.4426950408889634f;
    float M = -INFINITY;
    for (int i = 0; i < n; ++i) M = fmaxf(M, base[(long long)i * 18 * QPB]);
    float L = 0.f, o[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) o[d] = 0.f;
    for (int i = 0; i < n; ++i) {
        const float *pp = base + (long long)i * 18 * QPB;
        const float f = ga_ex2((pp[0] - M) * L2E);
        L = fmaf(pp[QPB], f, L);
#pragma unroll
        for (int d = 0; d < 16; ++d) o[d] = fmaf(pp[(2 + d) * QPB], f, o[d]);
    }
    const float inv = 1.f / L;
    uint32_t w[8];
#pragma unroll
    for (int d = 0; d < 16; d += 2) w[d >> 1] = pack_bf16x2(o[d] * inv, o[d + 1] * inv);
    uint4 *op = reinterpret_cast<uint4 *>(out + ((long)b * S + tok) * dim + h * 16);
    op[0] = make_uint4(w[0], w[1], w[2], w[3]);
    op[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

}  // namespace tu
// CTA shape of the mma.sync kernel (debug key "ga_shape", -1 = automatic): 0 = 4 warps x 32 queries (128 queries per CTA), 1 = 2 warps
// x 32 queries, 2 = 4 warps x 16 (64 per CTA), 3 = 8 warps x 16, 4 = shape 2 with the balanced launch (needs scratch)
thread_local int tu::g_ga_shape = -1;
namespace tu {
// resident CTAs of the balanced kernel with MT m16 tiles per warp (a property of the code, not of the device ordinal)
template <int MT> static int ga_bal_grid() {
    static const int occ = [] {
        int o = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, global_attn_mma_kernel<MT, 4, true>, 128, 0) != cudaSuccess || o < 1) o = 4;
        return o;
    }();
    return device_sm_count() * occ;
}
static void ga_bal_geometry(int B, int S, int heads, int grid, int qpb, GaBal &g) {
    g.ktiles = ceil_div(S, GA_KT);
    g.qtiles = ceil_div(S, qpb);
    g.heads = heads;
    g.units = (long long)B * heads * g.qtiles * g.ktiles;
    const long long per = g.units / grid > 0 ? g.units / grid : 1;
    g.max_parts = (int)((g.ktiles + per - 1) / per) + 1;
}
size_t global_attn_mma_scratch_bytes(int B, int S, int heads) {
    GaBal g1, g2;
    ga_bal_geometry(B, S, heads, ga_bal_grid<1>(), 64, g1);
    ga_bal_geometry(B, S, heads, ga_bal_grid<2>(), 128, g2);
    const size_t a = (size_t)B * heads * g1.qtiles * g1.max_parts * 18 * 64 * sizeof(float);
    const size_t b = (size_t)B * heads * g2.qtiles * g2.max_parts * 18 * 128 * sizeof(float);
    return a > b ? a : b;
}
template <int MT>
static bool launch_global_attn_bal(const bf16 *qkv, bf16 *out, int B, int S, int heads, int dim, float *scratch, size_t scratch_bytes,
                                   cudaStream_t st) {
    constexpr int QPB = 64 * MT;
    const int grid = ga_bal_grid<MT>();
    GaBal g;
    ga_bal_geometry(B, S, heads, grid, QPB, g);
    g.scratch = scratch;
    if (!scratch || scratch_bytes < global_attn_mma_scratch_bytes(B, S, heads) || g.units < grid) return false;
    global_attn_mma_kernel<MT, 4, true><<<grid, 128, 0, st>>>(qkv, out, S, dim, g);
    global_attn_mma_merge_kernel<QPB><<<B * heads * g.qtiles, QPB, 0, st>>>(g, grid, out, S, dim);
    return true;
}
// scratch (optional): global_attn_mma_scratch_bytes() of fp32 partials for the balanced launches
static void launch_global_attn_mma(const bf16 *qkv, bf16 *out, int B, int S, int heads, int dim, float *scratch, size_t scratch_bytes,
                                   cudaStream_t st) {
    // measured (profiles/r2b_attn_shape.log, S = 3600): see DESIGN.md section 3.5c
    int shape = g_ga_shape;
    const GaBal none{nullptr, 0, 0, 0, 0, 0};
    if (shape < 0) {
        const long n128 = (long)ceil_div(S, 128) * heads * B;
        shape = n128 <= 8L * device_sm_count() ? 2 : 0;
        if (scratch && n128 <= 4L * device_sm_count()) shape = 5;
    }
    if (shape == 4 && launch_global_attn_bal<1>(qkv, out, B, S, heads, dim, scratch, scratch_bytes, st)) return;
    if (shape == 5 && launch_global_attn_bal<2>(qkv, out, B, S, heads, dim, scratch, scratch_bytes, st)) return;
    if (shape == 4) shape = 2;
    if (shape == 5) shape = 0;
    switch (shape) {
        case 1: global_attn_mma_kernel<2, 2, false><<<dim3(ceil_div(S, 64), heads, B), 64, 0, st>>>(qkv, out, S, dim, none); break;
        case 2: global_attn_mma_kernel<1, 4, false><<<dim3(ceil_div(S, 64), heads, B), 128, 0, st>>>(qkv, out, S, dim, none); break;
        case 3: global_attn_mma_kernel<1, 8, false><<<dim3(ceil_div(S, 128), heads, B), 256, 0, st>>>(qkv, out, S, dim, none); break;
        default: global_attn_mma_kernel<2, 4, false><<<dim3(ceil_div(S, 128), heads, B), 128, 0, st>>>(qkv, out, S, dim, none); break;
    }
}

// ------------------------------------------------------------------ global attention (S tokens per frame, head_dim 16)
// qkv (B*S, 3*dim) T with q pre-scaled; out (B*S, dim) T.  CTA = 128 queries of one (frame, head);
// keys/values streamed through shared memory in tiles of 128 with an online softmax (never materialises SxS).
template <typename T>
__global__ void __launch_bounds__(128) global_attn_kernel(const T *__restrict__ qkv, T *__restrict__ out, int S, int dim) {
    __shared__ float ks[128][16];
    __shared__ float vs[128][16];
    const int b = blockIdx.z, h = blockIdx.y;
    const int qi = blockIdx.x * 128 + threadIdx.x;
    const bool active = qi < S;
    const long frame = (long)b * S;
    float q[16], o[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) { q[d] = 0.f; o[d] = 0.f; }
    if (active) {
        const T *qp = qkv + (frame + qi) * (3 * dim) + h * 16;
#pragma unroll
        for (int d = 0; d < 16; d += 4) {
            float4 a = load4(qp + d);
            q[d] = a.x; q[d + 1] = a.y; q[d + 2] = a.z; q[d + 3] = a.w;
        }
    }
    float mx = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < S; k0 += 128) {
        __syncthreads();
        int kj = k0 + threadIdx.x;
        if (kj < S) {
            const T *kp = qkv + (frame + kj) * (3 * dim) + dim + h * 16;
#pragma unroll
            for (int d = 0; d < 16; d += 4) {
                float4 k4 = load4(kp + d), v4 = load4(kp + dim + d);
                ks[threadIdx.x][d] = k4.x; ks[threadIdx.x][d + 1] = k4.y; ks[threadIdx.x][d + 2] = k4.z; ks[threadIdx.x][d + 3] = k4.w;
                vs[threadIdx.x][d] = v4.x; vs[threadIdx.x][d + 1] = v4.y; vs[threadIdx.x][d + 2] = v4.z; vs[threadIdx.x][d + 3] = v4.w;
            }
        }
        __syncthreads();
        const int nk = min(128, S - k0);
        // two passes over the tile: tile max first, then one rescale of the running state
        float tmx = mx;
        for (int j = 0; j < nk; ++j) {
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < 16; ++d) a = fmaf(q[d], ks[j][d], a);
            tmx = fmaxf(tmx, a);
        }
        const float corr = expf(mx - tmx);   // exp(-inf) = 0 on the first tile
        l *= corr;
#pragma unroll
        for (int d = 0; d < 16; ++d) o[d] *= corr;
        mx = tmx;
        for (int j = 0; j < nk; ++j) {
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < 16; ++d) a = fmaf(q[d], ks[j][d], a);
            float p = expf(a - mx);
            l += p;
#pragma unroll
            for (int d = 0; d < 16; ++d) o[d] = fmaf(p, vs[j][d], o[d]);
        }
    }
    if (!active) return;
    const float inv = 1.0f / l;
    T *op = out + (frame + qi) * dim + h * 16;
#pragma unroll
    for (int d = 0; d < 16; d += 4)
        store4(op + d, make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv));
}

// ------------------------------------------------------------------ host launchers
template <typename T>
int launch_layernorm(const float *x, const float *g, const float *b, T *out, int M, int dim, cudaStream_t st) {
    dim3 grid(ceil_div(M, 8));
    if (dim == 128)
        layernorm_kernel<T, 4><<<grid, 256, 0, st>>>(x, g, b, out, M);
    else if (dim == 192)
        layernorm_kernel<T, 6><<<grid, 256, 0, st>>>(x, g, b, out, M);
    else {
        set_error("tu: layernorm supports dim 128 or 192");
        return TU_ERR_ARG;
    }
    TU_CHECK_LAUNCH("layernorm");
    return TU_OK;
}

// workspace: ln (M,dim) T | qkv (M,3dim) T [reused as mlp hidden (M,4dim) T] | attn (M,dim) T
static size_t block_ws(int M, int dim, int dtype) {
    size_t e = dtype_size(dtype);
    return align_up((size_t)M * dim * e, 256) + align_up((size_t)M * 4 * dim * e, 256) + align_up((size_t)M * dim * e, 256);
}

// the global attention kernels (tcgen05, or mma.sync with the balanced launch) keep fp32 partials behind the three buffers above (bf16 path, window == 0)
size_t block_workspace_bytes_ex(int M, int dim, int dtype, int window, int S) {
    size_t n = block_ws(M, dim, dtype);
    if (!window && dtype == TU_BF16 && S > 0 && M % S == 0 && tc_available()) {
        const size_t a = tc_global_attention_scratch_bytes(M / S, S, dim / 16), b = global_attn_mma_scratch_bytes(M / S, S, dim / 16);
        n += align_up(a > b ? a : b, 256);
    }
    return n;
}

// Dense layer of the block: tensor-core GEMM for bf16 when enabled, else the exact-fp32 SIMT GEMM.
//   mode 0: out = A W^T + b   mode 2: out = gelu(A W^T + b)   mode 1: x += A W^T + b (fp32 stream)
template <typename T>
static int block_linear(const T *A, long lda, const void *W, const float *bias, int M, int N, int K, int mode, T *out, float *x,
                        bf16 *x_bf16, bool tc, cudaStream_t st, const char *what) {
    if (tc && sizeof(T) == 2 && lda == K) {
        int rc = tc_linear((const bf16 *)A, (const bf16 *)W, bias, M, N, K, mode == 2 ? 2 : 0, mode == 1 ? nullptr : (bf16 *)out,
                           mode == 1 ? x : nullptr, mode == 1 ? x_bf16 : nullptr, st);
        if (rc != TU_TC_UNSUPPORTED) return rc;
    }
    ARows<T> a{A, M, lda};
    WDesc<T> wd{(const T *)W, K, 0, 64L * K};
    if (mode == 1) {
        EpiResidual e{x, bias, N};
        return launch_gemm_simt<T>(a, wd, M, N, K, e, st, what);
    }
    if (mode == 2) {
        EpiStore<T, 2> e{out, bias, N};
        return launch_gemm_simt<T>(a, wd, M, N, K, e, st, what);
    }
    EpiStore<T, 0> e{out, bias, N};
    return launch_gemm_simt<T>(a, wd, M, N, K, e, st, what);
}

// x_bf16_out (optional, tensor-core path only): receives a bf16 copy of the block's output stream
template <typename T>
int transformer_block_impl(float *x, const TuBlockWeights *w, int M, int dim, int heads, int window, int S, void *ws, size_t ws_bytes,
                           bool tc, bf16 *x_bf16_out, cudaStream_t st) {
    char *p = (char *)ws;
    T *ln = (T *)p;
    p += align_up((size_t)M * dim * sizeof(T), 256);
    T *big = (T *)p;
    p += align_up((size_t)M * 4 * dim * sizeof(T), 256);
    T *att = (T *)p;
    int rc;
    if ((rc = launch_layernorm<T>(x, w->ln1_w, w->ln1_b, ln, M, dim, st))) return rc;
    if ((rc = block_linear<T>(ln, dim, w->qkv_w, w->qkv_b, M, 3 * dim, dim, 0, big, nullptr, nullptr, tc, st, "qkv"))) return rc;
    if (window && tc && sizeof(T) == 2) {
        dim3 grid(M / 64, heads);
        window_attn_mma_kernel<<<grid, 128, 0, st>>>((const bf16 *)big, w->rel_bias, (bf16 *)att, dim);
        TU_CHECK_LAUNCH("window_attn_mma");
    } else if (window) {
        dim3 grid(M / 64, heads);
        window_attn_kernel<T><<<grid, 64, 0, st>>>(big, w->rel_bias, att, dim, heads);
        TU_CHECK_LAUNCH("window_attn");
    } else if (tc && sizeof(T) == 2) {
        // tcgen05 flash attention when the caller's workspace holds its partials (block_workspace_bytes_ex); V^T lives in the unused
        // last quarter of the qkv / hidden buffer.  Otherwise the mma.sync kernel.
        rc = TU_TC_UNSUPPORTED;
        const size_t base = block_ws(M, dim, TU_BF16);
        if (ws_bytes > base)
            rc = tc_global_attention((const bf16 *)big, (bf16 *)att, (bf16 *)big + (size_t)M * 3 * dim, (float *)((char *)ws + base), ws_bytes - base,
                                     M / S, S, heads, st);
        if (rc != TU_OK && rc != TU_TC_UNSUPPORTED) return rc;
        if (rc == TU_TC_UNSUPPORTED) {
            launch_global_attn_mma((const bf16 *)big, (bf16 *)att, M / S, S, heads, dim, ws_bytes > base ? (float *)((char *)ws + base) : nullptr,
                                   ws_bytes > base ? ws_bytes - base : 0, st);
            TU_CHECK_LAUNCH("global_attn_mma");
        }
    } else {
        dim3 grid(ceil_div(S, 128), heads, M / S);
        global_attn_kernel<T><<<grid, 128, 0, st>>>(big, att, S, dim);
        TU_CHECK_LAUNCH("global_attn");
    }
    if ((rc = block_linear<T>(att, dim, w->proj_w, w->proj_b, M, dim, dim, 1, nullptr, x, nullptr, tc, st, "proj"))) return rc;
    if ((rc = launch_layernorm<T>(x, w->ln2_w, w->ln2_b, ln, M, dim, st))) return rc;
    if ((rc = block_linear<T>(ln, dim, w->fc1_w, w->fc1_b, M, 4 * dim, dim, 2, big, nullptr, nullptr, tc, st, "fc1"))) return rc;
    if ((rc = block_linear<T>(big, 4L * dim, w->fc2_w, w->fc2_b, M, dim, 4 * dim, 1, nullptr, x, x_bf16_out, tc, st, "fc2"))) return rc;
    return TU_OK;
}

template int transformer_block_impl<float>(float *, const TuBlockWeights *, int, int, int, int, int, void *, size_t, bool, bf16 *, cudaStream_t);
template int transformer_block_impl<bf16>(float *, const TuBlockWeights *, int, int, int, int, int, void *, size_t, bool, bf16 *, cudaStream_t);

// One ResidualTransformer layer: resid_pre (LN1 + in_proj) -> global attention -> resid_post (out_proj, LN2, MLP), bf16 tensor-core path.
// Workspace layout as in transformer_block_impl: ln (unused here) | qkv in the 4 dim buffer (+ V^T in its last quarter) | att [| partials]
int resid_layer_fused(float *x, const TuModelWeights *w, int layer, int M, int S, void *ws, size_t ws_bytes, bf16 *x_bf16_out,
                      cudaStream_t st) {
    const int dim = w->dim, heads = w->heads;
    if (dim != 128 || !w->stack_w || !w->stack_p || !tc_enabled() || S <= 0 || M % S) return TU_TC_UNSUPPORTED;
    const size_t base = block_ws(M, dim, TU_BF16);
    if (ws_bytes < base) return TU_TC_UNSUPPORTED;
    char *p = (char *)ws;
    p += align_up((size_t)M * dim * sizeof(bf16), 256);
    bf16 *big = (bf16 *)p;
    p += align_up((size_t)M * 4 * dim * sizeof(bf16), 256);
    bf16 *att = (bf16 *)p;
    int rc = tc_resid_pre(x, big, M, layer, w->n_blocks, (const bf16 *)w->stack_w, w->stack_p, st);
    if (rc) return rc;             // TU_TC_UNSUPPORTED before anything was launched, or an error
    rc = TU_TC_UNSUPPORTED;
    if (ws_bytes > base)
        rc = tc_global_attention(big, att, big + (size_t)M * 3 * dim, (float *)((char *)ws + base), ws_bytes - base, M / S, S, heads, st);
    if (rc != TU_OK && rc != TU_TC_UNSUPPORTED) return rc;
    if (rc == TU_TC_UNSUPPORTED) {
        launch_global_attn_mma(big, att, M / S, S, heads, dim, ws_bytes > base ? (float *)((char *)ws + base) : nullptr, ws_bytes > base ? ws_bytes - base : 0, st);
        TU_CHECK_LAUNCH("global_attn_mma");
    }
    rc = tc_resid_post(x, x_bf16_out, att, M, layer, w->n_blocks, (const bf16 *)w->stack_w, w->stack_p, st);
    if (rc == TU_TC_UNSUPPORTED) {
        set_error("tu: resid_post unavailable after resid_pre ran");
        return TU_ERR_ARG;
    }
    return rc;
}

static int block_check(float *x, const TuBlockWeights *w, int M, int dim, int heads, int window, int S, int dtype, void *workspace,
                       size_t workspace_bytes) {
    TU_CHECK_ARG(x && w && workspace && M > 0, "transformer_block: bad argument");
    TU_CHECK_ARG(dim == heads * 16 && (dim == 128 || dim == 192), "transformer_block: dim must be heads*16 and 128|192");
    TU_CHECK_ARG(window ? (M % 64 == 0 && w->rel_bias) : (S > 0 && M % S == 0), "transformer_block: bad token count");
    TU_CHECK_ARG(dtype == TU_F32 || dtype == TU_BF16, "transformer_block: bad dtype");
    if (workspace_bytes < block_ws(M, dim, dtype)) {
        set_error("tu: transformer_block workspace too small");
        return TU_ERR_WORKSPACE;
    }
    return TU_OK;
}

int transformer_block_ex(float *x, const TuBlockWeights *w, int M, int dim, int heads, int window, int S, int dtype,
                         void *workspace, size_t workspace_bytes, bf16 *x_bf16_out, cudaStream_t st) {
    int rc = block_check(x, w, M, dim, heads, window, S, dtype, workspace, workspace_bytes);
    if (rc) return rc;
    if (dtype == TU_F32) return transformer_block_impl<float>(x, w, M, dim, heads, window, S, workspace, workspace_bytes, false, nullptr, st);
    return transformer_block_impl<bf16>(x, w, M, dim, heads, window, S, workspace, workspace_bytes, tc_enabled(), x_bf16_out, st);
}

}  // namespace tu

using namespace tu;

extern "C" size_t tu_block_workspace_bytes(int M, int dim, int dtype) { return block_ws(M, dim, dtype); }
extern "C" size_t tu_block_workspace_bytes_for(int M, int dim, int dtype, int window, int S) { return block_workspace_bytes_ex(M, dim, dtype, window, S); }

// Stand-alone window attention (softmax(q k^T + bias) v per 8x8 window and head, head_dim 16) on qkv rows (nWin*64, 3*dim):
// the op inside WindowAttention.forward between `qkv` and `proj` (WindowTransformer/model.py:104-127).  Exported so that the
// "attention TFLOP/s" of the headline metric can be measured on its own; the forward pass itself runs it fused
// (window_stack_tcgen05.cu) or, for dim 192, through tu_transformer_block.
extern "C" int tu_window_attention(const void *qkv, const float *rel_bias, void *out, int nWin, int dim, int heads, int dtype,
                                   void *stream) {
    TU_CHECK_ARG(qkv && rel_bias && out && nWin > 0, "window_attention: bad argument");
    TU_CHECK_ARG(dim == heads * 16 && (dim == 128 || dim == 192), "window_attention: dim must be heads*16 and 128|192");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(nWin, heads);
    if (dtype == TU_BF16) {
        window_attn_mma_kernel<<<grid, 128, 0, st>>>((const bf16 *)qkv, rel_bias, (bf16 *)out, dim);
    } else if (dtype == TU_F32) {
        window_attn_kernel<float><<<grid, 64, 0, st>>>((const float *)qkv, rel_bias, (float *)out, dim, heads);
    } else {
        TU_CHECK_ARG(false, "window_attention: bad dtype");
    }
    TU_CHECK_LAUNCH("window_attention");
    return TU_OK;
}

// Stand-alone global attention (ResidualTransformer: softmax(q k^T) v over the S tokens of a frame per head, head_dim 16) on qkv rows
// (B*S, 3*dim) bf16 with q pre-scaled; out (B*S, dim) bf16.  workspace: tu_global_attention_workspace_bytes(B, S, heads) enables the
// tcgen05 kernel; with workspace == NULL (or too small / unsupported S) the mma.sync kernel runs.  Exported so that the "attention
// TFLOP/s" of the headline metric can be measured on its own for this model too.
extern "C" size_t tu_global_attention_workspace_bytes(int B, int S, int heads) {
    if (B <= 0 || S <= 0 || heads <= 0 || !tc_available()) return 0;
    const size_t a = tc_global_attention_scratch_bytes(B, S, heads), b = global_attn_mma_scratch_bytes(B, S, heads);
    return align_up((size_t)B * S * heads * 16 * sizeof(bf16), 256) + (a > b ? a : b);
}
extern "C" int tu_global_attention(const void *qkv, void *out, int B, int S, int heads, void *workspace, size_t workspace_bytes, void *stream) {
    TU_CHECK_ARG(qkv && out && B > 0 && S > 0 && (heads == 8 || heads == 12), "global_attention: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int dim = heads * 16;
    int rc = TU_TC_UNSUPPORTED;
    const size_t vt_bytes = align_up((size_t)B * S * dim * sizeof(bf16), 256);
    if (workspace && workspace_bytes > vt_bytes && tc_enabled())
        rc = tc_global_attention((const bf16 *)qkv, (bf16 *)out, (bf16 *)workspace, (float *)((char *)workspace + vt_bytes), workspace_bytes - vt_bytes,
                                 B, S, heads, st);
    if (rc != TU_TC_UNSUPPORTED) return rc;
    launch_global_attn_mma((const bf16 *)qkv, (bf16 *)out, B, S, heads, dim, workspace && workspace_bytes > vt_bytes ? (float *)((char *)workspace + vt_bytes) : nullptr,
                           workspace && workspace_bytes > vt_bytes ? workspace_bytes - vt_bytes : 0, st);
    TU_CHECK_LAUNCH("global_attn_mma");
    return TU_OK;
}

extern "C" int tu_transformer_block(float *x, const TuBlockWeights *w, int M, int dim, int heads, int window, int S,
                                    int dtype, void *workspace, size_t workspace_bytes, void *stream) {
    return transformer_block_ex(x, w, M, dim, heads, window, S, dtype, workspace, workspace_bytes, nullptr, (cudaStream_t)stream);
}
