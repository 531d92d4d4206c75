// ResidualTransformer's global attention on the 5th-generation tensor cores: softmax(q k^T) v over all S = 3600 tokens of a frame
// per head (head_dim 16), nn.MultiheadAttention(128, 8, batch_first=True) inside TransformerBlock (ResidualTransformer/model.py:31,44);
// q arrives pre-scaled by head_dim^-0.5 (packing.py).  The S x S score matrix never exists.
//
// Work decomposition.  An "item" is (frame, head, pair of 128-query tiles); its key loop has ceil(S / 128) tiles.  All (item, key
// tile) units of the launch are laid out in one sequence and cut into gridDim.x equal contiguous ranges, one per persistent CTA, so
// every SM does the same number of 128 x 128 score tiles whatever the number of items (240 items on 148 SMs would otherwise run
// two rounds, the second 62 % full).  A CTA therefore processes "segments" (item, kt0..kt1) and leaves an online-softmax PARTIAL
// (row max m, row sum l, un-normalised o[16]) per query row in a scratch buffer; `global_attn_merge_kernel` combines the (at most a
// few) partials of a row and writes the normalised bf16 output.  A partial of the whole key range merges to itself.
//
// CTA = 12 warps:
//   warp 0      TMA producer: the two Q tiles of a segment, then per key tile the K slab and the V^T slabs (3-stage ring)
//   warp 1      MMA issuer: S_w = Q_w K^T (one 128x128x16 UMMA per tile and query tile), O_w = P_w V (eight 128x16x16 UMMAs)
//   warps 4-7   softmax warpgroup 0 (query tile 0), warps 8-11 softmax warpgroup 1 (query tile 1): a thread owns one query row
//               (its TMEM lane): row max over the tile, alpha = 2^((m_old - m_new) c), P = 2^(S c - m_new c) rounded to bf16 into
//               a 128-byte-swizzled K-major shared-memory operand, o = (o + O_prev) alpha in registers.
// The two warpgroups run half a tile apart: while one waits for its next S tile or its PV product the other keeps the MUFU busy
// (the kernel is bound by the exponentials: 16 per clock and SM).
//
// Operand layouts: only the K-major 128-byte-swizzle descriptors every other kernel of this library uses.  q / k of FOUR heads are
// one 128-byte row of the qkv matrix, so the TMA box is (64 columns, 128 rows) and head h of the slab is the K = 16 slice at byte
// offset 32 (h mod 4) of every row -- the same descriptor advance a K loop over a 64-wide slab takes.  V is needed K-major in the
// key dimension: `global_attn_vt_kernel` writes V^T (frame, head, d) x S once per layer (1.8 MB for two frames), loaded as two
// (64 keys, 16 rows) boxes per tile.  Keys past S are zero-filled by TMA (V^T) and masked in the softmax (scores).
#include <cuda.h>

#include "ptx.cuh"
#include "tc_api.cuh"

namespace tu {

namespace {

constexpr int GA_THREADS = 384;
constexpr int GA_KV_STAGES = 3;
constexpr int GA_Q_BYTES = 128 * 128;            // one query tile: 128 rows x 64 bf16 (four heads)
constexpr int GA_K_BYTES = 128 * 128;            // one key tile, same shape
constexpr int GA_VT_SLAB = 16 * 128;             // 16 rows (d) x 64 keys
constexpr int GA_KV_STAGE = GA_K_BYTES + 2 * GA_VT_SLAB;
constexpr int GA_P_SLAB = 128 * 128;             // 128 rows x 64 keys bf16
constexpr int GA_P_BYTES = 2 * GA_P_SLAB;        // per warpgroup
constexpr int GA_OFF_Q = 0;
constexpr int GA_OFF_KV = GA_OFF_Q + 2 * GA_Q_BYTES;
constexpr int GA_OFF_P = GA_OFF_KV + GA_KV_STAGES * GA_KV_STAGE;
constexpr int GA_OFF_BAR = GA_OFF_P + 2 * GA_P_BYTES;
constexpr int GA_SMEM = GA_OFF_BAR + 256 + 1024;
constexpr int GA_TMEM_COLS = 512;                // S0 [0,128) S1 [128,256) O0 [256,272) O1 [272,288)
constexpr int GA_PART_FLOATS = 18;               // m, l, o[16]
static_assert(GA_OFF_KV % 1024 == 0 && GA_OFF_P % 1024 == 0 && GA_KV_STAGE % 1024 == 0 && GA_SMEM <= 232448, "shared memory layout");

struct GaParams {
    int B, S, heads, dim;
    int qpairs;              // ceil(S / 256)
    int ktiles;              // ceil(S / 128)
    long long units;         // B * heads * qpairs * ktiles
    int max_parts;
    float *scratch;          // [item][part][18][256]
};

struct GaBars {
    uint64_t q_full, q_free;
    uint64_t kv_full[GA_KV_STAGES], kv_empty[GA_KV_STAGES];
    uint64_t s_full[2], p_full[2], o_full[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// first CTA whose range of units contains unit u: ranges start at floor(U i / G)
__host__ __device__ __forceinline__ int ga_owner(long long u, long long U, int G) { return (int)(((u + 1) * G - 1) / U); }
__host__ __device__ __forceinline__ long long ga_range_start(int i, long long U, int G) { return U * i / G; }

struct Segment { int item, b, h, qp, kt0, kt1, part, act1; };
__device__ __forceinline__ bool next_segment(const GaParams &p, long long &u, long long u1, Segment &s) {
    if (u >= u1) return false;
    s.item = (int)(u / p.ktiles);
    s.kt0 = (int)(u - (long long)s.item * p.ktiles);
    const long long left = u1 - u;
    s.kt1 = (int)min((long long)p.ktiles, s.kt0 + left);
    s.qp = s.item % p.qpairs;
    const int bh = s.item / p.qpairs;
    s.h = bh % p.heads;
    s.b = bh / p.heads;
    s.part = (int)blockIdx.x - ga_owner((long long)s.item * p.ktiles, p.units, (int)gridDim.x);
    s.act1 = s.qp * 256 + 128 < p.S;          // the second query tile holds at least one real row
    u += s.kt1 - s.kt0;
    return true;
}

__global__ void __launch_bounds__(GA_THREADS, 1)
global_attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_vt, const GaParams p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *sm = smem_raw + (smem0 - ptx::smem_u32(smem_raw));
    GaBars *bars = reinterpret_cast<GaBars *>(sm + GA_OFF_BAR);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        ptx::mbar_init(ptx::smem_u32(&bars->q_full), 1);
        ptx::mbar_init(ptx::smem_u32(&bars->q_free), 1);
        for (int i = 0; i < GA_KV_STAGES; ++i) {
            ptx::mbar_init(ptx::smem_u32(&bars->kv_full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->kv_empty[i]), 1);
        }
        for (int w = 0; w < 2; ++w) {
            ptx::mbar_init(ptx::smem_u32(&bars->s_full[w]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->p_full[w]), 128);
            ptx::mbar_init(ptx::smem_u32(&bars->o_full[w]), 1);
        }
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&tmap_qkv);
        ptx::prefetch_tmap(&tmap_vt);
    }
    // P starts as zeros: key chunks past S are never written and must multiply the zero-filled V^T as finite numbers
    for (int i = threadIdx.x; i < 2 * GA_P_BYTES / 16; i += GA_THREADS) reinterpret_cast<uint4 *>(sm + GA_OFF_P)[i] = make_uint4(0, 0, 0, 0);
    ptx::fence_proxy_async();
    if (warp == 1) {
        ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), GA_TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_wait();

    const long long u_begin = ga_range_start((int)blockIdx.x, p.units, (int)gridDim.x);
    const long long u_end = ga_range_start((int)blockIdx.x + 1, p.units, (int)gridDim.x);
    const int qcol_per_slab = 64;

    if (warp == 0) {
        if (lane == 0) {
            // ================================ TMA producer ================================
            int stage = 0;
            uint32_t phase = 0, qphase = 0;
            long long u = u_begin;
            Segment s;
            while (next_segment(p, u, u_end, s)) {
                ptx::mbar_wait(ptx::smem_u32(&bars->q_free), qphase ^ 1);
                const uint32_t qf = ptx::smem_u32(&bars->q_full);
                ptx::mbar_expect_tx(qf, (s.act1 ? 2 : 1) * GA_Q_BYTES);
                const int slab = s.h >> 2;
                tma_load_3d(smem0 + GA_OFF_Q, &tmap_qkv, qf, slab * qcol_per_slab, s.qp * 256, s.b);
                if (s.act1) tma_load_3d(smem0 + GA_OFF_Q + GA_Q_BYTES, &tmap_qkv, qf, slab * qcol_per_slab, s.qp * 256 + 128, s.b);
                qphase ^= 1;
                for (int kt = s.kt0; kt < s.kt1; ++kt) {
                    ptx::mbar_wait(ptx::smem_u32(&bars->kv_empty[stage]), phase ^ 1);
                    const uint32_t fb = ptx::smem_u32(&bars->kv_full[stage]);
                    const uint32_t dst = smem0 + GA_OFF_KV + stage * GA_KV_STAGE;
                    ptx::mbar_expect_tx(fb, GA_KV_STAGE);
                    tma_load_3d(dst, &tmap_qkv, fb, p.dim + slab * qcol_per_slab, kt * 128, s.b);
                    const int vrow = (s.b * p.heads + s.h) * 16;
                    ptx::tma_load_2d(dst + GA_K_BYTES, &tmap_vt, fb, kt * 128, vrow);
                    ptx::tma_load_2d(dst + GA_K_BYTES + GA_VT_SLAB, &tmap_vt, fb, kt * 128 + 64, vrow);
                    if (++stage == GA_KV_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (whole warp converged, elected lane issues) ================================
        const uint32_t leader = ptx::elect_one();
        const uint32_t idesc_qk = ptx::make_idesc_bf16(128, 128), idesc_pv = ptx::make_idesc_bf16(128, 16);
        int stage = 0;
        uint32_t phase = 0, qphase = 0, pph[2] = {0u, 0u};
        long long u = u_begin;
        Segment s;
        while (next_segment(p, u, u_end, s)) {
            const int T = s.kt1 - s.kt0, nw = s.act1 ? 2 : 1;
            const uint32_t hoff = (uint32_t)(s.h & 3) * 2u;                    // 32 bytes per head inside the four-head slab
            ptx::mbar_wait(ptx::smem_u32(&bars->q_full), qphase);
            qphase ^= 1;
            ptx::mbar_wait(ptx::smem_u32(&bars->kv_full[stage]), phase);
            ptx::tc_fence_after();
            for (int w = 0; w < nw; ++w) {
                ptx::umma_bf16_lo<0>(tmem_base + w * 128, ptx::sdesc_lo(smem0 + GA_OFF_Q + w * GA_Q_BYTES) + hoff,
                                     ptx::sdesc_lo(smem0 + GA_OFF_KV + stage * GA_KV_STAGE) + hoff, idesc_qk, leader);
                ptx::umma_commit_pred(ptx::smem_u32(&bars->s_full[w]), leader);
            }
            for (int t = 0; t < T; ++t) {
                int nstage = stage + 1;
                uint32_t nphase = phase;
                if (nstage == GA_KV_STAGES) { nstage = 0; nphase ^= 1; }
                if (t + 1 < T) ptx::mbar_wait(ptx::smem_u32(&bars->kv_full[nstage]), nphase);
                const uint32_t vt_lo = ptx::sdesc_lo(smem0 + GA_OFF_KV + stage * GA_KV_STAGE + GA_K_BYTES);
                for (int w = 0; w < nw; ++w) {
                    ptx::mbar_wait(ptx::smem_u32(&bars->p_full[w]), pph[w]);   // P_w(t) written, S_w(t) and O_w(t-1) read
                    pph[w] ^= 1;
                    ptx::tc_fence_after();
                    const uint32_t p_lo = ptx::sdesc_lo(smem0 + GA_OFF_P + w * GA_P_BYTES);
                    const uint32_t od = tmem_base + 256 + w * 16;
                    ptx::umma_bf16_lo<0>(od, p_lo, vt_lo, idesc_pv, leader);
#pragma unroll
                    for (int ks = 1; ks < 8; ++ks)
                        ptx::umma_bf16_lo<1>(od, p_lo + (ks >> 2) * (GA_P_SLAB >> 4) + (ks & 3) * 2, vt_lo + (ks >> 2) * (GA_VT_SLAB >> 4) + (ks & 3) * 2,
                                             idesc_pv, leader);
                    ptx::umma_commit_pred(ptx::smem_u32(&bars->o_full[w]), leader);
                    if (t + 1 < T) {
                        ptx::umma_bf16_lo<0>(tmem_base + w * 128, ptx::sdesc_lo(smem0 + GA_OFF_Q + w * GA_Q_BYTES) + hoff,
                                             ptx::sdesc_lo(smem0 + GA_OFF_KV + nstage * GA_KV_STAGE) + hoff, idesc_qk, leader);
                        ptx::umma_commit_pred(ptx::smem_u32(&bars->s_full[w]), leader);
                    }
                }
                ptx::umma_commit_pred(ptx::smem_u32(&bars->kv_empty[stage]), leader);
                stage = nstage;
                phase = nphase;
            }
            ptx::umma_commit_pred(ptx::smem_u32(&bars->q_free), leader);
        }
    } else if (warp >= 4) {
        // ================================ softmax warpgroups ================================
        const int w = (warp - 4) >> 2, q = warp & 3, row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t s_addr = tmem_base + lane_off + w * 128, o_addr = tmem_base + lane_off + 256 + w * 16;
        uint8_t *prow = sm + GA_OFF_P + w * GA_P_BYTES + row * 128;
        const uint32_t sw = (uint32_t)(row & 7);
        const float L2E = 1.4426950408889634f;
        uint32_t sph = 0, oph = 0;
        long long u = u_begin;
        Segment s;
        while (next_segment(p, u, u_end, s)) {
            if (w == 1 && !s.act1) continue;
            const int T = s.kt1 - s.kt0;
            float m = -INFINITY, l = 0.f, o[16];
#pragma unroll
            for (int d = 0; d < 16; ++d) o[d] = 0.f;
            for (int t = 0; t < T; ++t) {
                const int nvalid = min(128, p.S - (s.kt0 + t) * 128);
                ptx::mbar_wait(ptx::smem_u32(&bars->s_full[w]), sph);
                sph ^= 1;
                ptx::tc_fence_after();
                // ---- pass A: row max of the tile
                float mx = -INFINITY;
                if (nvalid == 128) {
#pragma unroll
                    for (int c = 0; c < 4; c += 2) {
                        uint32_t v0[32], v1[32];
                        ptx::tmem_ld_x32(s_addr + c * 32, v0);
                        ptx::tmem_ld_x32(s_addr + c * 32 + 32, v1);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            mx = fmaxf(mx, fmaxf(__uint_as_float(v0[j]), __uint_as_float(v0[j + 1])));
                            mx = fmaxf(mx, fmaxf(__uint_as_float(v1[j]), __uint_as_float(v1[j + 1])));
                        }
                    }
                } else {
                    for (int c = 0; c * 32 < nvalid; ++c) {
                        uint32_t v0[32];
                        ptx::tmem_ld_x32(s_addr + c * 32, v0);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c * 32 + j < nvalid) mx = fmaxf(mx, __uint_as_float(v0[j]));
                    }
                }
                const float m_new = fmaxf(m, mx);
                const float alpha = ex2f((m - m_new) * L2E);                   // 2^(-inf) = 0 on a segment's first tile
                if (t > 0) {
                    ptx::mbar_wait(ptx::smem_u32(&bars->o_full[w]), oph);       // O_w(t-1) = P_w(t-1) V(t-1): also frees the P buffer
                    oph ^= 1;
                    ptx::tc_fence_after();
                    uint32_t ov[16];
                    ptx::tmem_ld_x16(o_addr, ov);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int d = 0; d < 16; ++d) o[d] = (o[d] + __uint_as_float(ov[d])) * alpha;
                }
                l *= alpha;
                m = m_new;
                // ---- pass B: P = 2^(S c - m c) -> bf16, 128-byte swizzled K-major rows of two 64-key slabs
                const ptx::f32x2 c2 = ptx::pk2(L2E, L2E), nb2 = ptx::pk2(-m_new * L2E, -m_new * L2E);
                ptx::f32x2 rs = ptx::pk2(0.f, 0.f);
                for (int c = 0; c * 32 < nvalid; ++c) {
                    uint32_t v[32];
                    ptx::tmem_ld_x32(s_addr + c * 32, v);
                    ptx::tmem_ld_wait();
                    uint32_t pk[16];
                    const int lim = nvalid - c * 32;                         // >= 32 for a full chunk
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        float e0, e1;
                        ptx::up2(ptx::fma2(ptx::pk2u(v[j], v[j + 1]), c2, nb2), e0, e1);
                        float p0 = ex2f(e0), p1 = ex2f(e1);
                        if (lim < 32) {
                            if (j >= lim) p0 = 0.f;
                            if (j + 1 >= lim) p1 = 0.f;
                        }
                        rs = ptx::add2(rs, ptx::pk2(p0, p1));
                        pk[j >> 1] = pack_bf16(p0, p1);
                    }
                    uint8_t *slab = prow + (c >> 1) * GA_P_SLAB;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t unit = (uint32_t)((c & 1) * 4 + g) ^ sw;
                        *reinterpret_cast<uint4 *>(slab + unit * 16) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
                    }
                }
                {
                    float a, b2;
                    ptx::up2(rs, a, b2);
                    l += a + b2;
                }
                ptx::tc_fence_before();            // our reads of S_w / O_w are ordered before the MMAs that p_full releases
                ptx::fence_proxy_async();          // P is read by the tensor core (async proxy)
                ptx::mbar_arrive(ptx::smem_u32(&bars->p_full[w]));
            }
            ptx::mbar_wait(ptx::smem_u32(&bars->o_full[w]), oph);
            oph ^= 1;
            ptx::tc_fence_after();
            {
                uint32_t ov[16];
                ptx::tmem_ld_x16(o_addr, ov);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int d = 0; d < 16; ++d) o[d] += __uint_as_float(ov[d]);
            }
            ptx::tc_fence_before();
            float *dst = p.scratch + ((long long)s.item * p.max_parts + s.part) * (GA_PART_FLOATS * 256) + w * 128 + row;
            dst[0] = m;
            dst[256] = l;
#pragma unroll
            for (int d = 0; d < 16; ++d) dst[(2 + d) * 256] = o[d];
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, GA_TMEM_COLS);
}

// combine the partials of every query row and write softmax(q k^T) v as bf16 at (frame, token, head * 16 + d)
__global__ void __launch_bounds__(256) global_attn_merge_kernel(const GaParams p, int grid_attn, bf16 *__restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int item = blockIdx.x, r = threadIdx.x;
    const int qp = item % p.qpairs, bh = item / p.qpairs, h = bh % p.heads, b = bh / p.heads;
    const int tok = qp * 256 + r;
    if (tok >= p.S) return;
    const int first = ga_owner((long long)item * p.ktiles, p.units, grid_attn);
    const int last = ga_owner((long long)item * p.ktiles + p.ktiles - 1, p.units, grid_attn);
    const int n = last - first + 1;
    const float *base = p.scratch + (long long)item * p.max_parts * (GA_PART_FLOATS * 256) + r;
    const float L2E = 1.4426950408889634f;
    float M = -INFINITY;
    for (int i = 0; i < n; ++i) M = fmaxf(M, base[(long long)i * GA_PART_FLOATS * 256]);
    float L = 0.f, o[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) o[d] = 0.f;
    for (int i = 0; i < n; ++i) {
        const float *pp = base + (long long)i * GA_PART_FLOATS * 256;
        const float f = ex2f((pp[0] - M) * L2E);
        L = fmaf(pp[256], f, L);
#pragma unroll
        for (int d = 0; d < 16; ++d) o[d] = fmaf(pp[(2 + d) * 256], f, o[d]);
    }
    const float inv = 1.f / L;
    uint32_t w[8];
#pragma unroll
    for (int d = 0; d < 16; d += 2) w[d >> 1] = pack_bf16(o[d] * inv, o[d + 1] * inv);
    uint4 *op = reinterpret_cast<uint4 *>(out + ((long long)b * p.S + tok) * p.dim + h * 16);
    op[0] = make_uint4(w[0], w[1], w[2], w[3]);
    op[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// V^T: vt[((b * heads + h) * 16 + d) * S + s] = qkv[(b * S + s) * 3 dim + 2 dim + h * 16 + d]; 64 tokens per CTA
__global__ void __launch_bounds__(256) global_attn_vt_kernel(const bf16 *__restrict__ qkv, bf16 *__restrict__ vt, int S, int dim) {
    pdl_trigger();
    pdl_wait();
    __shared__ bf16 tile[64][192 + 2];
    const int b = blockIdx.y, s0 = blockIdx.x * 64;
    const int nvec = dim / 8;                                  // 16-byte pieces of a token's V row
    for (int e = threadIdx.x; e < 64 * nvec; e += 256) {
        const int t = e / nvec, c = e - t * nvec;
        if (s0 + t < S) {
            const uint4 v = *reinterpret_cast<const uint4 *>(qkv + ((long long)b * S + s0 + t) * 3 * dim + 2 * dim + c * 8);
            const bf16 *pv = reinterpret_cast<const bf16 *>(&v);
#pragma unroll
            for (int i = 0; i < 8; ++i) tile[t][c * 8 + i] = pv[i];
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < dim * 32; e += 256) {        // (row of V^T, pair of tokens)
        const int rowd = e >> 5, tp = (e & 31) * 2;
        if (s0 + tp < S) {
            bf16 *dst = vt + ((long long)b * dim + rowd) * S + s0 + tp;
            if (s0 + tp + 1 < S) *reinterpret_cast<__nv_bfloat162 *>(dst) = __halves2bfloat162(tile[tp][rowd], tile[tp + 1][rowd]);
            else dst[0] = tile[tp][rowd];
        }
    }
}

PerDeviceFlag g_ga_attr;
thread_local int g_ga_enable = 1;

}  // namespace

void tc_set_global_attn(int on) { g_ga_enable = on; }

static int ga_geometry(int B, int S, int heads, int grid, GaParams &p) {
    p.B = B; p.S = S; p.heads = heads; p.dim = heads * 16;
    p.qpairs = ceil_div(S, 256);
    p.ktiles = ceil_div(S, 128);
    p.units = (long long)B * heads * p.qpairs * p.ktiles;
    int mp = 1;
    const long long items = (long long)B * heads * p.qpairs;
    // parts of an item = CTAs its key range touches; bounded by how many range boundaries fit into one item
    const long long per_cta = p.units / grid;
    mp = per_cta > 0 ? (int)((p.ktiles + per_cta - 1) / per_cta) + 1 : p.ktiles;
    if (mp > p.ktiles) mp = p.ktiles;
    p.max_parts = mp;
    (void)items;
    return TU_OK;
}

static int ga_grid(int B, int S, int heads) {
    const long long units = (long long)B * heads * ceil_div(S, 256) * ceil_div(S, 128);
    const int sms = device_sm_count();
    return (int)(units < sms ? units : sms);
}

size_t tc_global_attention_scratch_bytes(int B, int S, int heads) {
    GaParams p;
    ga_geometry(B, S, heads, ga_grid(B, S, heads), p);
    return (size_t)B * heads * p.qpairs * p.max_parts * GA_PART_FLOATS * 256 * sizeof(float);
}

// qkv (B*S, 3*dim) bf16 rows [q | k | v], q pre-scaled; out (B*S, dim) bf16; vt: B*dim*S bf16 scratch; scratch: see above
int tc_global_attention(const bf16 *qkv, bf16 *out, bf16 *vt, float *scratch, size_t scratch_bytes, int B, int S, int heads,
                        cudaStream_t st) {
    TcEncodeFn enc = tc_encode_fn();
    const int dim = heads * 16;
    if (!g_ga_enable || !enc || (dim % 64) || dim > 192 || (S % 8) || S < 128 || (reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(vt) & 15) ||
        (reinterpret_cast<uintptr_t>(out) & 15) || !scratch || scratch_bytes < tc_global_attention_scratch_bytes(B, S, heads))
        return TU_TC_UNSUPPORTED;
    if (!g_ga_attr.is_set()) {
        cudaError_t e = cudaFuncSetAttribute(global_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GA_SMEM);
        if (e != cudaSuccess) return cuda_fail(e, "global_attn_tc smem attribute");
        g_ga_attr.set();
    }
    const int grid = ga_grid(B, S, heads);
    GaParams p;
    ga_geometry(B, S, heads, grid, p);
    p.scratch = scratch;
    CUtensorMap tq, tv;
    {
        cuuint64_t d3[3] = {(cuuint64_t)3 * dim, (cuuint64_t)S, (cuuint64_t)B};
        cuuint64_t s3[2] = {(cuuint64_t)3 * dim * 2, (cuuint64_t)S * 3 * dim * 2};
        cuuint32_t b3[3] = {64, 128, 1}, e3[3] = {1, 1, 1};
        CUresult r = enc(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void *)qkv, d3, s3, b3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t d2[2] = {(cuuint64_t)S, (cuuint64_t)B * dim}, s2[1] = {(cuuint64_t)S * 2};
        cuuint32_t b2[2] = {64, 16}, e2[2] = {1, 1};
        if (r == CUDA_SUCCESS)
            r = enc(&tv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)vt, d2, s2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("tu: cuTensorMapEncodeTiled(global attention) failed with code " + std::to_string((int)r));
            return TU_ERR_CUDA;
        }
    }
    launch_pdl(global_attn_vt_kernel, dim3(ceil_div(S, 64), B), dim3(256), 0, st, qkv, vt, S, dim);
    TU_CHECK_LAUNCH("global_attn_vt");
    launch_pdl(global_attn_tc_kernel, dim3(grid), dim3(GA_THREADS), (size_t)GA_SMEM, st, tq, tv, p);
    TU_CHECK_LAUNCH("global_attn_tc");
    launch_pdl(global_attn_merge_kernel, dim3(B * heads * p.qpairs), dim3(256), 0, st, p, grid, out);
    TU_CHECK_LAUNCH("global_attn_merge");
    return TU_OK;
}

}  // namespace tu
