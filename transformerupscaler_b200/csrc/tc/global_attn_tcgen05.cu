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
// CTA = 18 warps:
//   warps 0-7   softmax group 0 (query tile 0), warps 8-15 softmax group 1 (query tile 1).  TWO threads share a query row (its TMEM
//               lane): warp w of a group owns lane quadrant w mod 4 and key half w / 4 of every tile (64 scores, = one 64-key slab of
//               the P operand).  The 64 scores are read from TMEM ONCE (the S buffer goes straight back to the MMA warp, which issues
//               the next tile's Q K^T while this one is being exponentiated); the two halves exchange their row maxima through
//               shared memory (one 256-thread named barrier per tile), alpha = 2^((m_old - m_new) c),
//               P = 2^(S c - m_new c) rounded to bf16 into 128-byte-swizzled K-major shared memory; each half keeps 8 of the 16
//               running outputs, acc = (acc + O_prev) alpha, in spare TMEM columns (updated half way through the exponentials,
//               when half of the score registers are free).
//   warp 16     TMA producer: the two Q tiles of a segment (double buffered), then per key tile the K slab and the V^T slabs
//               (3-stage ring)
//   warp 17     MMA issuer: S_w = Q_w K^T (one 128x128x16 UMMA per tile and query tile), O_w = P_w V (eight 128x16x16 UMMAs)
// The two groups run half a tile apart: while one waits for its S tile or its PV product the other keeps the MUFU busy
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

constexpr int GA_THREADS = 576;
constexpr int GA_KV_STAGES = 3;
constexpr int GA_Q_BYTES = 128 * 128;            // one query tile: 128 rows x 64 bf16 (four heads)
constexpr int GA_K_BYTES = 128 * 128;            // one key tile, same shape
constexpr int GA_VT_SLAB = 16 * 128;             // 16 rows (d) x 64 keys
constexpr int GA_KV_STAGE = GA_K_BYTES + 2 * GA_VT_SLAB;
constexpr int GA_P_SLAB = 128 * 128;             // 128 rows x 64 keys bf16
constexpr int GA_P_BYTES = 2 * GA_P_SLAB;        // per warpgroup
constexpr int GA_OFF_Q = 0;
constexpr int GA_OFF_KV = GA_OFF_Q + 4 * GA_Q_BYTES;       // two segments' worth of Q tiles
constexpr int GA_OFF_P = GA_OFF_KV + GA_KV_STAGES * GA_KV_STAGE;
constexpr int GA_OFF_X = GA_OFF_P + 2 * GA_P_BYTES;        // row-max / row-sum exchange: [parity 2][group 2][half 2][128] floats
constexpr int GA_OFF_BAR = GA_OFF_X + 2 * 2 * 2 * 128 * 4;
constexpr int GA_SMEM = GA_OFF_BAR + 256 + 1024;
constexpr int GA_TMEM_COLS = 512;                // S0 [0,128) S1 [128,256); P V products [256,288): 16 per group; running o [288,320)
constexpr int GA_PART_FLOATS = 18;               // m, l, o[16]
static_assert(GA_OFF_KV % 1024 == 0 && GA_OFF_P % 1024 == 0 && GA_KV_STAGE % 1024 == 0 && GA_SMEM <= 232448, "shared memory layout");

struct GaParams {
    int B, S, heads, dim;
    int qpairs;              // ceil(S / 256)
    int ktiles;              // ceil(S / 128)
    long long units;         // B * heads * qpairs * ktiles
    int max_parts;
    float *scratch;          // [item][part][18][256]
    unsigned long long *trace;   // debug (tu_debug_trace): clock64 at the phase boundaries of CTA 0's softmax warps 0 and 8, 8 words per tile
};

struct GaBars {
    uint64_t q_full[2], q_free[2];
    uint64_t kv_full[GA_KV_STAGES], kv_empty[GA_KV_STAGES];
    uint64_t s_full[2], s_free[2], p_full[2], o_full[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void named_barrier(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_barrier_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// The two softmax groups take turns on the MUFU pipe: group 0 exponentiates tile t, then group 1 its tile t, then group 0 tile t + 1 ...
// (left alone they fall into lock step -- both in the exponentials at half rate, then both out of them: MUFU 52 % busy, profiles/r2_ga1).
// Barrier GA_BAR_GO0 is "group 0 may go" (group 1 arrives on it when its exponentials are done), GA_BAR_GO1 the reverse.
constexpr int GA_BAR_GO0 = 3, GA_BAR_GO1 = 4;
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

// 18 warps are allocated as 20 (granularity 4): 96 registers per thread.  A softmax thread holds 64 scores of its row; its share of the
// running output therefore lives in spare TMEM columns between tiles.
__global__ void __launch_bounds__(GA_THREADS, 1)
global_attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_vt, const GaParams p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *sm = smem_raw + (smem0 - ptx::smem_u32(smem_raw));
    GaBars *bars = reinterpret_cast<GaBars *>(sm + GA_OFF_BAR);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(ptx::smem_u32(&bars->q_full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->q_free[i]), 1);
        }
        for (int i = 0; i < GA_KV_STAGES; ++i) {
            ptx::mbar_init(ptx::smem_u32(&bars->kv_full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->kv_empty[i]), 1);
        }
        for (int w = 0; w < 2; ++w) {
            ptx::mbar_init(ptx::smem_u32(&bars->s_full[w]), 1);
            ptx::mbar_init(ptx::smem_u32(&bars->s_free[w]), 8);      // one arrival per warp of the group
            ptx::mbar_init(ptx::smem_u32(&bars->p_full[w]), 8);
            ptx::mbar_init(ptx::smem_u32(&bars->o_full[w]), 1);
        }
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&tmap_qkv);
        ptx::prefetch_tmap(&tmap_vt);
    }
    // P starts as zeros (never read uninitialised by the tensor core)
    for (int i = threadIdx.x; i < 2 * GA_P_BYTES / 16; i += GA_THREADS) reinterpret_cast<uint4 *>(sm + GA_OFF_P)[i] = make_uint4(0, 0, 0, 0);
    ptx::fence_proxy_async();
    if (warp == 17) {
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

    if (warp == 16) {
        if (lane == 0) {
            // ================================ TMA producer ================================
            int stage = 0, seg = 0;
            uint32_t phase = 0;
            long long u = u_begin;
            Segment s;
            while (next_segment(p, u, u_end, s)) {
                const int qb = seg & 1;
                ptx::mbar_wait(ptx::smem_u32(&bars->q_free[qb]), ((seg >> 1) & 1) ^ 1);
                const uint32_t qf = ptx::smem_u32(&bars->q_full[qb]);
                ptx::mbar_expect_tx(qf, (s.act1 ? 2 : 1) * GA_Q_BYTES);
                const int slab = s.h >> 2;
                const uint32_t qdst = smem0 + GA_OFF_Q + qb * 2 * GA_Q_BYTES;
                tma_load_3d(qdst, &tmap_qkv, qf, slab * 64, s.qp * 256, s.b);
                if (s.act1) tma_load_3d(qdst + GA_Q_BYTES, &tmap_qkv, qf, slab * 64, s.qp * 256 + 128, s.b);
                ++seg;
                for (int kt = s.kt0; kt < s.kt1; ++kt) {
                    ptx::mbar_wait(ptx::smem_u32(&bars->kv_empty[stage]), phase ^ 1);
                    const uint32_t fb = ptx::smem_u32(&bars->kv_full[stage]);
                    const uint32_t dst = smem0 + GA_OFF_KV + stage * GA_KV_STAGE;
                    ptx::mbar_expect_tx(fb, GA_KV_STAGE);
                    tma_load_3d(dst, &tmap_qkv, fb, p.dim + slab * 64, kt * 128, s.b);
                    const int vrow = (s.b * p.heads + s.h) * 16;
                    ptx::tma_load_2d(dst + GA_K_BYTES, &tmap_vt, fb, kt * 128, vrow);
                    ptx::tma_load_2d(dst + GA_K_BYTES + GA_VT_SLAB, &tmap_vt, fb, kt * 128 + 64, vrow);
                    if (++stage == GA_KV_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 17) {
        // ================================ MMA issuer (whole warp converged, elected lane issues) ================================
        const uint32_t leader = ptx::elect_one();
        const uint32_t idesc_qk = ptx::make_idesc_bf16(128, 128), idesc_pv = ptx::make_idesc_bf16(128, 16);
        int stage = 0, seg = 0;
        uint32_t phase = 0, pph[2] = {0u, 0u}, fph[2] = {0u, 0u};
        long long u = u_begin;
        Segment s;
        while (next_segment(p, u, u_end, s)) {
            const int T = s.kt1 - s.kt0, nw = s.act1 ? 2 : 1, qb = seg & 1;
            const uint32_t hoff = (uint32_t)(s.h & 3) * 2u;                    // 32 bytes per head inside the four-head slab
            const uint32_t q_lo = ptx::sdesc_lo(smem0 + GA_OFF_Q + qb * 2 * GA_Q_BYTES) + hoff;
            ptx::mbar_wait(ptx::smem_u32(&bars->q_full[qb]), (seg >> 1) & 1);
            ptx::mbar_wait(ptx::smem_u32(&bars->kv_full[stage]), phase);
            ptx::tc_fence_after();
            for (int w = 0; w < nw; ++w) {
                ptx::umma_bf16_lo<0>(tmem_base + w * 128, q_lo + w * (GA_Q_BYTES >> 4), ptx::sdesc_lo(smem0 + GA_OFF_KV + stage * GA_KV_STAGE) + hoff,
                                     idesc_qk, leader);
                ptx::umma_commit_pred(ptx::smem_u32(&bars->s_full[w]), leader);
            }
            for (int t = 0; t < T; ++t) {
                int nstage = stage + 1;
                uint32_t nphase = phase;
                if (nstage == GA_KV_STAGES) { nstage = 0; nphase ^= 1; }
                // next tile's scores as soon as the warpgroup has S(t) in registers
                if (t + 1 < T) ptx::mbar_wait(ptx::smem_u32(&bars->kv_full[nstage]), nphase);
                for (int w = 0; w < nw; ++w) {
                    ptx::mbar_wait(ptx::smem_u32(&bars->s_free[w]), fph[w]);
                    fph[w] ^= 1;
                    if (t + 1 < T) {
                        ptx::tc_fence_after();
                        ptx::umma_bf16_lo<0>(tmem_base + w * 128, q_lo + w * (GA_Q_BYTES >> 4),
                                             ptx::sdesc_lo(smem0 + GA_OFF_KV + nstage * GA_KV_STAGE) + hoff, idesc_qk, leader);
                        ptx::umma_commit_pred(ptx::smem_u32(&bars->s_full[w]), leader);
                    }
                }
                if (t + 1 == T) ptx::umma_commit_pred(ptx::smem_u32(&bars->q_free[qb]), leader);     // the segment's last Q K^T has been issued
                const uint32_t vt_lo = ptx::sdesc_lo(smem0 + GA_OFF_KV + stage * GA_KV_STAGE + GA_K_BYTES);
                for (int w = 0; w < nw; ++w) {
                    ptx::mbar_wait(ptx::smem_u32(&bars->p_full[w]), pph[w]);   // P_w(t) written, O_w(t-1) read
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
                }
                ptx::umma_commit_pred(ptx::smem_u32(&bars->kv_empty[stage]), leader);
                stage = nstage;
                phase = nphase;
            }
            ++seg;
        }
    } else {
        // ================================ softmax groups ================================
        const int w = warp >> 3, half = (warp >> 2) & 1, q = warp & 3, row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t s_addr = tmem_base + lane_off + w * 128 + half * 64;
        const uint32_t o_addr = tmem_base + lane_off + 256 + w * 16 + half * 8;          // this thread's 8 of the 16 outputs of P V
        const uint32_t acc_addr = tmem_base + lane_off + 288 + w * 16 + half * 8;        // ... and of the running un-normalised output
        uint8_t *prow = sm + GA_OFF_P + w * GA_P_BYTES + half * GA_P_SLAB + row * 128;
        float *xch = reinterpret_cast<float *>(sm + GA_OFF_X);                            // [parity][group][half][row]
        const uint32_t sw = (uint32_t)(row & 7);
        const float L2E = 1.4426950408889634f;
        uint32_t sph = 0, oph = 0;
        int xpar = 0;
        // phase timeline for tools/probes/attn_trace.py: compiled in only with -DTU_GA_TRACE (the marks cost registers in a loop that has none to spare)
#ifdef TU_GA_TRACE
        int trc = 0;
        const bool tracing = p.trace && blockIdx.x == 0 && lane == 0 && q == 0 && half == 0;
#define GA_MARK(k) do { if (tracing && trc < 96) p.trace[(w * 96 + trc) * 8 + (k)] = (unsigned long long)clock64(); } while (0)
#define GA_NEXT_TILE() ++trc
#else
#define GA_MARK(k) do { } while (0)
#define GA_NEXT_TILE() do { } while (0)
#endif
        long long u = u_begin;
        Segment s;
        if (w == 1) named_barrier_arrive(GA_BAR_GO0, 512);      // group 0 starts; the arrival group 1 leaves behind after its last tile of a
                                                                // segment is the "go" for group 0's first tile of the next one
        while (next_segment(p, u, u_end, s)) {
            if (w == 1 && !s.act1) continue;
            const bool pingpong = s.act1 != 0;                  // with the second query tile empty, group 0 runs alone
            const int T = s.kt1 - s.kt0;
            float m = -INFINITY, l = 0.f;
            for (int t = 0; t < T; ++t) {
                const int nvalid = min(128, p.S - (s.kt0 + t) * 128) - half * 64;      // valid columns of this half (may be <= 0)
                GA_MARK(0);
                ptx::mbar_wait(ptx::smem_u32(&bars->s_full[w]), sph);
                sph ^= 1;
                ptx::tc_fence_after();
                GA_MARK(1);
                // ---- this thread's 64 scores into registers, then the S buffer goes back to the MMA warp
                uint32_t v[64];
                ptx::tmem_ld_x32(s_addr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                ptx::tmem_ld_x32(s_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->s_free[w]));
                GA_MARK(2);
                if (nvalid < 64) {                                           // keys past S: score -inf -> probability 0
#pragma unroll
                    for (int j = 0; j < 64; ++j)
                        if (j >= nvalid) v[j] = 0xff800000u;
                }
                // ---- row max: four independent chains over the own half, then the other half's through shared memory
                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
                for (int j = 0; j < 64; j += 8) {
                    mx0 = fmaxf(mx0, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
                    mx1 = fmaxf(mx1, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
                    mx2 = fmaxf(mx2, fmaxf(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])));
                    mx3 = fmaxf(mx3, fmaxf(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7])));
                }
                const float mloc = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                float *xp = xch + ((xpar * 2 + w) * 2) * 128 + row;
                xp[half * 128] = mloc;
                named_barrier(1 + w, 256);
                const float m_new = fmaxf(m, fmaxf(mloc, xp[(half ^ 1) * 128]));
                xpar ^= 1;
                GA_MARK(3);
                const float alpha = ex2f((m - m_new) * L2E);                   // 2^(-inf) = 0 on a segment's first tile
                if (t > 0) {
                    ptx::mbar_wait(ptx::smem_u32(&bars->o_full[w]), oph);       // P_w(t-1) V(t-1) is complete: the P buffer is free
                    oph ^= 1;
                    ptx::tc_fence_after();
                }
                GA_MARK(4);
                m = m_new;
                // ---- my group's turn on the MUFU: P = 2^(S c - m c) -> bf16, one 64-key slab of 128-byte swizzled K-major rows
                const ptx::f32x2 c2 = ptx::pk2(L2E, L2E), nb2 = ptx::pk2(-m_new * L2E, -m_new * L2E);
                ptx::f32x2 rs0 = ptx::pk2(0.f, 0.f), rs1 = rs0;
                if (pingpong) named_barrier(w == 0 ? GA_BAR_GO0 : GA_BAR_GO1, 512);
                GA_MARK(5);
#pragma unroll
                for (int g = 0; g < 8; ++g) {                                // 8 keys = one 16-byte unit of the operand row
                    if (g == 4 && t > 0) {
                        // half of the scores have been consumed: registers are free for the running output
                        // acc = (acc + O(t-1)) alpha, kept in spare TMEM columns between tiles
                        uint32_t ov[8], oa[8];
                        tmem_ld_x8(o_addr, ov);
                        if (t > 1) tmem_ld_x8(acc_addr, oa);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int d = 0; d < 8; ++d)
                            oa[d] = __float_as_uint(((t > 1 ? __uint_as_float(oa[d]) : 0.f) + __uint_as_float(ov[d])) * alpha);
                        tmem_st_x8(acc_addr, oa);
                        ptx::tmem_st_wait();
                    }
                    float e[8];
                    ptx::up2(ptx::fma2(ptx::pk2u(v[8 * g + 0], v[8 * g + 1]), c2, nb2), e[0], e[1]);
                    ptx::up2(ptx::fma2(ptx::pk2u(v[8 * g + 2], v[8 * g + 3]), c2, nb2), e[2], e[3]);
                    ptx::up2(ptx::fma2(ptx::pk2u(v[8 * g + 4], v[8 * g + 5]), c2, nb2), e[4], e[5]);
                    ptx::up2(ptx::fma2(ptx::pk2u(v[8 * g + 6], v[8 * g + 7]), c2, nb2), e[6], e[7]);
#pragma unroll
                    for (int i = 0; i < 8; ++i) e[i] = ex2f(e[i]);
                    rs0 = ptx::add2(rs0, ptx::add2(ptx::pk2(e[0], e[1]), ptx::pk2(e[2], e[3])));
                    rs1 = ptx::add2(rs1, ptx::add2(ptx::pk2(e[4], e[5]), ptx::pk2(e[6], e[7])));
                    const uint32_t unit = (uint32_t)g ^ sw;
                    *reinterpret_cast<uint4 *>(prow + unit * 16) =
                        make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
                }
                GA_MARK(6);
                if (pingpong) named_barrier_arrive(w == 0 ? GA_BAR_GO1 : GA_BAR_GO0, 512);
                {
                    float a0, a1;
                    ptx::up2(ptx::add2(rs0, rs1), a0, a1);
                    l = fmaf(l, alpha, a0 + a1);
                }
                ptx::tc_fence_before();            // our read of O_w is ordered before the PV MMAs that p_full releases
                ptx::fence_proxy_async();          // P is read by the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->p_full[w]));
                GA_MARK(7);
                GA_NEXT_TILE();
            }
            ptx::mbar_wait(ptx::smem_u32(&bars->o_full[w]), oph);
            oph ^= 1;
            ptx::tc_fence_after();
            float o[8];
            {
                uint32_t ov[8], oa[8];
                tmem_ld_x8(o_addr, ov);
                if (T > 1) tmem_ld_x8(acc_addr, oa);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int d = 0; d < 8; ++d) o[d] = (T > 1 ? __uint_as_float(oa[d]) : 0.f) + __uint_as_float(ov[d]);
            }
            ptx::tc_fence_before();
            // the row sum of the two halves (same exchange buffer and parity discipline as the row maxima)
            float *xp = xch + ((xpar * 2 + w) * 2) * 128 + row;
            xp[half * 128] = l;
            named_barrier(1 + w, 256);
            const float lsum = l + xp[(half ^ 1) * 128];
            xpar ^= 1;
            float *dst = p.scratch + ((long long)s.item * p.max_parts + s.part) * (GA_PART_FLOATS * 256) + w * 128 + row;
            if (half == 0) {
                dst[0] = m;
                dst[256] = lsum;
            }
#pragma unroll
            for (int d = 0; d < 8; ++d) dst[(2 + half * 8 + d) * 256] = o[d];
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 17) ptx::tmem_dealloc(tmem_base, GA_TMEM_COLS);
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
// Measured on B200 (tools/probes/attn_probe.py, profiles/r2_attn_*.log): 109 us per layer for 2 frames and 756 us for 16 against 103 / 599 us of the
// mma.sync kernel it was meant to replace -- at head_dim 16 the tensor pipe is 5 % busy either way and the kernel is bound by the
// exponentials and by moving S (TMEM -> registers) and P (registers -> shared memory) around them, which mma.sync keeps in registers.
// The forward therefore uses this kernel only when asked to: tu_debug_set("global_attn_tc", 1).
thread_local int g_ga_enable = 0;

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
    p.trace = g_trace_buf;
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
