// Work distribution of the fused window-stack kernels (window_stack_tcgen05.cu, window_stack192_tcgen05.cu).
// The enumeration itself (seg_of) is plain C++ so that tests/test_host_logic.py can compile and check it on the host.
#pragma once

#ifndef __CUDACC__
#define TU_HD inline
#else
#define TU_HD __host__ __device__ __forceinline__
#endif

namespace tu {

struct Seg { int tile, lo, hi; };   // blocks [lo, hi) of a 128-token tile

// k-th work segment of CTA `cta` of `grid` in processing order.
// Whole-tile mode (split == false): tiles cta, cta + grid, ... through all blocks.
// Split mode: the n_tiles * n_blocks (tile, block) units, tile-major, are dealt out in contiguous shares of units_per_cta
// (>= n_blocks); the CTA owns units [start, end).  It first runs the leading blocks of the tile its share ends in (the next CTA
// continues that tile and waits for it), then its whole tiles, and last the remaining blocks of the tile its share begins in
// (started by the previous CTA as ITS first segment, so the wait never depends on anything the waiting CTA has yet to do).
TU_HD bool seg_of(int cta, int grid, int n_tiles, int nb, bool split, int units_per_cta, int k, Seg &s) {
    if (!split) {
        const int t = cta + k * grid;
        s.tile = t; s.lo = 0; s.hi = nb;
        return t < n_tiles;
    }
    const int total = n_tiles * nb;
    const int start = cta * units_per_cta < total ? cta * units_per_cta : total;
    const int end = start + units_per_cta < total ? start + units_per_cta : total;
    if (start >= end) return false;
    const int o = start % nb, tA = start / nb, e = end % nb, tB = end / nb;
    const int first_full = tA + (o ? 1 : 0), nfull = tB > first_full ? tB - first_full : 0;
    if (e) {
        if (k == 0) { s.tile = tB; s.lo = 0; s.hi = e; return true; }
        --k;
    }
    if (k < nfull) { s.tile = first_full + k; s.lo = 0; s.hi = nb; return true; }
    k -= nfull;
    if (o && k == 0) { s.tile = tA; s.lo = o; s.hi = nb; return true; }
    return false;
}

#ifdef __CUDACC__
// Spin until flag != 0 (set by another CTA with __threadfence + atomicExch); a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void wait_flag_acquire(const int *flag) {
    const long long t0 = clock64();
    int v;
    do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (!v) {
            __nanosleep(100);
            if (clock64() - t0 > 4000000000LL) __trap();      // ~2 s
        }
    } while (!v);
}

__device__ __forceinline__ bool get_seg(int n_tiles, int nb, bool split, int units_per_cta, int k, Seg &s) {
    return seg_of((int)blockIdx.x, (int)gridDim.x, n_tiles, nb, split, units_per_cta, k, s);
}
#endif

}  // namespace tu
