"""CPU: bounds proof of the unrolled row-schedule bicubic kernel's shared-memory indexing (transformerupscaler_b200/csrc/resample.cu,
bicubic_add_clamp_r32_kernel + the tile plan in tu_bicubic_add_clamp).  compute-sanitizer is not available on the GPU pool, so the
kernel's index arithmetic is restated here in numpy (ATen's fp32 source coordinate, UpSample.h:259-312) and every shared-memory read
of every thread of every tile is checked against the staged tile's extent, for the BASELINE shapes and for ragged widths:
  * a column pair reads six consecutive x elements from the even column p0 = (iA - 1) & ~1 and five residual elements from jA - 1;
  * both columns' four taps must lie inside those six / five elements (the zero-padded weight vectors);
  * the rows a CTA walks (4 + the schedule's steps per tile) must fit the planned box."""
import numpy as np
import pytest

BS_W = 128


def src_floor(dst, n_in, n_out):
    scale = np.float32(n_in) / np.float32(n_out)
    src = (np.float64(scale) * (np.asarray(dst, dtype=np.float64) + 0.5) - 0.5).astype(np.float32)       # one rounding, like fmaf
    return np.minimum(np.floor(src).astype(np.int64), n_in - 1)


def sched(pat, big):
    """period, rows per CTA (RowSched<PAT>::TILE / ::BIG), x step?, residual step?"""
    if pat == 0:
        return 12, (60 if big else 36), (lambda q: q % 3 != 0), (lambda q: q % 3 == 1)
    tile = {2: 80, 3: 72, 4: 96, 6: 96}[pat] if big else (48 if pat == 2 else 8 * pat)
    return 8 * pat, tile, (lambda q: q % pat == pat // 2), (lambda q: q % (2 * pat) == pat)


CASES = [  # (H, W), (rH, rW), (oH, oW), schedule, element bytes of x
    ((720, 1280), (360, 640), (1080, 1920), 0, 2), ((720, 1280), (360, 640), (1080, 1920), 0, 1),
    ((720, 1280), (360, 640), (1440, 2560), 2, 2), ((720, 1280), (360, 640), (2160, 3840), 3, 2),
    ((720, 1280), (360, 640), (2880, 5120), 4, 1), ((720, 1280), (360, 640), (4320, 7680), 6, 2),
    ((1080, 1920), (540, 960), (1620, 2880), 0, 2), ((48, 64), (24, 40), (72, 250), 0, 2), ((24, 304), (12, 152), (36, 456), 0, 1),
    ((16, 64), (8, 40), (48, 200), 3, 2), ((48, 64), (24, 40), (96, 250), 2, 1), ((16, 32), (8, 24), (64, 150), 4, 2),
    ((720, 1280), (360, 320), (1080, 5120), 0, 2),      # wide up-scaling: s = 0.25 and 0.0625 (the +8 slack of the plan)
]


@pytest.mark.parametrize("big", [False, True])
@pytest.mark.parametrize("case", CASES)
def test_every_shared_memory_read_is_inside_the_tile(case, big):
    (H, W), (rH, rW), (oH, oW), pat, eb = case
    P, TILE, xstep, rstep = sched(pat, big)
    assert oH % P == 0 and oW % 2 == 0 and TILE % P == 0 and 2 * TILE <= 192
    # ---- the host's plan (tu_bicubic_add_clamp: plan(bh) with bh = TILE)
    xr = (TILE - 1) * H // oH + 6
    xc = ((BS_W - 1) * W // oW + 8 + (16 // eb - 1) + 15) & ~15
    rr = (TILE - 1) * rH // oH + 6
    rc = ((BS_W - 1) * rW // oW + 8 + 3 + 3) & ~3
    XA = 16 // eb
    # ---- rows: window start + 4 rows + one row per step of the tile
    nx = 4 + sum(xstep(q % P) for q in range(TILE))
    nr = 4 + sum(rstep(q % P) for q in range(TILE))
    assert nx <= xr and nr <= rr, (nx, xr, nr, rr)
    # the tile's first row is the window's first row of its first output row, on every tile
    oy0 = np.arange(0, oH, TILE)
    ex = (4 * oy0 - 1) // 6 if pat == 0 else (2 * oy0 + 1 - pat) // (2 * pat)
    er = (oy0 - 1) // 3 if pat == 0 else (2 * oy0 + 1 - 2 * pat) // (4 * pat)
    assert np.array_equal(src_floor(oy0, H, oH), ex) and np.array_equal(src_floor(oy0, rH, oH), er)
    # ---- columns: every pair of every column tile
    ox = np.arange(0, oW, 2)
    ox0 = (ox // BS_W) * BS_W
    iA, iB = src_floor(ox, W, oW), src_floor(ox + 1, W, oW)
    xc0 = (src_floor(ox0, W, oW) - 1) & ~(XA - 1)
    p0 = (iA - 1) & ~1
    offA, offB = iA - 1 - p0, iB - 1 - p0
    assert ((offA >= 0) & (offA <= 1) & (offB >= offA) & (offB <= 2)).all()           # four taps of both columns inside e0..e5
    assert ((p0 - xc0 >= 0) & (p0 - xc0 + 5 <= xc - 1)).all(), (int((p0 - xc0 + 5).max()), xc)
    assert (xc0 % 2 == 0).all()                                                        # parity of tile and source columns agree
    jA, jB = src_floor(ox, rW, oW), src_floor(ox + 1, rW, oW)
    rc0 = (src_floor(ox0, rW, oW) - 1) & ~3
    d = jB - jA
    assert ((d == 0) | (d == 1)).all()
    assert ((jA - 1 - rc0 >= 0) & (jA - 1 - rc0 + 4 <= rc - 1)).all(), (int((jA - 1 - rc0 + 4).max()), rc)
    # ---- shared memory: tile + 128 bytes of alignment slack within the 96 KB the kernels request
    assert ((3 * xr * xc * eb + 127) & ~127) + 3 * rr * rc * 4 + 128 <= 96 * 1024
