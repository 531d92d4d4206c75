"""CPU: the C-ABI library loads and exports every symbol include/tu_b200.h declares; ctypes structs match the C layout.
No compute call is made (no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tu_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from transformerupscaler_b200 import _lib
    return _lib.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tu_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from transformerupscaler_b200 import _lib
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in tu_b200.h but not exported by libtu_b200.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) <= set(names), "ctypes binds a symbol the header does not declare"


def test_struct_layout_matches_c(lib, tmp_path):
    from transformerupscaler_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu\\n",'
                   'sizeof(TuBlockWeights),sizeof(TuUpsamplerStage),sizeof(TuModelWeights),offsetof(TuModelWeights,blocks),'
                   'offsetof(TuModelWeights,up1),offsetof(TuModelWeights,finconv_b));return 0;}\n' % HEADER)
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", str(src), "-o", str(exe)])
    got = [int(t) for t in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_lib.TuBlockWeights), C.sizeof(_lib.TuUpsamplerStage), C.sizeof(_lib.TuModelWeights),
            _lib.TuModelWeights.blocks.offset, _lib.TuModelWeights.up1.offset, _lib.TuModelWeights.finconv_b.offset]
    assert got == want


def test_host_side_queries_work_without_gpu(lib):
    from transformerupscaler_b200 import _lib
    assert lib.tu_version() >= 100
    # workspace sizing is pure host arithmetic
    n = lib.tu_forward_workspace_bytes(0, 8, 720, 1280, 1080, 1920, 0, _lib.TU_BF16)
    assert 2 * 2**30 < n < 4 * 2**30
    assert lib.tu_forward_workspace_bytes(1, 1, 64, 64, 320, 320, 5, _lib.TU_BF16) == 0
    assert "was not built" in _lib.last_error()
    assert lib.tu_forward_workspace_bytes(2, 1, 64, 64, 96, 96, 0, _lib.TU_F32) == 0
    assert "must match" in _lib.last_error()


def test_bicubic_row_schedule_matches_aten_coordinates(lib):
    """Host logic of the unrolled bicubic kernel: the schedule the library picks must reproduce, on EVERY output row, the source row of
    ATen's fp32 coordinate arithmetic (UpSample.h: scale = (float)in / out, src = scale * (dst + 0.5) - 0.5 as one fma, floor) —
    restated here in numpy float32 with the fma emulated in float64 (exact for these magnitudes)."""
    import numpy as np

    def aten_rows(n_in, n_out):
        scale = np.float32(n_in) / np.float32(n_out)
        dst = np.arange(n_out, dtype=np.float64) + 0.5
        src = (np.float64(scale) * dst - 0.5).astype(np.float32)         # one rounding, like fmaf
        return np.minimum(np.floor(src).astype(np.int64), n_in - 1)

    def sched_rows(pat, n_out):
        oy = np.arange(n_out, dtype=np.int64)
        if pat == 0:
            return np.floor_divide(4 * oy - 1, 6), np.floor_divide(oy - 1, 3)
        return np.floor_divide(2 * oy + 1 - pat, 2 * pat), np.floor_divide(2 * oy + 1 - 2 * pat, 4 * pat)

    cases = {(720, 360, 1080): 0, (720, 360, 1440): 2, (720, 360, 2160): 3, (720, 360, 2880): 4, (720, 360, 4320): 6,
             (1080, 540, 1620): 0, (72, 36, 108): 0, (24, 12, 72): 3}
    for (H, rH, oH), want in cases.items():
        got = lib.tu_bicubic_row_schedule(H, rH, oH)
        ex, er = sched_rows(want, oH)
        exact = np.array_equal(aten_rows(H, oH), ex) and np.array_equal(aten_rows(rH, oH), er)
        assert got == (want if exact else -1), (H, rH, oH, got, exact)
        assert exact, (H, rH, oH)            # the BASELINE shapes do follow their schedule
    # heights without a schedule, and a period that does not divide outH
    for H, rH, oH in [(720, 360, 1000), (720, 240, 1080), (720, 360, 3600), (20, 10, 30), (0, 0, 0)]:
        assert lib.tu_bicubic_row_schedule(H, rH, oH) == -1


def test_missing_library_fails_loudly(monkeypatch):
    from transformerupscaler_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libtu_b200.so")
    with pytest.raises(ImportError, match="no CPU"):
        _lib.load()
