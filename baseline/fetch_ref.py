#!/usr/bin/env python
"""Copy the UNMODIFIED reference files the tests / bench need into baseline/_ref (build container only: needs /root/reference).

baseline/_ref is git-ignored (the reference's sources never enter the history) but travels to the GPU box with `gpurun`:
  models/                         the reference's model classes  (bench.py --impl reference, live-oracle tests)
  speed_test.py, inference.py     the reference's entry points   (tests/test_zz_gpu_scripts.py runs them against the drop-in models/)
  data_handling/, tools/utils.py  what those two scripts import
`__graft_entry__.build()` calls this when /root/reference exists.
"""
import filecmp
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["speed_test.py", "inference.py", "data_handling/__init__.py", "data_handling/data_class.py", "tools/utils.py"]
MODEL_DIRS = ["WindowTransformer", "FastTransformer", "ResidualTransformer", "BicubicInterpolation"]


def fetch(verbose: bool = False) -> bool:
    if not os.path.isdir(REF):
        return False
    for m in MODEL_DIRS:
        src = os.path.join(REF, "models", m)
        for f in os.listdir(src):
            if f.endswith(".py"):
                FILES.append(os.path.join("models", m, f))
    for rel in sorted(set(FILES)):
        s, d = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not os.path.exists(d) or not filecmp.cmp(s, d, shallow=False):
            shutil.copyfile(s, d)
            if verbose:
                print("copied", rel)
    return True


if __name__ == "__main__":
    print("fetched" if fetch(True) else "no /root/reference here")
