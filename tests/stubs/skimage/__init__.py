"""TEST SCAFFOLDING: minimal stand-in for scikit-image (absent from this image) so that the reference's
inference.py can be executed unmodified.  Only what inference.py:23-26,129-141 touches."""
import numpy as np


def img_as_float(a):
    a = np.asarray(a)
    return a.astype(np.float64) / 255.0 if a.dtype == np.uint8 else a.astype(np.float64)
