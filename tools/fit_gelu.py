#!/usr/bin/env python
"""Fit of the GELU used by the bf16 tensor-core kernels (window_stack_tcgen05.cu, gemm_tcgen05.cu):

    GELU(x) = x * Phi(x),  Phi(x) ~= 0.5 * (1 + tanh(x * (c0 + c1 x^2 + c2 x^4))),  x^2 clamped to 64

The odd polynomial approximates atanh(erf(x / sqrt 2)); the coefficients minimise the maximum absolute error against the
exact erf form (nn.GELU default, the reference's MLP activation: WindowTransformer/model.py:146).  Prints the
coefficients and the error; the exact-erf form stays in the fp32 path (transformer_simt.cu)."""
import numpy as np
from scipy.optimize import minimize
from scipy.special import erf

x = np.linspace(-8, 8, 400001)
exact = 0.5 * x * (1 + erf(x / np.sqrt(2)))


def gelu(c, x):
    x2 = np.minimum(x * x, 64.0)
    return 0.5 * x * (1 + np.tanh(x * (c[0] + x2 * (c[1] + x2 * c[2]))))


r = minimize(lambda c: np.abs(gelu(c, x) - exact).max(), [0.7978845608, 0.0356774, 0.0], method="Nelder-Mead",
             options=dict(xatol=1e-10, fatol=1e-12, maxiter=20000))
print("coefficients", r.x.tolist())
print("max |error| on [-8, 8]", r.fun)
xx = np.linspace(-30, 30, 200001)
print("max |error| on [-30, 30]", np.abs(gelu(r.x, xx) - 0.5 * xx * (1 + erf(xx / np.sqrt(2)))).max())
