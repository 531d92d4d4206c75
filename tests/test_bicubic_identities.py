"""CPU: the arithmetic identities the row-schedule bicubic kernel's output stage relies on (transformerupscaler_b200/csrc/resample.cu,
clamp_store_pair / store_pair_unit / unit_to_u8), checked on the host with torch's own conversions.  They are what makes the kernel
bitwise equal to the pair kernel it replaces (tests/test_gpu_ops.py::test_bicubic_fixed_row_pattern_kernel) and to the reference's
`torch.clamp(out, 0, 1)` / `(out * 255).clamp(0, 255).to(uint8)` tails (WindowTransformer/model.py:304-305, app_overlay.py:382-386)."""
import numpy as np
import torch


def _samples():
    g = torch.Generator().manual_seed(7)
    v = torch.cat([torch.empty(400000).uniform_(-0.5, 1.5, generator=g), torch.empty(100000).uniform_(0.99, 1.01, generator=g),
                   torch.empty(100000).uniform_(-0.01, 0.01, generator=g),
                   torch.tensor([0.0, -0.0, 1.0, 1.0 + 2**-9, 1.0 + 2**-8, 1.0 - 2**-9, 1.0 - 2**-10, 2**-140, -2**-140, 0.5, 255.0 / 256,
                                 1.00390625, 0.99609375, 3.0e38, -3.0e38])])
    return v.float()


def test_bf16_relu_then_min_equals_clamp_then_round():
    """cvt.rn.relu.bf16x2.f32 + min.bf16x2 with 1.0 == round-to-bf16(clamp(v, 0, 1)): rounding is monotonic and 0, 1 are bf16 values."""
    v = _samples()
    want = v.clamp(0, 1).to(torch.bfloat16)
    relu_rounded = torch.where(v > 0, v, torch.zeros_like(v)).to(torch.bfloat16)          # relu: negatives (and -0) become +0
    got = torch.minimum(relu_rounded, torch.tensor(1.0, dtype=torch.bfloat16))
    nz = want.float() != 0
    assert torch.equal(got.view(torch.int16)[nz], want.view(torch.int16)[nz])              # bit patterns, not values
    assert torch.equal(got.float()[~nz], want.float()[~nz])                                # zeros: +0 here, torch.clamp keeps the sign of -0


def test_uint8_store_without_the_noop_clamp():
    """For v already in [0, 1]: floor(fp32(v * 255)) — what the low mantissa byte of (v * 255 +rd 2^23) holds — equals the reference's
    (v * 255).clamp(0, 255).to(uint8); the clamp to [0, 255] is a no-op because 1.0 * 255 is exact and the product is monotonic."""
    v = _samples().clamp(0, 1)
    want = (v * 255).clamp(0, 255).to(torch.uint8)
    t = (v * 255).numpy()                                                                  # fp32 product, round to nearest
    assert t.min() >= 0.0 and t.max() <= 255.0
    summed = (t.astype(np.float64) + 8388608.0)                                            # exact in float64
    rd = np.floor(summed)                                                                  # round-down to the fp32 grid at 2^23: spacing 1
    low_byte = (rd.astype(np.int64) & 0xFF).astype(np.uint8)
    assert np.array_equal(low_byte, want.numpy())


def test_zero_weight_taps_leave_the_sum_unchanged():
    """fma(e, 0, t) == t for finite e: the six-tap (x) / five-tap (residual) zero-padded filters give the four-tap sums bit for bit;
    a leading zero-weight tap only turns the first product into a zero of either sign, which the next fma absorbs."""
    rs = np.random.RandomState(3)
    e = rs.uniform(-300, 300, 100000).astype(np.float32)
    t = rs.uniform(-4, 4, 100000).astype(np.float32)
    w = rs.uniform(-1, 1, 100000).astype(np.float32)
    assert np.array_equal((e * np.float32(0) + t).view(np.int32), t.view(np.int32))        # e * 0 is exact: no fused rounding involved
    lead = e * np.float32(0.0)                                                             # +0 or -0
    prod = (e.astype(np.float64) * w.astype(np.float64))
    fused = (prod + lead.astype(np.float64)).astype(np.float32)                            # fma(e, w, +-0): one rounding of the exact product
    direct = prod.astype(np.float32)                                                       # mul.rn(e, w)
    nz = direct != 0
    assert np.array_equal(fused[nz].view(np.int32), direct[nz].view(np.int32))
