import numpy as np


def peak_signal_noise_ratio(a, b, data_range=1):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return float("inf") if mse == 0 else 10 * np.log10(data_range ** 2 / mse)


def structural_similarity(a, b, data_range=1, channel_axis=-1):
    """global (single-window) SSIM: enough for a smoke run of the script, not the windowed scikit-image value"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    ma, mb, va, vb = a.mean(), b.mean(), a.var(), b.var()
    cov = ((a - ma) * (b - mb)).mean()
    return float((2 * ma * mb + c1) * (2 * cov + c2) / ((ma ** 2 + mb ** 2 + c1) * (va + vb + c2)))
