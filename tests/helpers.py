"""Shared helpers of the parity tests."""
import numpy as np


def pair_sets(off, j, f, n):
    """Per-point sorted arrays of packed (j, f) codes -- order-insensitive list comparison."""
    out = []
    for i in range(n):
        a, b = int(off[i]), int(off[i + 1])
        out.append(np.sort(j[a:b].astype(np.int64) * 32 + f[a:b].astype(np.int64)))
    return out


def assert_close_scaled(got, want64, abs64, rtol, atol, what=""):
    """|got - want| <= atol + rtol * sum|terms|  (the oracle sums in another order; abs64 is the sum
    of the magnitudes of the terms, the natural scale of fp32 summation error)."""
    err = np.abs(got.astype(np.float64) - want64)
    bound = atol + rtol * abs64
    bad = err > bound
    if bad.any():
        k = np.unravel_index(np.argmax(err - bound), err.shape)
        raise AssertionError(f"{what}: {bad.sum()} elements out of tolerance; worst at {k}: "
                             f"got {got[k]!r} want {want64[k]!r} |err| {err[k]:.3e} bound {bound[k]:.3e}")
    return float((err / np.maximum(abs64, 1e-30)).max())
