"""GPU tests of the tensor-core plumbing (tcgen05 / TMEM / 128B-swizzled operand panels)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(128, 64), (64, 32), (256, 128), (16, 32), (48, 96)])
def test_tcgen05_3xtf32_matches_float64(N, K):
    from pointwise_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(N * 1000 + K)
    A = rng.uniform(-1, 1, (128, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (N, K)).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    errs = {}
    for split in (1, 0):
        d = torch.full((128, N), float("nan"), device="cuda")
        _lib.check(L.conv3p_selftest_tc(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K, split, None))
        torch.cuda.synchronize()
        got = d.cpu().numpy().astype(np.float64)
        assert np.isfinite(got).all()
        errs[split] = float((np.abs(got - want) / scale).max())
    assert errs[1] < 4e-6, errs      # 3xTF32: ~2^-20 per product, i.e. fp32-class
    assert errs[0] < 2e-3, errs      # plain TF32: 2^-10 per operand
    assert errs[1] < errs[0] / 50, errs


@pytest.mark.parametrize("N,K", [(64, 64), (32, 8), (128, 32), (64, 16)])
def test_tcgen05_mn_major_operands(N, K):
    """Contraction index outermost in memory (point-major panels): D = A^T B."""
    from pointwise_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(N * 1000 + K + 7)
    A = rng.uniform(-1, 1, (K, 128)).astype(np.float32)
    B = rng.uniform(-1, 1, (K, N)).astype(np.float32)
    want = A.astype(np.float64).T @ B.astype(np.float64)
    scale = np.abs(A).astype(np.float64).T @ np.abs(B).astype(np.float64)
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    d = torch.full((128, N), float("nan"), device="cuda")
    _lib.check(L.conv3p_selftest_tc_mn(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K, 1, None))
    torch.cuda.synchronize()
    got = d.cpu().numpy().astype(np.float64)
    assert np.isfinite(got).all()
    assert float((np.abs(got - want) / scale).max()) < 4e-6


@pytest.mark.parametrize("N,K", [(128, 64), (64, 32), (256, 128), (16, 32), (48, 96)])
def test_tcgen05_tf32_plus_bf16_corrections_matches_float64(N, K):
    """The production split: one TF32 product of the rounded hi parts + the two correction terms as a single BF16
    MMA chain ([A_lo | A_hi] x [B_hi | B_lo], K-major).  Worst case ~2^-18 per product; the operator's tolerance is
    1e-5 * sum|terms|."""
    from pointwise_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(N * 1000 + K)
    A = rng.uniform(-1, 1, (128, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (N, K)).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    d = torch.full((128, N), float("nan"), device="cuda")
    _lib.check(L.conv3p_selftest_tc(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K, 1 | 8, None))
    torch.cuda.synchronize()
    got = d.cpu().numpy().astype(np.float64)
    assert np.isfinite(got).all()
    err = float((np.abs(got - want) / scale).max())
    print("tf32+bf16 corrections, K-major: max err / sum|terms| =", err)
    assert err < 2.5e-6, err


@pytest.mark.parametrize("N,K", [(64, 64), (32, 8), (128, 32), (64, 16), (256, 32)])
def test_tcgen05_mn_major_bf16_corrections(N, K):
    """Same split with the contraction index outermost in memory (weight-gradient layout): [lo ; hi] x [hi ; lo]
    stacked along K in 128B-swizzled MN-major BF16 panels."""
    from pointwise_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(N * 1000 + K + 7)
    A = rng.uniform(-1, 1, (K, 128)).astype(np.float32)
    B = rng.uniform(-1, 1, (K, N)).astype(np.float32)
    want = A.astype(np.float64).T @ B.astype(np.float64)
    scale = np.abs(A).astype(np.float64).T @ np.abs(B).astype(np.float64)
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    d = torch.full((128, N), float("nan"), device="cuda")
    _lib.check(L.conv3p_selftest_tc_mn(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K, 1 | 8, None))
    torch.cuda.synchronize()
    got = d.cpu().numpy().astype(np.float64)
    assert np.isfinite(got).all()
    err = float((np.abs(got - want) / scale).max())
    print("tf32+bf16 corrections, MN-major: max err / sum|terms| =", err)
    assert err < 2.5e-6, err
