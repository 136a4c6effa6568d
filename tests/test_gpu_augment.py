"""GPU tests of the input pipeline: pointwise_b200.augment against the numpy checker
(oracle/augment_oracle.py) and against outputs of the reference's own functions (tests/golden/augment_*.npz).

Bar: the xyz sort is index work -> bit-exact.  rotate + jitter is floating point: the reference multiplies float32
rows by a float64 matrix through BLAS (summation order / FMA use unspecified) and rounds to float32, so the bar is
1 float32 ulp of the result; on these inputs it is met exactly."""
import os

import numpy as np
import pytest
import torch

from oracle import augment_oracle as ao

pytestmark = pytest.mark.gpu
GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def assert_within_one_ulp(got, want):
    got, want = np.asarray(got, np.float32), np.asarray(want, np.float32)
    ulp = np.spacing(np.abs(want)).astype(np.float64)
    assert (np.abs(got.astype(np.float64) - want.astype(np.float64)) <= ulp).all()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_rotate_jitter_matches_reference_outputs(tag):
    from pointwise_b200.augment import rotate_jitter
    d = np.load(os.path.join(GOLDEN_DIR, f"augment_rotate_jitter_{tag}.npz"))
    got = rotate_jitter(dev(d["data"]), dev(d["angles"]), dev(d["noise"])).cpu().numpy()
    assert_within_one_ulp(got, d["out"])
    assert (got == d["out"]).mean() > 0.999


@pytest.mark.parametrize("B,N,sigma,clip", [(4, 1024, 0.01, 0.05), (1, 1, 0.5, 0.2), (3, 4097, 0.02, 0.01), (0, 16, 0.01, 0.05)])
def test_rotate_jitter_matches_checker(B, N, sigma, clip):
    from pointwise_b200.augment import rotate_jitter
    rng = np.random.default_rng(B * 1000 + N)
    data = rng.uniform(-2, 2, (B, N, 3)).astype(np.float32)
    angles, noise = rng.uniform(0, 2 * np.pi, B), rng.standard_normal((B, N, 3))
    got = rotate_jitter(dev(data), dev(angles), dev(noise), sigma, clip).cpu().numpy()
    assert_within_one_ulp(got, ao.rotate_jitter(data, angles, noise, sigma, clip))
    # the stages on their own, and the properties they must have at any size
    rot = rotate_jitter(dev(data), dev(angles), None).cpu().numpy()
    assert_within_one_ulp(rot, ao.rotate_jitter(data, angles, None))
    assert np.array_equal(rot[:, :, 1], data[:, :, 1])                       # the up axis is untouched
    np.testing.assert_allclose(np.linalg.norm(rot, axis=2), np.linalg.norm(data, axis=2), rtol=1e-6, atol=1e-6)
    jit = rotate_jitter(dev(data), None, dev(noise), sigma, clip).cpu().numpy()
    assert_within_one_ulp(jit, ao.rotate_jitter(data, None, noise, sigma, clip))
    assert (np.abs(jit.astype(np.float64) - data) <= clip + 1e-6).all()       # clipped offsets


def test_rotate_jitter_rejects_bad_arguments():
    from pointwise_b200.augment import rotate_jitter
    x = torch.zeros(2, 8, 3, device="cuda")
    with pytest.raises(ValueError):
        rotate_jitter(x, None, None, clip=0.0)
    with pytest.raises(ValueError):
        rotate_jitter(torch.zeros(2, 8, 4, device="cuda"))
    with pytest.raises(RuntimeError):
        rotate_jitter(torch.zeros(2, 8, 3))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_sort_xyz_matches_reference_outputs(tag):
    from pointwise_b200.augment import sort_xyz
    d = np.load(os.path.join(GOLDEN_DIR, f"augment_sort_xyz_{tag}.npz"))
    assert np.array_equal(sort_xyz(dev(d["data"])).cpu().numpy(), d["sorted"])
    sp, sa = sort_xyz(dev(d["points"]), dev(d["attributes"]))
    assert np.array_equal(sp.cpu().numpy(), d["sorted_points"])
    assert np.array_equal(sa.cpu().numpy(), d["sorted_attributes"])


@pytest.mark.parametrize("B,N,K,lattice", [(5, 4096, 3, 0), (2, 1000, 6, 4), (1, 1, 3, 0), (3, 513, 4, 2), (0, 7, 3, 0)])
def test_sort_xyz_order_is_bit_exact(B, N, K, lattice):
    """Random clouds, lattice clouds with many equal x / y / whole points (ties -> original order), negative zeros."""
    from pointwise_b200.augment import sort_xyz
    rng = np.random.default_rng(N + K)
    data = rng.uniform(-1, 1, (B, N, K)).astype(np.float32)
    if lattice:
        data[:, :, :3] = np.round(data[:, :, :3] * lattice) / lattice           # includes -0.0
    out, order = sort_xyz(dev(data), return_order=True)
    want = ao.sort_xyz_order(data)
    assert np.array_equal(order.cpu().numpy(), want)
    assert np.array_equal(out.cpu().numpy(), ao.sort_xyz(data))
    if B and N:
        for k in range(B):                                                     # a permutation, lexicographically sorted
            assert np.array_equal(np.sort(order[k].cpu().numpy()), np.arange(N))
            keys = [tuple(r[:3] + 0.0) for r in out[k].cpu().numpy()]
            assert keys == sorted(keys)
