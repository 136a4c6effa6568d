"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's per-step input pipeline, the checker for
pointwise_b200/augment.py.  Only tests/ may import it; nothing in pointwise_b200/ does.

Parity pinned: tests/golden/augment_*.npz hold outputs of the reference's OWN functions (executed from
/root/reference by tests/golden/make_golden_augment.py) and tests/test_oracle.py checks this restatement against them.
"""
import numpy as np


def rotate_jitter(batch_data, angles=None, noise=None, sigma=0.01, clip=0.05):
    """rotate_point_cloud (modelnet_provider.py:23-41) then jitter_point_cloud (:64-75) with explicit draws:
    angles[k] is the value of `np.random.uniform() * 2 * np.pi` for cloud k (:33), noise the array
    `np.random.randn(B, N, C)` (:73).  Returns the float32 batch the provider stores (:211)."""
    data = np.asarray(batch_data, dtype=np.float32)
    out = data
    if angles is not None:
        out = np.zeros(data.shape, dtype=np.float32)                      # :31
        for k in range(data.shape[0]):
            c, s = np.cos(angles[k]), np.sin(angles[k])                    # :34-35
            rot = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])             # :36-38
            out[k, ...] = np.dot(data[k].reshape((-1, 3)), rot)            # :40 (float64 product -> float32)
    if noise is not None:
        assert clip > 0                                                    # :72
        j = np.clip(sigma * np.asarray(noise, dtype=np.float64), -1 * clip, clip)   # :73
        j += out                                                           # :74
        out = j
    return np.asarray(out, dtype=np.float32)


def sort_xyz_order(batch_data):
    """The permutation sort_point_cloud_xyz applies (util.py:66-70): argsort by z, then stable by y, then stable by
    x.  The reference's first pass uses numpy's default (unstable) sort; here it is stable too, which differs only
    in the order of rows whose x, y AND z are all equal."""
    data = np.asarray(batch_data)
    order = np.zeros(data.shape[:2], dtype=np.int32)
    for k in range(data.shape[0]):
        pc = data[k]
        idx = pc[:, 2].argsort(kind="mergesort")
        idx = idx[pc[idx, 1].argsort(kind="mergesort")]
        idx = idx[pc[idx, 0].argsort(kind="mergesort")]
        order[k] = idx
    return order


def sort_xyz(batch_data, batch_attributes=None):
    """sort_point_cloud_xyz (util.py:55-73) / sort_point_cloud_xyz2 (:75-109)."""
    order = sort_xyz_order(batch_data)
    take = lambda a: np.stack([a[k][order[k]] for k in range(a.shape[0])]) if a.shape[0] else a.copy()   # noqa: E731
    if batch_attributes is None:
        return take(np.asarray(batch_data, dtype=np.float32))
    return take(np.asarray(batch_data)), take(np.asarray(batch_attributes))
