"""Generates tests/golden/augment_*.npz by EXECUTING the reference's own functions where they lie under
/root/reference (nothing is copied: the function definitions are extracted from the source files at run time and
exec'd with numpy -- the modules themselves import TensorFlow / h5py, which this container does not have).

    python tests/golden/make_golden_augment.py        (needs /root/reference; the fixtures are committed)
"""
import ast
import os

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_functions(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return [ns[n] for n in names]


def main():
    rotate, jitter = load_functions(os.path.join(REF, "modelnet_provider.py"),
                                    ["rotate_point_cloud", "jitter_point_cloud"])
    sort1, sort2 = load_functions(os.path.join(REF, "util.py"), ["sort_point_cloud_xyz", "sort_point_cloud_xyz2"])
    rng = np.random.default_rng(2024)
    for tag, (B, N) in {"a": (3, 257), "b": (2, 1024)}.items():
        data = rng.uniform(-1, 1, (B, N, 3)).astype(np.float32)
        # the reference draws from the global numpy RNG: seed it, run, then replay the same draws
        np.random.seed(100 + B)
        out = jitter(rotate(data))
        np.random.seed(100 + B)
        angles = np.array([np.random.uniform() * 2 * np.pi for _ in range(B)])
        noise = np.random.randn(B, N, 3)
        np.savez_compressed(os.path.join(HERE, f"augment_rotate_jitter_{tag}.npz"), data=data, angles=angles,
                            noise=noise, out=np.asarray(out, dtype=np.float32))
        # xyz sort: coordinates on a coarse lattice so that x and y ties are common (the z pass matters), no exact
        # duplicate points (their order is unspecified in the reference: its first argsort is not stable)
        while True:
            pts = np.round(rng.uniform(-1, 1, (B, N, 3)) * 8) / 8
            pts[:, :, 2] += rng.permutation(N)[None, :] * 1e-3
            pts = pts.astype(np.float32)
            if all(len(np.unique(pts[k], axis=0)) == N for k in range(B)):
                break
        feats = rng.uniform(-1, 1, (B, N, 5)).astype(np.float32)
        both = np.concatenate([pts, feats[:, :, :2]], axis=2)
        s1 = sort1(both)
        s2d, s2a = sort2(pts, feats)
        np.savez_compressed(os.path.join(HERE, f"augment_sort_xyz_{tag}.npz"), data=both, sorted=s1, points=pts,
                            attributes=feats, sorted_points=s2d, sorted_attributes=s2a)
    print("wrote augment fixtures")


if __name__ == "__main__":
    main()
