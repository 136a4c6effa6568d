"""Generates tests/golden/*.npz from the REFERENCE's own CPU op (oracle/_ref, i.e.
/root/reference/tf_ops/conv3p/tf_conv3p_atrous.cpp compiled unmodified by oracle/Makefile).

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The vectors pin the oracle port and the CUDA path wherever the reference itself cannot travel.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from pointwise_b200.synth import make_problem  # noqa: E402

V = 0.1

KATS = {  # SURVEY section 8c; Cin = Cout = 1, W[f] = f
    "kat1": ([(0, 0, 0)], [2], [1], 1),
    "kat2": ([(0, 0, 0), (0.1, 0, 0)], [2, 3], [1, 10], 1),
    "kat3": ([(0, 0, 0), (0.15, 0, 0)], [2, 3], [1, 10], 1),
    "kat4": ([(0, 0, 0), (0.16, 0, 0)], [2, 3], [1, 10], 1),
    "kat5": ([(0, 0, 0), (0.06, 0, 0), (0.07, 0, 0)], [2, 3, 5], [1, 10, 100], 1),
    "kat6": ([(0, 0, 0), (0.1, 0, 0), (0.2, 0, 0)], [2, 3, 5], [1, 10, 100], 2),
    "kat7": ([(0, 0, 0), (0.1, 0.1, 0.1)], [2, 3], [1, 10], 1),
}

RANDOM = {  # name: (B, N, Cin, Cout, stride, dist, quantise, seed)
    "rand_sphere_s1": (2, 256, 4, 5, (1, 1, 1), "sphere", None, 11),
    "rand_room_s2": (2, 300, 3, 4, (2, 2, 2), "room", None, 12),
    "rand_cube_q_s1": (1, 400, 2, 3, (1, 1, 1), "cube", 0.05, 13),
    "rand_room_q_s3": (1, 350, 2, 2, (3, 3, 3), "room", 0.05, 14),
    "rand_room_aniso": (1, 300, 3, 2, (1, 2, 4), "room", None, 15),
}


# Filter shapes other than 3x3x3 (the reference reads fz, fy, fx from the tensor, tf_conv3p_atrous.cpp:425-427):
# name: (B, N, Cin, Cout, (fz, fy, fx), stride, dist, quantise, seed)
GENERAL = {
    "gen_555_s1": (2, 300, 3, 4, (5, 5, 5), (1, 1, 1), "room", None, 21),
    "gen_135_s2": (1, 400, 2, 3, (1, 3, 5), (2, 2, 2), "sphere", None, 22),
    "gen_222_q": (2, 256, 4, 2, (2, 2, 2), (1, 1, 1), "cube", 0.05, 23),
    "gen_331_aniso": (1, 350, 3, 3, (3, 3, 1), (1, 2, 3), "room", 0.05, 24),
    "gen_111": (1, 128, 5, 6, (1, 1, 1), (1, 1, 1), "cube", None, 25),
}

# T = double (register_op.cpp:45, 64; CPU kernels tf_conv3p_atrous.cpp:516, 727):
# name: (B, N, Cin, Cout, (fz, fy, fx), stride, dist or "lattice", seed)
DOUBLE = {
    "f64_333_s1": (2, 300, 3, 4, (3, 3, 3), (1, 1, 1), "room", 41),
    "f64_333_aniso": (1, 350, 2, 3, (3, 3, 3), (1, 2, 3), "sphere", 42),
    "f64_222": (2, 200, 3, 2, (2, 2, 2), (1, 1, 1), "cube", 43),
    # 7x7x7 lattice at half-voxel spacing, every coordinate moved by a double far below float resolution: in exact
    # arithmetic most points sit on cell boundaries, and only the double bits decide on which side
    "f64_333_subfloat": (1, 343, 2, 2, (3, 3, 3), (1, 1, 1), "lattice", 44),
}


def run64(R, points, inp, filt, gout, stride):
    v = 0.1
    out = R.forward64(points, inp, filt, stride, v)
    gi, gf = R.backward64(gout, points, inp, filt, stride, v)
    cnt = np.stack([R.neighbor_count64(points[b], stride, v, dims=filt.shape[:3]) for b in range(points.shape[0])])
    return dict(points=points, input=inp, filter=filt, grad_out=gout,
                stride=np.asarray(np.broadcast_to(stride, (3,)), np.int32), voxel=np.float64(v),
                output=out, grad_input=gi, grad_filter=gf, count_table=cnt)


def main64():
    R = oracle.Ref(single_thread=True)
    for name, (B, N, Ci, Co, dims, s, dist, seed) in DOUBLE.items():
        rng = np.random.default_rng(seed)
        if dist == "lattice":
            g = np.arange(7, dtype=np.float64) * 0.05 + 0.1
            P = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(1, -1, 3)
            P = P + rng.uniform(-1e-11, 1e-11, P.shape)
            P = P[:, rng.permutation(P.shape[1])]
        else:
            P = make_problem(B, N, Ci, Co, dist, seed=seed)["points"].astype(np.float64)
            P = P + rng.uniform(-1e-8, 1e-8, P.shape)
        X = rng.uniform(-1, 1, (B, N, Ci))
        W = rng.uniform(-0.1, 0.1, (*dims, Ci, Co))
        G = rng.uniform(-1, 1, (B, N, Co))
        fix = run64(R, P, X, W, G, s)
        if dist == "lattice":    # the fixture must tell a float predicate from a double one
            c32 = R.neighbor_count(P[0].astype(np.float32), s, np.float32(0.1), dims=dims)
            assert not np.array_equal(c32, fix["count_table"][0]), "lattice fixture does not discriminate"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **fix)
    print("wrote", len(DOUBLE), "double fixtures to", HERE)


def run(R, points, inp, filt, gout, stride):
    out = R.forward(points, inp, filt, stride, V)
    gi, gf = R.backward(gout, points, inp, filt, stride, V)
    cnt = np.stack([R.neighbor_count(points[b], stride, V, dims=filt.shape[:3]) for b in range(points.shape[0])])
    return dict(points=points, input=inp, filter=filt, grad_out=gout,
                stride=np.asarray(np.broadcast_to(stride, (3,)), np.int32), voxel=np.float32(V),
                output=out, grad_input=gi, grad_filter=gf, count_table=cnt)


def main():
    R = oracle.Ref(single_thread=True)   # single thread: grad_filter summed in one fixed order
    for name, (pts, inp, g, s) in KATS.items():
        P = np.array([pts], np.float32)
        X = np.array(inp, np.float32).reshape(1, -1, 1)
        G = np.array(g, np.float32).reshape(1, -1, 1)
        W = np.arange(27, dtype=np.float32).reshape(3, 3, 3, 1, 1)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run(R, P, X, W, G, s))
    for name, (B, N, Ci, Co, s, dist, q, seed) in RANDOM.items():
        pr = make_problem(B, N, Ci, Co, dist, seed=seed, quantise=q)
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            **run(R, pr["points"], pr["input"], pr["filter"], pr["grad_out"], s))
    for name, (B, N, Ci, Co, dims, s, dist, q, seed) in GENERAL.items():
        pr = make_problem(B, N, Ci, Co, dist, seed=seed, quantise=q)
        filt = np.random.default_rng(seed).uniform(-0.1, 0.1, (*dims, Ci, Co)).astype(np.float32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            **run(R, pr["points"], pr["input"], filt, pr["grad_out"], s))
    print("wrote", len(KATS) + len(RANDOM) + len(GENERAL), "fixtures to", HERE)


if __name__ == "__main__":
    if sys.argv[1:] == ["f64"]:      # only the double fixtures (the float ones stay byte-identical in git)
        main64()
    else:
        main()
        main64()
