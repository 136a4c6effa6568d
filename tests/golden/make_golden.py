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
    main()
