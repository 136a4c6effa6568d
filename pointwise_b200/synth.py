"""Synthetic (B, N, C) point clouds for parity tests and the benchmark.

There are no datasets in this environment, so inputs are generated with a portable numpy
generator (``numpy.random.default_rng(seed)``, everything cast to float32).  Three coordinate
distributions stand in for the reference's data:

* ``cube``    U(-1,1)^3 -- volumetric.
* ``sphere``  normalised Gaussian, unit radius -- a ModelNet40-like surface
  (reference normalises shapes to the unit sphere, modelnet_provider.py).
* ``room``    an S3DIS-like 1 x 1 x 3 m block (s3dis_provider.py feeds 4096-point blocks):
  floor, ceiling, two walls and clutter, one fifth each.

``quantise`` snaps coordinates to a lattice (e.g. 0.05) to put many points exactly on bin edges --
the stress case for the reference's non-symmetric neighbour quirk (tf_conv3p_atrous.cpp:679).
"""
from __future__ import annotations

import numpy as np


def make_points(B: int, N: int, dist: str = "room", seed: int = 0, quantise: float | None = None,
                sort_xyz: bool = False) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if dist == "cube":
        p = rng.uniform(-1.0, 1.0, size=(B, N, 3))
    elif dist == "sphere":
        g = rng.standard_normal(size=(B, N, 3))
        p = g / np.maximum(np.linalg.norm(g, axis=-1, keepdims=True), 1e-12)
    elif dist == "room":
        comp = rng.integers(0, 5, size=(B, N))
        u = rng.uniform(0.0, 1.0, size=(B, N, 3))
        x, y, z = u[..., 0].copy(), u[..., 1].copy(), u[..., 2] * 3.0
        thin = u[..., 2] * 0.01
        z = np.where(comp == 0, thin, z)                       # floor
        z = np.where(comp == 1, 2.99 + thin, z)                # ceiling
        x = np.where(comp == 2, u[..., 0] * 0.01, x)           # wall x ~ 0
        y = np.where(comp == 3, u[..., 1] * 0.01, y)           # wall y ~ 0
        z = np.where(comp == 4, u[..., 2] * 1.2, z)            # clutter
        p = np.stack([x, y, z], axis=-1)
    else:
        raise ValueError(f"unknown distribution {dist!r}")
    if quantise:
        p = np.round(p / quantise) * quantise
    p = p.astype(np.float32)
    if sort_xyz:
        # the reference feeds clouds sorted by x, then y, then z (util.py:55-73, param.json:8-9)
        for b in range(B):
            order = np.lexsort((p[b, :, 2], p[b, :, 1], p[b, :, 0]))
            p[b] = p[b, order]
    return p


def make_problem(B: int, N: int, Cin: int, Cout: int, dist: str = "room", seed: int = 0,
                 quantise: float | None = None, sort_xyz: bool = False):
    """-> dict(points[B,N,3], input[B,N,Cin], filter[3,3,3,Cin,Cout], grad_out[B,N,Cout]) float32."""
    points = make_points(B, N, dist, seed, quantise, sort_xyz)
    rng = np.random.default_rng(seed + 1000003)
    return dict(
        points=points,
        input=rng.uniform(-1.0, 1.0, size=(B, N, Cin)).astype(np.float32),
        filter=rng.uniform(-0.1, 0.1, size=(3, 3, 3, Cin, Cout)).astype(np.float32),
        grad_out=rng.uniform(-1.0, 1.0, size=(B, N, Cout)).astype(np.float32),
    )
