"""Input pipeline on the GPU: the per-step host work of the reference's data providers.

* ``rotate_jitter``  -- ``rotate_point_cloud`` + ``jitter_point_cloud`` (modelnet_provider.py:23-41, 64-75), applied
  by ``get_batch_point_cloud`` to every training batch (:195-198);
* ``sort_xyz``       -- ``sort_point_cloud_xyz`` / ``sort_point_cloud_xyz2`` (util.py:55-109), applied when
  ``sort_cloud`` is set (:202-204, param.json).

Same argument meaning as the reference; the random draws are explicit so that a given draw gives the same batch.
Everything runs in the CUDA library (include/conv3p_b200.h); there is no CPU path here.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib
from .ops import _ptr, _stream_ptr


def _cuda(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (pointwise_b200 has no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype} (got {t.dtype})")
    return t.contiguous()


def draw_augmentation(B: int, N: int, device, generator: Optional[torch.Generator] = None):
    """The draws the reference makes per batch: angles = U[0,1) * 2*pi per cloud (modelnet_provider.py:33),
    noise = N(0,1) per coordinate (:73), float64 like numpy's."""
    angles = torch.rand(B, dtype=torch.float64, device=device, generator=generator) * (2.0 * math.pi)
    noise = torch.randn(B, N, 3, dtype=torch.float64, device=device, generator=generator)
    return angles, noise


def rotate_jitter(batch_data: torch.Tensor, angles: Optional[torch.Tensor] = None,
                  noise: Optional[torch.Tensor] = None, sigma: float = 0.01, clip: float = 0.05) -> torch.Tensor:
    """jitter_point_cloud(rotate_point_cloud(batch_data)) for [B,N,3] float32 clouds with explicit draws
    (``angles`` [B] float64 radians, ``noise`` [B,N,3] float64 standard normal; None skips the stage)."""
    data = _cuda(batch_data, torch.float32, "batch_data")
    if data.dim() != 3 or data.shape[2] != 3:
        raise ValueError("rotate_jitter expects a [B, N, 3] batch")
    if not clip > 0:
        raise ValueError("clip must be positive")      # assert(clip > 0), modelnet_provider.py:72
    B, N = int(data.shape[0]), int(data.shape[1])
    if angles is not None:
        angles = _cuda(angles, torch.float64, "angles")
        if tuple(angles.shape) != (B,):
            raise ValueError("angles must have shape [B]")
    if noise is not None:
        noise = _cuda(noise, torch.float64, "noise")
        if tuple(noise.shape) != (B, N, 3):
            raise ValueError("noise must have shape [B, N, 3]")
    out = torch.empty_like(data)
    with torch.cuda.device(data.device):
        _lib.check(_lib.lib().conv3p_augment_rotate_jitter_f32(
            _ptr(data), _ptr(angles), _ptr(noise), float(sigma), float(clip), B, N, _ptr(out),
            _stream_ptr(data.device)))
    return out


def sort_xyz(batch_data: torch.Tensor, batch_attributes: Optional[torch.Tensor] = None, return_order: bool = False):
    """sort_point_cloud_xyz(batch_data) / sort_point_cloud_xyz2(batch_data, batch_attributes): rows of every cloud
    ordered by x, then y, then z (first three channels), ties by original position.
    -> sorted_data [, sorted_attributes] [, order (int32 [B,N], source row of every output row)]."""
    data = _cuda(batch_data, torch.float32, "batch_data")
    if data.dim() != 3 or data.shape[2] < 3:
        raise ValueError("sort_xyz expects a [B, N, K>=3] batch")
    B, N, K = (int(v) for v in data.shape)
    attrs, M = None, 0
    if batch_attributes is not None:
        attrs = _cuda(batch_attributes, torch.float32, "batch_attributes")
        if attrs.dim() != 3 or tuple(attrs.shape[:2]) != (B, N):
            raise ValueError("batch_attributes must have shape [B, N, M]")
        M = int(attrs.shape[2])
    L = _lib.lib()
    order = torch.empty((B, N), dtype=torch.int32, device=data.device)
    out = torch.empty_like(data)
    out_attrs = torch.empty_like(attrs) if attrs is not None else None
    nws = L.conv3p_xyz_sort_workspace_bytes(B, N)
    ws = torch.empty(max(nws, 1), dtype=torch.uint8, device=data.device)
    with torch.cuda.device(data.device):
        _lib.check(L.conv3p_xyz_sort_f32(_ptr(data), K, _ptr(attrs), M, B, N, _ptr(order), _ptr(out),
                                         _ptr(out_attrs), _ptr(ws), nws, _stream_ptr(data.device)))
    res = (out,) + ((out_attrs,) if attrs is not None else ()) + ((order,) if return_order else ())
    return res[0] if len(res) == 1 else res
