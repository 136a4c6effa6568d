"""Drop-in consumers of the operator: PyTorch restatements of the reference's two PointConvNets
(SURVEY section 8f, rows N1/N2).  They exist to prove the drop-in claim at the reference's own call sites and to
exercise plan sharing; they are not part of the hot path.

* ``PointConvNetCls``  -- pointcnn2_acsd.py:33-90: four Conv3p layers (Cin->9, 9->9 x3, strides 1..4), SELU, concat
  of the 36 channels, flatten [B, N*36] -> FC 512 (SELU) -> alpha-dropout -> FC num_class (SELU);
  sparse softmax cross-entropy.
* ``PointConvNetSeg``  -- scene_seg/pointcnn_scene_seg_acsd.py:32-71: four 9-channel layers (strides 1..4), concat,
  one 36->num_class layer at stride 1, every layer followed by SELU; softmax cross-entropy per point.

Every Conv3p is followed by SELU in both networks; the layers call ``conv3p(..., activation="selu")``, which applies it
in the kernel's epilogue, and ``features_fused`` (inference) additionally writes the four 9-channel
outputs straight into the [B, N, 36] concat buffer and lets each layer read its predecessor's slice of it.

A ``PlanCache`` builds one neighbour plan per (points, stride) and hands it to every layer that needs it and to
the backward pass; the reference rebuilds its grid twice per layer per step (tf_conv3p_atrous.cpp:463, 629).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .ops import NeighborPlan, conv3p, conv3p_forward, parse_stride, parse_voxel


# False: every conv3p call builds its own plan (what a literal drop-in call without `plan=` does) -- bench.py's A/B of
# plan sharing.  Results are identical either way.
SHARE_PLANS = True
# Passed to NeighborPlan(check=...): True = learned capacity with the deferred overflow check (the default); False =
# learned capacity, poison only, nothing read back -- what a CUDA-graph capture of a training step needs (no host
# polling inside the captured region; tools/graph_net.py).
PLAN_CHECK = True


class PlanCache:
    """Neighbour plans of ONE batch of points, keyed by stride (voxel size fixed)."""

    def __init__(self, points: torch.Tensor, voxel_size=0.1):
        self.points = points
        self.voxel = parse_voxel(voxel_size)
        self.plans = {}

    def get(self, stride) -> NeighborPlan:
        s = parse_stride(stride)
        if not SHARE_PLANS:
            return NeighborPlan(self.points, s, self.voxel, check=PLAN_CHECK)
        if s not in self.plans:
            self.plans[s] = NeighborPlan(self.points, s, self.voxel, check=PLAN_CHECK)
        return self.plans[s]


@torch.no_grad()
def features_fused(plans: PlanCache, input_tensor: torch.Tensor, filters) -> torch.Tensor:
    """The stack of 9-channel Conv3p+SELU layers at strides 1, 2, 3, ... (pointcnn2_acsd.py:48-69,
    scene_seg/pointcnn_scene_seg_acsd.py:51-56) written straight into the concat buffer: layer i stores
    selu(conv3p(.)) in channels [9i, 9i+9) and layer i+1 gathers from that slice -- no activation pass, no concat."""
    B, N = input_tensor.shape[0], input_tensor.shape[1]
    widths = [int(w.shape[4]) for w in filters]
    concat = torch.empty((B, N, sum(widths)), dtype=torch.float32, device=input_tensor.device)
    x, c0 = input_tensor, 0
    for i, w in enumerate(filters):
        out = concat[:, :, c0:c0 + widths[i]]
        conv3p_forward(plans.get([i + 1] * 3), x, w, activation="selu", out=out)
        x, c0 = out, c0 + widths[i]
    return concat


def _filter(cin: int, cout: int) -> nn.Parameter:
    # tf.get_variable default: glorot-uniform over the [3,3,3,cin,cout] tensor (fan = 27*cin, 27*cout)
    w = torch.empty(3, 3, 3, cin, cout)
    bound = math.sqrt(6.0 / (27 * cin + 27 * cout))
    return nn.Parameter(w.uniform_(-bound, bound))


class PointConvNetCls(nn.Module):
    """pointcnn2_acsd.py:33-77."""

    def __init__(self, num_class: int, num_points: int, in_channels: int, voxel_size: float = 0.1):
        super().__init__()
        self.voxel = voxel_size
        self.filters = nn.ParameterList([_filter(in_channels, 9), _filter(9, 9), _filter(9, 9), _filter(9, 9)])
        self.fc1 = nn.Linear(num_points * 36, 512)
        self.fc2 = nn.Linear(512, num_class)

    def model(self, points_tensor: torch.Tensor, input_tensor: torch.Tensor, is_training: bool = True):
        plans = PlanCache(points_tensor, self.voxel)
        x, feats = input_tensor, []
        for i, w in enumerate(self.filters):                      # strides 1,2,3,4 -- :48-67
            stride = [i + 1] * 3
            x = conv3p(points_tensor, x, w, stride, [self.voxel], plan=plans.get(stride), activation="selu")
            feats.append(x)
        feat = torch.cat(feats, dim=2)                            # :69
        view = feat.reshape(feat.shape[0], -1)                    # :70
        fc1 = F.selu(self.fc1(view))                              # :71
        drop = F.alpha_dropout(fc1, p=0.5, training=is_training)  # selu.dropout_selu, :73
        return F.selu(self.fc2(drop))                             # :75

    @torch.no_grad()
    def infer(self, points_tensor: torch.Tensor, input_tensor: torch.Tensor) -> torch.Tensor:
        """Inference with the concat-free layout (no dropout): same values as ``model(..., is_training=False)``."""
        feat = features_fused(PlanCache(points_tensor, self.voxel), input_tensor, list(self.filters))
        return F.selu(self.fc2(F.selu(self.fc1(feat.reshape(feat.shape[0], -1)))))

    forward = model

    @staticmethod
    def loss(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        return F.cross_entropy(logits, labels)                    # sparse softmax CE, mean -- :88-89


class PointConvNetSeg(nn.Module):
    """scene_seg/pointcnn_scene_seg_acsd.py:39-58."""

    def __init__(self, num_class: int, in_channels: int, voxel_size: float = 0.1):
        super().__init__()
        self.voxel = voxel_size
        self.filters = nn.ParameterList([_filter(in_channels, 9), _filter(9, 9), _filter(9, 9), _filter(9, 9),
                                         _filter(36, num_class)])

    def model(self, points_tensor: torch.Tensor, input_tensor: torch.Tensor, is_training: bool = True):
        plans = PlanCache(points_tensor, self.voxel)
        x, feats = input_tensor, []
        for i in range(4):                                         # :51-54
            stride = [i + 1] * 3
            x = conv3p(points_tensor, x, self.filters[i], stride, [self.voxel], plan=plans.get(stride),
                       activation="selu")
            feats.append(x)
        concat = torch.cat(feats, dim=2)                           # :56
        # layer 5 reuses the stride-1 plan of layer 1
        return conv3p(points_tensor, concat, self.filters[4], [1, 1, 1], [self.voxel],
                      plan=plans.get([1, 1, 1]), activation="selu")  # :57

    @torch.no_grad()
    def infer(self, points_tensor: torch.Tensor, input_tensor: torch.Tensor) -> torch.Tensor:
        """Inference with the concat-free layout: same values as ``model`` without the SELU passes and the concat."""
        plans = PlanCache(points_tensor, self.voxel)
        concat = features_fused(plans, input_tensor, list(self.filters)[:4])
        return conv3p_forward(plans.get([1, 1, 1]), concat, self.filters[4], activation="selu")

    forward = model

    @staticmethod
    def loss(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        # tf.losses.softmax_cross_entropy(one_hot, logits): mean over all points -- :66-67
        return F.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1))
