"""pointwise_b200 -- a Blackwell-native (sm_100a) Conv3p pointwise-convolution operator.

Public surface (mirrors the reference's Python boundary, pointcnn2_acsd.py:9-31):

    conv3p(points, input, kernel, stride, voxel_size) -> output        (differentiable)
    conv3p_grad(grad, points, input, kernel, stride, voxel_size) -> (input_grad, filter_grad)
    NeighborPlan(points, stride, voxel_size)                            (shareable neighbour structure)

Everything computes in hand-written CUDA behind the C ABI of include/conv3p_b200.h; there is no CPU
or PyTorch fallback.
"""
from .ops import (NeighborPlan, conv3p, conv3p_backward, conv3p_forward, conv3p_grad,  # noqa: F401
                  launch_count, selu_backward, set_engine)
from ._lib import Conv3pError  # noqa: F401

__all__ = ["conv3p", "conv3p_grad", "conv3p_forward", "conv3p_backward", "NeighborPlan",
           "Conv3pError", "launch_count", "selu_backward", "set_engine"]
