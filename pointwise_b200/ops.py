"""The Conv3p operator on PyTorch CUDA tensors -- host-side mirror of the reference's Python boundary.

Reference interface being mirrored (hkust-vgd/pointwise):

* ``conv3p(points_tensor, input_tensor, kernel_tensor, stride, voxel_size)``
  -- ``pointcnn2_acsd.py:12-13``, ``scene_seg/pointcnn_scene_seg_acsd.py:11-12``
* the registered gradient returning ``[None, input_grad, filter_grad, None, None]``
  -- ``pointcnn2_acsd.py:15-31``
* argument validation of the op -- ``tf_ops/conv3p/tf_conv3p_atrous.cpp:410-444, 583-585``
  (same conditions and messages, raised as ``ValueError``).

PyTorch is plumbing here (device memory, the current stream, autograd bookkeeping); all compute is
hand-written sm_100a CUDA behind the C ABI of ``include/conv3p_b200.h``.  There is no CPU path.
"""
from __future__ import annotations

import os

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib

NCELL = 27


# --------------------------------------------------------------------------------------------------
# argument handling
# --------------------------------------------------------------------------------------------------
def _host_list(x, what: str):
    """stride / voxel_size may be a python scalar, a sequence, a numpy array or a tensor.  A CUDA
    tensor forces one device->host read (the reference's GPU op does the same, blocking, at
    tf_conv3p_atrous.cu:577,586); pass host values to stay asynchronous."""
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().reshape(-1).tolist()
    if hasattr(x, "tolist"):
        x = x.tolist()
    if isinstance(x, (int, float)):
        return [x]
    try:
        return list(x)
    except TypeError as e:
        raise ValueError(f"Conv3p: cannot interpret {what}={x!r}") from e


def parse_stride(stride) -> tuple:
    s = _host_list(stride, "stride")
    if len(s) == 1 and not isinstance(stride, torch.Tensor) and not hasattr(stride, "shape"):
        s = s * 3  # python scalar convenience: isotropic
    if len(s) != 3:
        raise ValueError("Conv3p expects stride tensor to have size 3.")  # tf_conv3p_atrous.cpp:437
    s = tuple(int(v) for v in s)
    if any(v < 1 for v in s):
        raise ValueError("Conv3p expects strides >= 1")
    return s


def parse_voxel(voxel_size) -> float:
    v = _host_list(voxel_size, "voxel_size")
    if len(v) != 1:
        raise ValueError("Conv3p expects voxel tensor to have dimension 1.")  # :443
    v = float(v[0])
    if not v > 0:
        raise ValueError("Conv3p expects voxel_size > 0")
    return v


def _check_cuda_f32(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"Conv3p: {name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"Conv3p: {name} must be a CUDA tensor (pointwise_b200 has no CPU fallback)")
    if t.dtype != dtype:
        what = "float32" if dtype == torch.float32 else "float64 like the other tensors of the call"
        raise TypeError(f"Conv3p: {name} must be {what} (got {t.dtype})")
    return t.contiguous()


def _op_dtype(points) -> torch.dtype:
    """T of the call (register_op.cpp:45: {float, double}): float64 when the points are float64 -- every tensor of the
    call must then be float64, as in the reference where one attribute T types them all."""
    return torch.float64 if isinstance(points, torch.Tensor) and points.dtype == torch.float64 else torch.float32


def validate(points, input, kernel):
    """Shape checks of Conv3pOp::Compute (tf_conv3p_atrous.cpp:410-430), same messages."""
    if points.dim() != 3:
        raise ValueError("Conv3p expects (batch_size, num_points, 3) points shape")  # :410
    if points.shape[2] != 3:
        raise ValueError("Conv3p expects (batch_size, num_points, 3) points shape")
    if input.dim() != 3 or input.shape[0] != points.shape[0]:
        raise ValueError("Conv3p expects points and input tensor to have the same batch size")  # :417
    if input.shape[1] != points.shape[1]:
        raise ValueError("Conv3p expects points and input tensor to have the same number of points")  # :418
    if kernel.dim() != 5:
        raise ValueError("Conv3p expects a [fz, fy, fx, in_channels, out_channels] filter")
    if kernel.shape[3] != input.shape[2]:
        raise ValueError("Conv3p expects filter channels to be matched with input channels")  # :430
    if any(int(d) < 1 for d in kernel.shape[:3]):
        raise ValueError("Conv3p expects a [fz, fy, fx, in_channels, out_channels] filter")


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


# CONV3P_PREFETCH_BACKWARD=0 keeps the backward lists on the caller's stream (A/B timing).  Measured (one B200,
# fwd+bwd step with / without the side stream): 262,144 points 3.956 / 3.973 ms, 65,536 points 0.833 / 0.864 ms
# (9->9) and 2.155 / 2.210 ms (36->13) -- but 32,768 points 0.528 / 0.358 ms: below ~50k points the step is so short
# that the cross-stream hand-off and the allocator's deferred reuse of the plan buffer cost more than the overlap
# gains, so small batches stay on one stream.
PREFETCH_BACKWARD = os.environ.get("CONV3P_PREFETCH_BACKWARD", "1") != "0"
PREFETCH_MIN_POINTS = int(os.environ.get("CONV3P_PREFETCH_MIN_POINTS", "49152"))
_side_streams = {}


def _side_stream(device, role: str = "lists") -> torch.cuda.Stream:
    key = (torch.device(device).index, role)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device)
    return _side_streams[key]


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


# --------------------------------------------------------------------------------------------------
# neighbour plan
# --------------------------------------------------------------------------------------------------
_capacity_hint: dict = {}
HEADROOM = float(os.environ.get("CONV3P_CAPACITY_HEADROOM", "2.0"))   # learned pair capacity = pairs seen x this
_pending_checks: list = []     # plans whose deferred overflow check has not been read yet (weak references)


def _poll_pending_checks() -> None:
    """Non-blocking: reads the overflow flag of earlier plans whose header copy has arrived; raises if one of them
    overflowed (its outputs were NaN-poisoned on the device)."""
    alive = []
    err = None
    for ref in _pending_checks:
        plan = ref()
        if plan is None or plan._stats_event is None:
            continue
        try:
            if not plan.verify(block=False):
                alive.append(ref)
        except _lib.Conv3pError as e:      # report the first, keep draining the list
            err = err or e
    _pending_checks[:] = alive
    if err is not None:
        raise err


class _StatsSlots:
    """Pool of 128-byte slots of pinned (device-addressable) host memory that plans publish their counters into."""
    _free: list = []
    _busy: list = []       # (slot, event): released by a plan that died before its counters were read
    _blocks: list = []

    @classmethod
    def take(cls) -> torch.Tensor:
        if cls._busy:
            still = []
            for slot, ev in cls._busy:
                if ev.query():
                    cls._free.append(slot)
                else:
                    still.append((slot, ev))
            cls._busy = still
        if not cls._free:
            block = torch.zeros(64 * 16, dtype=torch.int64).pin_memory()
            cls._blocks.append(block)
            cls._free.extend(block[i * 16:(i + 1) * 16] for i in range(64))
        return cls._free.pop()

    @classmethod
    def give(cls, slot: torch.Tensor, pending_event) -> None:
        if pending_event is None:
            cls._free.append(slot)
        else:
            cls._busy.append((slot, pending_event))     # the publishing kernel may still be in flight


class NeighborPlan:
    """Voxel-sorted neighbour structure of one batch of clouds for one (stride, voxel_size).

    Holds the count table ``[B,N,27]`` (bit-identical to the reference's ``neighbor_count`` table,
    tf_conv3p_atrous.cpp:369-379), the cell-grouped forward lists and, on demand, the backward lists.
    A plan depends on ``points`` only, so it can be shared by every layer that uses the same stride and
    by the forward and backward pass (the reference rebuilds its grid twice per layer per step,
    tf_conv3p_atrous.cpp:463, 629).

    ``capacity`` bounds the total number of (point, neighbour) pairs (data dependent).  ``capacity=None`` sizes it
    automatically and WITHOUT a host synchronisation in steady state ("no host synchronisation inside
    forward/backward"; the reference's GPU op blocks twice per call, tf_conv3p_atrous.cu:577, 586):

    * the first plan of a (B, N, stride, voxel) shape is built with a synchronous check (one 128-byte read-back,
      rebuilt larger if the guess was too small) and leaves a grow-only estimate with 100 % head-room (``HEADROOM``:
      the lists cost 12 bytes per pair of CAPACITY in the plan buffer and nothing at run time -- the kernels only touch
      what is filled -- so the estimate is generous: real scans vary far more from batch to batch than synthetic ones);
    * every later plan of that shape uses the estimate and only ENQUEUES a copy of the plan's counters into pinned
      host memory.  They are looked at later, without waiting -- when the backward pass starts, when the next plan
      is built -- or on ``verify()`` / ``stats``.  Should a batch ever exceed the estimate,
      the affected outputs were NaN-poisoned on the device (never silently wrong), the estimate is raised and a
      ``Conv3pError(ERR_PAIR_OVERFLOW)`` is raised at that point: rerun the step.

    ``check="sync"`` always checks synchronously, ``check="deferred"`` never does (also with an explicit
    ``capacity``, which is otherwise checked synchronously); ``check=False`` never reads anything back (poison only).
    """

    def __init__(self, points: torch.Tensor, stride, voxel_size, capacity: Optional[int] = None,
                 check=True):
        points = _check_cuda_f32(points, "points")
        if points.dim() != 3 or points.shape[2] != 3:
            raise ValueError("Conv3p expects (batch_size, num_points, 3) points shape")
        self.points = points
        self.stride = parse_stride(stride)
        self.voxel_size = parse_voxel(voxel_size)
        self.B, self.N = int(points.shape[0]), int(points.shape[1])
        self.device = points.device
        self.has_backward = False
        self._stats = None
        self._stats_event = None
        self._stats_slot = None
        # Side streams that read or write the plan buffer (deferred counters, backward-list prefetch) leave their events
        # here; __del__ makes the ALLOCATING stream wait for them before the buffer goes back to the caching allocator.
        # (Tensor.record_stream would do the same job lazily -- but then a buffer can only be reused once the
        # allocator has SEEN the event complete, and a host that enqueues many steps ahead makes the pool grow by one
        # plan buffer per step in flight: half-GB cudaMallocs, 15-45 ms each, at arbitrary steps.)
        self._alloc_stream = torch.cuda.current_stream(self.device)
        self._side_events = []
        L = _lib.lib()
        self._key = key = (self.B, self.N, self.stride, self.voxel_size)
        pts = self.B * self.N
        if check:
            _poll_pending_checks()
        learned = capacity is None and key in _capacity_hint
        sync_check = check == "sync" or (check is True and not learned)
        self._auto = capacity is None
        cap = int(capacity) if capacity is not None else _capacity_hint.get(key, max(1024, 48 * pts))
        while True:
            self.capacity = cap
            self.geom = _lib.make_geom(self.B, self.N, self.stride, self.voxel_size, cap)
            nbytes = L.conv3p_plan_bytes(self.geom)
            if nbytes == 0:
                raise ValueError("Conv3p: invalid geometry")
            self.buffer = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            with torch.cuda.device(self.device):
                _lib.check(L.conv3p_plan_build_f32(self.geom, _ptr(points), _ptr(self.buffer), nbytes,
                                                   _stream_ptr(self.device)))
            if not sync_check:
                break
            st = self.read_stats()
            if not st.overflow:
                break
            if capacity is not None:
                raise _lib.Conv3pError(_lib.ERR_PAIR_OVERFLOW,
                                       f"pair capacity {cap} too small, {st.total_pairs} pairs needed")
            cap = int(st.total_pairs * 1.125) + 1024
        if self._auto and sync_check:
            self._learn(self._stats.total_pairs)
        lay = _lib.PlanLayout()
        _lib.check(L.conv3p_plan_layout(self.geom, lay))
        self.layout = lay
        if check and not sync_check and pts > 0:     # check == "deferred", or True with a learned estimate
            self._enqueue_header_copy()
        # the forward lists are complete at this point of the stream: a later prefetch_backward() starts from here
        self._searched = torch.cuda.Event()
        self._searched.record(torch.cuda.current_stream(self.device))
        self._bwd_ready = None

    def _learn(self, total_pairs: int) -> None:
        """grow-only capacity estimate for the next batch of this shape"""
        want = int(total_pairs * HEADROOM) + 1024
        _capacity_hint[self._key] = max(_capacity_hint.get(self._key, 0), want)

    # ---- deferred overflow check -----------------------------------------------------------------
    def _enqueue_header_copy(self) -> None:
        import weakref
        self._stats_slot = _StatsSlots.take()
        # On its own stream, ordered after the plan build only: the store into host memory crosses PCIe and may sit
        # behind an application's bulk device->host copies; nothing on the compute stream ever waits for it.
        built = torch.cuda.Event()
        built.record(torch.cuda.current_stream(self.device))
        side = _side_stream(self.device, "stats")
        side.wait_event(built)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().conv3p_plan_publish_stats(self.geom, _ptr(self.buffer),
                                                            C.c_void_p(self._stats_slot.data_ptr()),
                                                            C.c_void_p(side.cuda_stream)))
        self._stats_event = torch.cuda.Event()
        self._stats_event.record(side)
        self._side_events.append(self._stats_event)
        _pending_checks.append(weakref.ref(self))

    def __del__(self):
        try:      # the buffer is released in the allocating stream's order: that stream waits for the side streams first
            for ev in getattr(self, "_side_events", ()):
                self._alloc_stream.wait_event(ev)
        except Exception:      # interpreter shutdown, CUDA context already gone
            pass
        slot = getattr(self, "_stats_slot", None)
        if slot is not None:
            ev = getattr(self, "_stats_event", None)
            _StatsSlots.give(slot, ev)
            self._stats_slot = None

    def verify(self, block: bool = True) -> bool:
        """Looks at the deferred copy of the plan's counters.  Returns False when it has not arrived yet and
        ``block`` is False; raises Conv3pError(ERR_PAIR_OVERFLOW) if the lists were incomplete."""
        ev = self._stats_event
        if ev is None:
            return True
        if block:
            ev.synchronize()          # waits for the plan build only (recorded right after it), not for the stream
        elif not ev.query():
            return False
        self._stats_event = None
        h = self._stats_slot.tolist()
        _StatsSlots.give(self._stats_slot, None)
        self._stats_slot = None
        st = _lib.PlanStats()
        st.total_pairs, st.backward_pairs = int(h[0]), int(h[2])
        st.overflow = 1 if (h[1] != 0 or h[0] > self.capacity) else 0
        st.has_backward = 1 if h[3] != 0 else 0
        self._stats = st
        if self._auto:
            self._learn(st.total_pairs)
        if st.overflow:
            raise _lib.Conv3pError(
                _lib.ERR_PAIR_OVERFLOW,
                f"pair capacity {self.capacity} was too small for this batch ({st.total_pairs} pairs): the affected "
                "outputs were NaN-poisoned; the capacity estimate has been raised, rerun the step")
        return True

    @property
    def stats(self):
        """Counters of the plan (reads them back on first use)."""
        if self._stats is None:
            if self._stats_event is not None:
                self.verify(block=True)
            else:
                self.read_stats()
        return self._stats

    # ---- stats / views -------------------------------------------------------------------------
    def read_stats(self):
        """Counters of the plan (synchronises the current stream)."""
        st = _lib.PlanStats()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().conv3p_plan_stats(self.geom, _ptr(self.buffer), st,
                                                    _stream_ptr(self.device)))
        self._stats = st
        return st

    def _view(self, offset: int, count: int, dtype: torch.dtype) -> torch.Tensor:
        nbytes = count * torch.empty((), dtype=dtype).element_size()
        return self.buffer[offset:offset + nbytes].view(dtype)

    @property
    def count_table(self) -> torch.Tensor:
        return self._view(self.layout.count_table, self.B * self.N * NCELL, torch.int32) \
            .view(self.B, self.N, NCELL)

    @property
    def backward_count_table(self) -> torch.Tensor:
        return self._view(self.layout.bwd_count, self.B * self.N * NCELL, torch.int32) \
            .view(self.B, self.N, NCELL)

    @property
    def pair_begin(self) -> torch.Tensor:
        return self._view(self.layout.pair_begin, self.B * self.N, torch.int64).view(self.B, self.N)

    @property
    def pair_len(self) -> torch.Tensor:
        return self._view(self.layout.pair_len, self.B * self.N, torch.int32).view(self.B, self.N)

    @property
    def pair_row(self) -> torch.Tensor:
        return self._view(self.layout.pair_row, self.capacity, torch.int32)

    @property
    def backward_row(self) -> torch.Tensor:
        return self._view(self.layout.bwd_row, self.capacity, torch.int32)

    @property
    def backward_weight(self) -> torch.Tensor:
        return self._view(self.layout.bwd_weight, self.capacity, torch.float32)

    @property
    def sorted_xyzi(self) -> torch.Tensor:
        return self._view(self.layout.sorted_xyzi, self.B * self.N * 4, torch.float32) \
            .view(self.B, self.N, 4)

    @property
    def sorted_key(self) -> torch.Tensor:
        return self._view(self.layout.sorted_key, self.B * self.N, torch.int32).view(self.B, self.N)

    # ---- backward lists ----------------------------------------------------------------------------
    def prefetch_backward(self) -> "NeighborPlan":
        """Builds the backward lists on a side stream, ordered after the neighbour search only -- called once the
        forward kernels are enqueued, the list kernel fills whatever the persistent forward CTAs leave free instead
        of sitting between forward and backward on the main stream.  ensure_backward() joins it."""
        if (self.has_backward or self._bwd_ready is not None or not PREFETCH_BACKWARD
                or self.B * self.N < max(1, PREFETCH_MIN_POINTS)):
            return self
        side = _side_stream(self.device)
        side.wait_event(self._searched)
        with torch.cuda.device(self.device), torch.cuda.stream(side):
            _lib.check(_lib.lib().conv3p_plan_build_backward(
                self.geom, _ptr(self.points), _ptr(self.buffer), self.buffer.numel(), C.c_void_p(side.cuda_stream)))
            self._bwd_ready = torch.cuda.Event()
            self._bwd_ready.record(side)
        self._side_events.append(self._bwd_ready)
        self.points.record_stream(side)
        return self

    def ensure_backward(self) -> "NeighborPlan":
        if self._bwd_ready is not None:
            torch.cuda.current_stream(self.device).wait_event(self._bwd_ready)
            self._bwd_ready = None
            self.has_backward = True
        if not self.has_backward:
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().conv3p_plan_build_backward(
                    self.geom, _ptr(self.points), _ptr(self.buffer), self.buffer.numel(),
                    _stream_ptr(self.device)))
            self.has_backward = True
        return self

    def matches(self, points: torch.Tensor, stride: tuple, voxel: float) -> bool:
        return (points.data_ptr() == self.points.data_ptr() and tuple(points.shape) == (self.B, self.N, 3)
                and stride == self.stride and voxel == self.voxel_size)


# --------------------------------------------------------------------------------------------------
# forward / backward on a plan
# --------------------------------------------------------------------------------------------------
ACTIVATIONS = {None: 0, "none": 0, "selu": 1}


def row_stride_of(shape, strides, B: int, N: int) -> Optional[int]:
    """Row stride (in elements) of a [B,N,C] tensor whose rows sit at a uniform distance inside a wider row-major
    buffer -- a channel slice ``buf[:, :, a:b]`` of a contiguous [B,N,W] tensor -- or None when the layout is anything
    else (the caller then makes it contiguous).  Pure function of shape and strides: tested on the CPU."""
    if len(shape) != 3 or shape[0] != B or shape[1] != N or B * N * shape[2] == 0:
        return None
    C = shape[2]
    sb, sn, sc = strides
    if (sc == 1 or C == 1) and sn >= C and (B == 1 or sb == N * sn):
        return int(sn)
    return None


def _rows_view(t: torch.Tensor, name: str, B: int, N: int):
    """-> (tensor, row stride in floats) for a [B,N,C] float32 CUDA tensor whose rows may sit inside a wider
    row-major buffer (a channel slice ``buf[:, :, a:b]`` of a contiguous [B,N,W] tensor); anything else is made
    contiguous."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"Conv3p: {name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"Conv3p: {name} must be a CUDA tensor (pointwise_b200 has no CPU fallback)")
    if t.dtype != torch.float32:
        raise TypeError(f"Conv3p: {name} must be float32 (got {t.dtype})")
    if t.dim() == 3:
        stride = row_stride_of(tuple(t.shape), t.stride(), B, N)
        if stride is not None:
            return t, stride
    t = t.contiguous()
    return t, int(t.shape[-1]) if t.dim() else 0


def conv3p_forward(plan: NeighborPlan, input: torch.Tensor, kernel: torch.Tensor,
                   activation: Optional[str] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Conv3p forward on a built plan.  ``activation="selu"`` fuses the SELU the reference's networks apply to every
    Conv3p output into the kernel's epilogue; ``input`` may be a channel slice of a wider [B,N,W] buffer and
    ``out`` (optional) such a slice to write into -- the concat of the 9-channel layers then needs no copy
    (scene_seg/pointcnn_scene_seg_acsd.py:35-36, :56)."""
    if activation not in ACTIVATIONS:
        raise ValueError(f"Conv3p: unknown activation {activation!r}")
    kernel = _check_cuda_f32(kernel, "kernel")
    Cin, Cout = int(kernel.shape[3]), int(kernel.shape[4])
    input, in_stride = _rows_view(input, "input", plan.B, plan.N)
    if out is None:
        out = torch.empty((plan.B, plan.N, Cout), dtype=torch.float32, device=plan.device)
        out_stride = Cout
    else:
        if tuple(out.shape) != (plan.B, plan.N, Cout):
            raise ValueError("Conv3p: out must have shape [B, N, Cout]")
        view, out_stride = _rows_view(out, "out", plan.B, plan.N)
        if view.data_ptr() != out.data_ptr():
            raise ValueError("Conv3p: out must be a channel slice of a contiguous [B, N, W] buffer")
    L = _lib.lib()
    nscratch = L.conv3p_scratch_bytes(plan.geom, Cin, Cout)
    scratch = torch.empty(nscratch, dtype=torch.uint8, device=plan.device)
    with torch.cuda.device(plan.device):
        _lib.check(L.conv3p_forward_ex_f32(plan.geom, _ptr(plan.buffer), _ptr(input), in_stride, _ptr(kernel),
                                           Cin, Cout, _ptr(out), out_stride, ACTIVATIONS[activation],
                                           _ptr(scratch), nscratch, _stream_ptr(plan.device)))
    return out


def selu_backward(y: torch.Tensor, grad: torch.Tensor) -> torch.Tensor:
    """grad * selu'(x) from the activated value y = selu(x) (the backward of the fused epilogue) -> dense [B,N,C]."""
    B, N, Cc = int(y.shape[0]), int(y.shape[1]), int(y.shape[2])
    y, ys = _rows_view(y, "y", B, N)
    grad, gs = _rows_view(grad, "grad", B, N)
    out = torch.empty((B, N, Cc), dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        _lib.check(_lib.lib().conv3p_selu_backward_f32(_ptr(y), ys, _ptr(grad), gs, _ptr(out), B * N, Cc,
                                                      _stream_ptr(y.device)))
    return out


# upper bound on the G store a backward call may allocate (3.6 GB at 64 x 4096 points, Cout = 128)
G_STORE_LIMIT_BYTES = int(os.environ.get("CONV3P_G_STORE_LIMIT_GB", "32")) << 30


def conv3p_backward(plan: NeighborPlan, grad_output: torch.Tensor, input: torch.Tensor,
                    kernel: torch.Tensor, need_input_grad: bool = True,
                    need_filter_grad: bool = True):
    grad_output = _check_cuda_f32(grad_output, "grad_output")
    input = _check_cuda_f32(input, "input")
    kernel = _check_cuda_f32(kernel, "kernel")
    Cin, Cout = int(kernel.shape[3]), int(kernel.shape[4])
    # shape checks of Conv3pGradOp::Compute, tf_conv3p_atrous.cpp:583-585
    if grad_output.shape[0] != plan.B:
        raise ValueError("backprop grad tensor has wrong size for dim 0")
    if grad_output.shape[1] != plan.N:
        raise ValueError("backprop grad tensor has wrong size for dim 1")
    if grad_output.shape[2] != Cout:
        raise ValueError("backprop grad tensor has wrong size for dim 2")
    # deferred overflow check of the plan: looked at if its counters have arrived (they have when backward runs after a
    # loss was computed); never waited for -- a caller that enqueues backward right behind forward (bench, the
    # host-buffer pipeline) would otherwise stall until the plan build has run.  A late flag is raised by the next
    # plan build, by HostConv3p.fetch() or by plan.verify().
    plan.verify(block=False)
    plan.ensure_backward()
    L = _lib.lib()
    gi = torch.empty((plan.B, plan.N, Cin), dtype=torch.float32, device=plan.device) \
        if need_input_grad else None
    gf = torch.empty_like(kernel) if need_filter_grad else None
    nscratch = L.conv3p_scratch_bytes(plan.geom, Cin, Cout)
    if need_input_grad and need_filter_grad:
        # room for the G store (27*Cout floats per point) lets the two gradient kernels share one gather
        nshared = L.conv3p_backward_scratch_bytes(plan.geom, Cin, Cout)
        if nshared - nscratch <= G_STORE_LIMIT_BYTES:
            nscratch = nshared
    try:
        scratch = torch.empty(nscratch, dtype=torch.uint8, device=plan.device)
    except torch.OutOfMemoryError:
        # no room for the G store: the smaller scratch gives identical results (each gradient kernel gathers)
        nscratch = L.conv3p_scratch_bytes(plan.geom, Cin, Cout)
        scratch = torch.empty(nscratch, dtype=torch.uint8, device=plan.device)
    with torch.cuda.device(plan.device):
        _lib.check(L.conv3p_backward_f32(plan.geom, _ptr(plan.buffer), _ptr(grad_output), _ptr(input),
                                         _ptr(kernel), Cin, Cout, _ptr(gi), _ptr(gf), _ptr(scratch),
                                         nscratch, _stream_ptr(plan.device)))
    return gi, gf


class _Conv3pFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, input, kernel, plan, activation=None):
        ctx.plan = plan
        ctx.activation = activation
        out = conv3p_forward(plan, input, kernel, activation=activation)
        if any(ctx.needs_input_grad):
            plan.prefetch_backward()     # overlaps the rest of the forward pass
        if activation == "selu":
            ctx.save_for_backward(input, kernel, out)
        else:
            ctx.save_for_backward(input, kernel)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.activation == "selu":
            input, kernel, out = ctx.saved_tensors
            grad_output = selu_backward(out, grad_output)
        else:
            input, kernel = ctx.saved_tensors
        gi, gf = conv3p_backward(ctx.plan, grad_output, input, kernel,
                                 need_input_grad=ctx.needs_input_grad[1],
                                 need_filter_grad=ctx.needs_input_grad[2])
        # the reference returns [None, input_grad, filter_grad, None, None] (pointcnn2_acsd.py:31)
        return None, gi, gf, None, None


# --------------------------------------------------------------------------------------------------
# general filter shapes (anything but 3x3x3): one-shot calls of the general fp32 path
# --------------------------------------------------------------------------------------------------
_generic_capacity_hint: dict = {}


def _generic_call(points, stride, voxel, dims, Cin, Cout, backward, run, f64=False):
    """Runs a one-shot C entry point on a workspace sized for `capacity` pairs; grows the capacity and retries when
    the neighbour lists overflowed (one 128-byte read-back per call: this path is the reference-compatible fallback
    for filter shapes no model uses, not the tuned one)."""
    L = _lib.lib()
    B, N = int(points.shape[0]), int(points.shape[1])
    key = (B, N, stride, voxel, dims, f64)
    vol = dims[0] * dims[1] * dims[2]
    cap = _generic_capacity_hint.get(key, max(1024, 2 * vol * B * N))
    dims_c = (C.c_int * 3)(*dims)
    stride_c = (C.c_int * 3)(*stride)
    while True:
        geom = _lib.make_geom(B, N, stride, voxel, cap)
        if f64:
            nbytes = L.conv3p_op_workspace_bytes_f64(geom, dims_c, Cin, Cout)
        else:
            nbytes = L.conv3p_op_workspace_bytes_ex(geom, dims_c, Cin, Cout, 1 if backward else 0)
        if nbytes == 0:
            raise _lib.Conv3pError(_lib.ERR_UNSUPPORTED, f"filter shape {dims} is not supported (more than 512 cells)")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=points.device)
        with torch.cuda.device(points.device):
            _lib.check(run(L, dims_c, stride_c, cap, ws, nbytes))
            st = _lib.PlanStats()
            _lib.check(L.conv3p_plan_stats(geom, _ptr(ws), st, _stream_ptr(points.device)))
        if not st.overflow:
            _generic_capacity_hint[key] = max(_generic_capacity_hint.get(key, 0), int(st.total_pairs * 1.25) + 1024)
            return
        cap = int(st.total_pairs * 1.125) + 1024


class _Conv3pGenericFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, input, kernel, stride, voxel):
        dims = tuple(int(d) for d in kernel.shape[:3])
        Cin, Cout = int(kernel.shape[3]), int(kernel.shape[4])
        B, N = int(points.shape[0]), int(points.shape[1])
        f64 = points.dtype == torch.float64
        out = torch.empty((B, N, Cout), dtype=points.dtype, device=points.device)
        _generic_call(points, stride, voxel, dims, Cin, Cout, False,
                      lambda L, d, s, cap, ws, nb: (L.conv3p_op_forward_f64 if f64 else L.conv3p_op_forward_f32)(
                          _ptr(points), _ptr(input), _ptr(kernel), d, s, voxel, B, N, Cin, Cout, cap, _ptr(out),
                          _ptr(ws), nb, _stream_ptr(points.device)), f64=f64)
        ctx.save_for_backward(points, input, kernel)
        ctx.geometry = (stride, voxel)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        points, input, kernel = ctx.saved_tensors
        stride, voxel = ctx.geometry
        f64 = points.dtype == torch.float64
        grad_output = _check_cuda_f32(grad_output, "grad_output", points.dtype)
        dims = tuple(int(d) for d in kernel.shape[:3])
        Cin, Cout = int(kernel.shape[3]), int(kernel.shape[4])
        B, N = int(points.shape[0]), int(points.shape[1])
        gi = torch.empty_like(input)
        gf = torch.empty_like(kernel)
        _generic_call(points, stride, voxel, dims, Cin, Cout, True,
                      lambda L, d, s, cap, ws, nb: (L.conv3p_op_backward_f64 if f64 else L.conv3p_op_backward_f32)(
                          _ptr(grad_output), _ptr(points), _ptr(input), _ptr(kernel), d, s, voxel, B, N, Cin, Cout, cap,
                          _ptr(gi), _ptr(gf), _ptr(ws), nb, _stream_ptr(points.device)), f64=f64)
        return None, gi, gf, None, None


def conv3p(points_tensor, input_tensor, kernel_tensor, stride, voxel_size,
           plan: Optional[NeighborPlan] = None, activation: Optional[str] = None) -> torch.Tensor:
    """Drop-in for the reference's ``conv3p`` (pointcnn2_acsd.py:12-13): same positional signature.

    points [B,N,3], input [B,N,Cin], kernel [fz,fy,fx,Cin,Cout] (z,y,x,in,out; 3x3x3 in every reference model and on
    the tuned engines, any shape up to 512 cells on the general path), stride = 3 ints (x,y,z),
    voxel_size = 1 float; returns [B,N,Cout].  Differentiable w.r.t. input and kernel only
    (pointcnn2_acsd.py:31).  float32 tensors, or float64 throughout (the reference's T = double registration,
    register_op.cpp:45: predicate and sums in double, general path).  ``plan`` optionally reuses a NeighborPlan built
    for the same points,
    stride and voxel size (e.g. across layers).  ``activation="selu"`` returns ``selu(conv3p(...))`` with the
    activation fused into the kernel's epilogue (and its derivative applied to the incoming gradient in backward).
    """
    T = _op_dtype(points_tensor)
    points = _check_cuda_f32(points_tensor, "points", T)
    input = _check_cuda_f32(input_tensor, "input", T)
    kernel = _check_cuda_f32(kernel_tensor, "kernel", T)
    validate(points, input, kernel)
    s, v = parse_stride(stride), parse_voxel(voxel_size)
    if activation not in ACTIVATIONS:
        raise ValueError(f"Conv3p: unknown activation {activation!r}")
    if tuple(kernel.shape[:3]) != (3, 3, 3) or T == torch.float64:
        # the reference is generic in the filter shape (tf_conv3p_atrous.cpp:425-427) and registered for double as well
        # (register_op.cpp:45); anything but float32 3x3x3 takes the general path (no plan reuse, no fused epilogue)
        if plan is not None:
            raise ValueError("Conv3p: a NeighborPlan serves float32 3x3x3 filters only")
        y = _Conv3pGenericFunction.apply(points, input, kernel, s, v)
        return torch.nn.functional.selu(y) if activation == "selu" else y
    if plan is None:
        plan = NeighborPlan(points, s, v)
    elif not plan.matches(points, s, v):
        raise ValueError("Conv3p: the supplied NeighborPlan was built for different points/stride/voxel_size")
    return _Conv3pFunction.apply(points, input, kernel, plan, None if activation in (None, "none") else activation)


def conv3p_grad(grad_from_next, points, input, filter, stride, voxel_size,
                plan: Optional[NeighborPlan] = None):
    """Mirror of the reference's ``conv3p_grad`` op (register_op.cpp:63-75):
    -> (input_grad, filter_grad)."""
    T = _op_dtype(points)
    points = _check_cuda_f32(points, "points", T)
    input = _check_cuda_f32(input, "input", T)
    filter = _check_cuda_f32(filter, "filter", T)
    validate(points, input, filter)
    s, v = parse_stride(stride), parse_voxel(voxel_size)
    if tuple(filter.shape[:3]) != (3, 3, 3) or T == torch.float64:
        class _Ctx:      # the autograd node's backward, called directly
            pass
        ctx = _Ctx()
        ctx.saved_tensors = (points, input, filter)
        ctx.geometry = (s, v)
        _, gi, gf, _, _ = _Conv3pGenericFunction.backward(ctx, grad_from_next)
        return gi, gf
    if plan is None:
        plan = NeighborPlan(points, s, v)
    return conv3p_backward(plan, grad_from_next, input, filter)


def set_engine(engine: str = "auto") -> str:
    """Selects the contraction engine: "auto" (tensor cores where the shape allows, else SIMT),
    "simt" (fp32 CUDA cores only: warp-per-point or tile kernels by channel count), "tc", or "tile" (the generic fp32
    tile kernels only, no tensor cores).  Returns the previous setting."""
    names = ["auto", "simt", "tc", "tile"]
    prev = _lib.lib().conv3p_set_engine(names.index(engine))
    return names[prev & 7] if (prev & 7) < len(names) else "auto"   # higher bits are ablation flags


def launch_count(reset: bool = False) -> int:
    """Kernels launched by the library on this thread since the last reset."""
    return int(_lib.lib().conv3p_launch_count(1 if reset else 0))
