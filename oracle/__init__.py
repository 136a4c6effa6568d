"""oracle -- TEST INFRASTRUCTURE for the Conv3p hot path.  Not product code.

Two CPU checkers, both reached through ctypes:

* ``port``  -- ``oracle/libconv3p_oracle.so``: the plain-C restatement in ``conv3p_oracle.c``
  (each function cites the reference file:line it follows).
* ``ref``   -- ``oracle/_ref/libconv3p_ref.so``: the reference's own CPU op
  (``/root/reference/tf_ops/conv3p/tf_conv3p_atrous.cpp``) compiled UNMODIFIED against
  ``oracle/tf_shim`` by ``oracle/Makefile``.  Present wherever it was built (the build container;
  it travels to the GPU box as a prebuilt library).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package.  ``pointwise_b200`` never does: the product has no CPU
fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "libconv3p_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libconv3p_ref.so")
REF_ST_SO = os.path.join(_HERE, "_ref", "libconv3p_ref_st.so")
REF_GPU_SO = os.path.join(_HERE, "_ref", "libconv3p_ref_gpu.so")
NCELL = 27

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def build(verbose: bool = False) -> None:
    """Compile the checkers (``make -C oracle all``).  Building the checker is not using it."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout + out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _stride3(stride):
    s = np.ascontiguousarray(np.broadcast_to(np.asarray(stride, dtype=np.int32), (3,)))
    return s


def _opt64(shape, want):
    return np.zeros(shape, np.float64) if want else None


def _ptr64(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class Port:
    """The C restatement (``conv3p_oracle.c``)."""

    kind = "port"

    def __init__(self, path: str = PORT_SO):
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        L.oracle_neighbor_count_f32.argtypes = [_f32p, C.c_int, _i32p, C.c_float, _i32p]
        L.oracle_neighbor_count_bruteforce_f32.argtypes = [_f32p, C.c_int, _i32p, C.c_float, _i32p]
        L.oracle_neighbors_f32.argtypes = [_f32p, C.c_int, _i32p, C.c_float, _i64p, _i32p, _i32p,
                                           C.c_longlong]
        L.oracle_neighbors_f32.restype = C.c_longlong
        L.oracle_conv3p_forward_f32.argtypes = [_f32p, _f32p, _f32p, _i32p, C.c_float, C.c_int,
                                                C.c_int, C.c_int, C.c_int, _f32p, dp, dp]
        L.oracle_conv3p_backward_f32.argtypes = [_f32p, _f32p, _f32p, _f32p, _i32p, C.c_float,
                                                 C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p,
                                                 dp, dp, dp, dp]
        L.oracle_num_threads.restype = C.c_int

    @property
    def threads(self) -> int:
        return int(self.lib.oracle_num_threads())

    def neighbor_count(self, xyz, stride, voxel, bruteforce: bool = False):
        xyz = _f32(xyz).reshape(-1, 3)
        n = xyz.shape[0]
        out = np.zeros((n, NCELL), np.int32)
        fn = (self.lib.oracle_neighbor_count_bruteforce_f32 if bruteforce
              else self.lib.oracle_neighbor_count_f32)
        rc = fn(xyz, n, _stride3(stride), float(np.float32(voxel)), out)
        assert rc == 0
        return out

    def neighbors(self, xyz, stride, voxel):
        """-> (off[n+1] int64, j[total] int32, f[total] int32) in reference emission order."""
        xyz = _f32(xyz).reshape(-1, 3)
        n = xyz.shape[0]
        off = np.zeros(n + 1, np.int64)
        cap = max(1, n) * 64
        while True:
            j = np.zeros(cap, np.int32)
            f = np.zeros(cap, np.int32)
            total = self.lib.oracle_neighbors_f32(xyz, n, _stride3(stride),
                                                  float(np.float32(voxel)), off, j, f, cap)
            assert total >= 0
            if total <= cap:
                return off, j[:total].copy(), f[:total].copy()
            cap = int(total)

    def forward(self, points, input, filter, stride, voxel, with64: bool = False):
        points, input, filter = _f32(points), _f32(input), _f32(filter)
        B, N, _ = points.shape
        Cin, Cout = filter.shape[-2], filter.shape[-1]
        assert filter.shape[:3] == (3, 3, 3) and input.shape == (B, N, Cin)
        out = np.zeros((B, N, Cout), np.float32)
        a64, s64 = _opt64(out.shape, with64), _opt64(out.shape, with64)
        rc = self.lib.oracle_conv3p_forward_f32(points, input, filter, _stride3(stride),
                                                float(np.float32(voxel)), B, N, Cin, Cout, out,
                                                _ptr64(a64), _ptr64(s64))
        assert rc == 0
        return (out, a64, s64) if with64 else out

    def backward(self, grad_out, points, input, filter, stride, voxel, with64: bool = False):
        grad_out, points, input, filter = _f32(grad_out), _f32(points), _f32(input), _f32(filter)
        B, N, _ = points.shape
        Cin, Cout = filter.shape[-2], filter.shape[-1]
        assert grad_out.shape == (B, N, Cout)
        gi = np.zeros((B, N, Cin), np.float32)
        gf = np.zeros(filter.shape, np.float32)
        e = [_opt64(gi.shape, with64), _opt64(gi.shape, with64),
             _opt64(gf.shape, with64), _opt64(gf.shape, with64)]
        rc = self.lib.oracle_conv3p_backward_f32(grad_out, points, input, filter, _stride3(stride),
                                                 float(np.float32(voxel)), B, N, Cin, Cout, gi, gf,
                                                 *[_ptr64(a) for a in e])
        assert rc == 0
        return (gi, gf, *e) if with64 else (gi, gf)


class Ref:
    """The reference's own CPU op, compiled unmodified (``oracle/_ref``)."""

    kind = "reference"

    def __init__(self, single_thread: bool = False):
        path = REF_ST_SO if single_thread else REF_SO
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path)
        L.ref_last_error.restype = C.c_char_p
        L.ref_num_threads.restype = C.c_int
        L.ref_hardware_concurrency.restype = C.c_int
        L.ref_conv3p_forward_f32.argtypes = [_f32p, _f32p, _f32p, _i32p, C.c_int, _f32p, C.c_int] + \
            [C.c_int] * 11 + [_f32p]
        L.ref_conv3p_backward_f32.argtypes = [_f32p, _f32p, _f32p, _f32p, _i32p, _f32p] + \
            [C.c_int] * 10 + [_f32p, _f32p]
        L.ref_neighbor_count_f32.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p,
                                             C.c_float, _i32p]
        L.ref_neighbor_count_f32.restype = None
        L.ref_neighbors_f32.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, C.c_float,
                                        _i64p, _i32p, _i32p, C.c_longlong]
        L.ref_neighbors_f32.restype = C.c_longlong
        if hasattr(L, "ref_conv3p_forward_f64"):
            L.ref_conv3p_forward_f64.argtypes = [_f64p, _f64p, _f64p, _i32p, _f64p] + [C.c_int] * 7 + [_f64p]
            L.ref_conv3p_backward_f64.argtypes = [_f64p, _f64p, _f64p, _f64p, _i32p, _f64p] + [C.c_int] * 7 + \
                [_f64p, _f64p]
            L.ref_neighbor_count_f64.argtypes = [_f64p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, C.c_double, _i32p]
            L.ref_neighbor_count_f64.restype = None

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    @property
    def threads(self) -> int:
        return int(self.lib.ref_num_threads())

    def _err(self):
        return self.lib.ref_last_error().decode()

    def forward(self, points, input, filter, stride, voxel, *, raw_stride=None, raw_voxel=None,
                input_shape=None, filter_cin=None, points_rank3=True):
        """Runs Conv3p/CPU/float.  The keyword arguments let tests provoke the reference's own
        OP_REQUIRES failures (a ValueError carrying the reference's message is raised)."""
        points, input, filter = _f32(points), _f32(input), _f32(filter)
        B, N = points.shape[0], points.shape[1]
        fz, fy, fx, Cin, Cout = filter.shape
        st = np.ascontiguousarray(raw_stride, np.int32) if raw_stride is not None else _stride3(stride)
        vx = np.ascontiguousarray(raw_voxel, np.float32) if raw_voxel is not None else \
            np.array([voxel], np.float32)
        iB, iN = (input_shape if input_shape is not None else input.shape[:2])
        out = np.zeros((B, N, Cout), np.float32)
        rc = self.lib.ref_conv3p_forward_f32(points, input, filter, st, st.size, vx, vx.size,
                                             B, N, input.shape[-1], Cout, fz, fy, fx, iB, iN,
                                             Cin if filter_cin is None else filter_cin,
                                             1 if points_rank3 else 0, out)
        if rc:
            raise ValueError(self._err())
        return out

    def backward(self, grad_out, points, input, filter, stride, voxel, *, grad_shape=None):
        grad_out, points, input, filter = _f32(grad_out), _f32(points), _f32(input), _f32(filter)
        B, N = points.shape[0], points.shape[1]
        fz, fy, fx, Cin, Cout = filter.shape
        gB, gN, gC = grad_shape if grad_shape is not None else grad_out.shape
        gi = np.zeros((B, N, Cin), np.float32)
        gf = np.zeros(filter.shape, np.float32)
        rc = self.lib.ref_conv3p_backward_f32(grad_out, points, input, filter, _stride3(stride),
                                              np.array([voxel], np.float32), B, N, Cin, Cout,
                                              fz, fy, fx, gB, gN, gC, gi, gf)
        if rc:
            raise ValueError(self._err())
        return gi, gf

    def neighbor_count(self, xyz, stride, voxel, dims=(3, 3, 3)):
        """dims = (fz, fy, fx): the reference is generic in the filter shape (tf_conv3p_atrous.cpp:425-427)."""
        xyz = _f32(xyz).reshape(-1, 3)
        n = xyz.shape[0]
        fz, fy, fx = (int(d) for d in dims)
        out = np.zeros((n, fz * fy * fx), np.int32)
        self.lib.ref_neighbor_count_f32(xyz, n, fz, fy, fx, _stride3(stride), float(np.float32(voxel)), out)
        return out

    # ---- T = double (register_op.cpp:45, 64; tf_conv3p_atrous.cpp:516, 727) ---------------------------------
    def forward64(self, points, input, filter, stride, voxel):
        points, input, filter = _f64(points), _f64(input), _f64(filter)
        B, N = points.shape[0], points.shape[1]
        fz, fy, fx, Cin, Cout = filter.shape
        out = np.zeros((B, N, Cout), np.float64)
        rc = self.lib.ref_conv3p_forward_f64(points, input, filter, _stride3(stride), np.array([voxel], np.float64),
                                             B, N, Cin, Cout, fz, fy, fx, out)
        if rc:
            raise ValueError(self._err())
        return out

    def backward64(self, grad_out, points, input, filter, stride, voxel):
        grad_out, points, input, filter = _f64(grad_out), _f64(points), _f64(input), _f64(filter)
        B, N = points.shape[0], points.shape[1]
        fz, fy, fx, Cin, Cout = filter.shape
        gi = np.zeros((B, N, Cin), np.float64)
        gf = np.zeros(filter.shape, np.float64)
        rc = self.lib.ref_conv3p_backward_f64(grad_out, points, input, filter, _stride3(stride),
                                              np.array([voxel], np.float64), B, N, Cin, Cout, fz, fy, fx, gi, gf)
        if rc:
            raise ValueError(self._err())
        return gi, gf

    def neighbor_count64(self, xyz, stride, voxel, dims=(3, 3, 3)):
        xyz = _f64(xyz).reshape(-1, 3)
        n = xyz.shape[0]
        fz, fy, fx = (int(d) for d in dims)
        out = np.zeros((n, fz * fy * fx), np.int32)
        self.lib.ref_neighbor_count_f64(xyz, n, fz, fy, fx, _stride3(stride), float(voxel), out)
        return out

    def neighbors(self, xyz, stride, voxel):
        xyz = _f32(xyz).reshape(-1, 3)
        n = xyz.shape[0]
        off = np.zeros(n + 1, np.int64)
        cap = max(1, n) * 64
        while True:
            j = np.zeros(cap, np.int32)
            f = np.zeros(cap, np.int32)
            total = self.lib.ref_neighbors_f32(xyz, n, 3, 3, 3, _stride3(stride),
                                               float(np.float32(voxel)), off, j, f, cap)
            if total <= cap:
                return off, j[:total].copy(), f[:total].copy()
            cap = int(total)


class RefGpu:
    """The reference's own GPU op (tf_conv3p_atrous.cu compiled unmodified for sm_100a, ``make -C oracle refgpu``):
    a TIMING baseline for bench.py ("the reference's GPU code on the same box").  Built with the reference's
    -use_fast_math, so it is not a parity reference.  Takes CUDA torch tensors."""

    kind = "reference-gpu"

    def __init__(self):
        if not os.path.exists(REF_GPU_SO):
            raise FileNotFoundError(REF_GPU_SO)
        self.lib = L = C.CDLL(REF_GPU_SO)
        vp = C.c_void_p
        L.refgpu_conv3p_forward_f32.argtypes = [vp] * 5 + [C.c_int] * 4 + [vp]
        L.refgpu_conv3p_backward_f32.argtypes = [vp] * 6 + [C.c_int] * 4 + [vp, vp]
        L.refgpu_last_error.restype = C.c_char_p

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_GPU_SO)

    def forward(self, points, input, filter, stride_dev, voxel_dev, out):
        B, N, Cin = input.shape
        rc = self.lib.refgpu_conv3p_forward_f32(points.data_ptr(), input.data_ptr(), filter.data_ptr(),
                                                stride_dev.data_ptr(), voxel_dev.data_ptr(), B, N, Cin,
                                                filter.shape[-1], out.data_ptr())
        if rc:
            raise RuntimeError(f"reference GPU op failed ({rc}): {self.lib.refgpu_last_error().decode()}")
        return out

    def backward(self, grad_out, points, input, filter, stride_dev, voxel_dev, grad_input, grad_filter):
        B, N, Cin = input.shape
        rc = self.lib.refgpu_conv3p_backward_f32(grad_out.data_ptr(), points.data_ptr(), input.data_ptr(),
                                                 filter.data_ptr(), stride_dev.data_ptr(), voxel_dev.data_ptr(), B, N,
                                                 Cin, filter.shape[-1], grad_input.data_ptr(), grad_filter.data_ptr())
        if rc:
            raise RuntimeError(f"reference GPU op failed ({rc}): {self.lib.refgpu_last_error().decode()}")
        return grad_input, grad_filter


_port = None
_ref = None


def port() -> Port:
    global _port
    if _port is None:
        _port = Port()
    return _port


def ref() -> Ref:
    global _ref
    if _ref is None:
        _ref = Ref()
    return _ref


def best():
    """The strongest checker available: the reference's own object code, else the port."""
    return ref() if Ref.available() else port()
