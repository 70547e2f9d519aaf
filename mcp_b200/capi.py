"""ctypes binding of libmcpb200.so — a 1:1 mirror of `include/mcpb200.h` (the same calls the Julia
`ccall` wrapper in INTEGRATION.md makes).  No numerical logic lives here.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import build as _build

OK = 0
ERR_INVALID_ARGUMENT = -1
ERR_UNSUPPORTED = -2
ERR_COMPILE = -3
ERR_CUDA = -4
ERR_NO_SENSITIVITIES = -5
ERR_INTERNAL = -6

COMPILE_ONLY = 1
NO_CACHE = 2

EXPORTS = [
    "mcpb200_create", "mcpb200_destroy", "mcpb200_last_error", "mcpb200_global_error", "mcpb200_get_info",
    "mcpb200_get_timing", "mcpb200_get_source", "mcpb200_default_opts", "mcpb200_set_devices",
    "mcpb200_solve_batched", "mcpb200_solve_batched_device", "mcpb200_sensitivities",
    "mcpb200_sensitivities_device", "mcpb200_measure_fp64_peak", "mcpb200_flush_l2",
]

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class ProblemDesc(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("ntheta", C.c_int32), ("n_nodes", C.c_int32),
        ("op", _i32p), ("a", _i32p), ("b", _i32p),
        ("n_consts", C.c_int32), ("consts", _f64p), ("gh_nodes", _i32p),
        ("jz_nnz", C.c_int32), ("jz_rows", _i32p), ("jz_cols", _i32p), ("jz_nodes", _i32p),
        ("jt_nnz", C.c_int32), ("jt_rows", _i32p), ("jt_cols", _i32p), ("jt_nodes", _i32p),
    ]


class SolverOpts(C.Structure):
    _fields_ = [("tol", C.c_double), ("max_inner_iters", C.c_int32), ("max_outer_iters", C.c_int32),
                ("tightening_rate", C.c_double), ("loosening_rate", C.c_double), ("min_stepsize", C.c_double)]


class Info(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "nx", "ny", "ntheta", "n_reduced", "kl", "ku", "window_rows", "window_cols", "row_stride",
        "n_jac_computed", "n_jac_constant", "n_assembly_dests", "n_assembly_terms", "threads_per_instance",
        "instances_per_cta", "ctas_per_sm", "smem_bytes_per_cta", "regs_solve", "regs_sens",
        "has_sensitivities", "cache_hit")] + [("flops_per_newton_step_band", C.c_double)]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Timing(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
                ("launches", C.c_int64), ("newton_steps", C.c_int64), ("solved", C.c_int64),
                ("pass0_ms", C.c_double), ("deferred", C.c_int64)]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class MCPB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libmcpb200 error {code}: {msg}")
        self.code = code


_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """Builds (if stale) and loads the in-tree libmcpb200.so.  There is no fallback: if the CUDA
    toolchain is missing this raises."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if os.environ.get("MCPB200_NO_AUTOBUILD") != "1" or not os.path.exists(path):
        path = _build.build_library()
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.mcpb200_create.argtypes = [C.POINTER(ProblemDesc), C.c_uint32, C.POINTER(vp)]
    lib.mcpb200_destroy.argtypes = [vp]
    lib.mcpb200_last_error.argtypes = [vp]
    lib.mcpb200_last_error.restype = C.c_char_p
    lib.mcpb200_global_error.restype = C.c_char_p
    lib.mcpb200_get_info.argtypes = [vp, C.POINTER(Info)]
    lib.mcpb200_get_timing.argtypes = [vp, C.POINTER(Timing)]
    lib.mcpb200_get_source.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_int64)]
    lib.mcpb200_default_opts.argtypes = [C.POINTER(SolverOpts)]
    lib.mcpb200_default_opts.restype = None
    lib.mcpb200_set_devices.argtypes = [vp, _i32p, C.c_int32]
    solve_args = [vp, C.c_int64, vp, vp, vp, vp, C.POINTER(SolverOpts), vp, vp, vp, vp, vp, vp, vp, vp]
    lib.mcpb200_solve_batched.argtypes = solve_args
    lib.mcpb200_solve_batched_device.argtypes = solve_args + [vp]
    sens_args = [vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int32, vp, vp, vp]
    lib.mcpb200_sensitivities.argtypes = sens_args
    lib.mcpb200_sensitivities_device.argtypes = sens_args + [vp]
    lib.mcpb200_measure_fp64_peak.argtypes = [_f64p]
    lib.mcpb200_flush_l2.argtypes = [vp]
    _lib = lib
    return lib


def _i32(arr):
    arr = np.ascontiguousarray(arr, dtype=np.int32)
    return arr, arr.ctypes.data_as(_i32p)


class Handle:
    """Owns one `mcpb200_handle`."""

    def __init__(self, ir, flags: int = 0):
        lib = load_library()
        self._lib = lib
        keep = []
        d = ProblemDesc()
        d.nx, d.ny, d.ntheta, d.n_nodes = ir.nx, ir.ny, ir.ntheta, len(ir.op)
        for name in ("op", "a", "b", "gh_nodes", "jz_rows", "jz_cols", "jz_nodes"):
            arr, ptr = _i32(getattr(ir, name))
            keep.append(arr)
            setattr(d, name, ptr)
        consts = np.ascontiguousarray(ir.consts, dtype=np.float64)
        keep.append(consts)
        d.n_consts, d.consts = len(consts), consts.ctypes.data_as(_f64p)
        d.jz_nnz = len(ir.jz_rows)
        if ir.jt_nodes is None:
            d.jt_nnz = -1
        else:
            d.jt_nnz = len(ir.jt_rows)
            for name in ("jt_rows", "jt_cols", "jt_nodes"):
                arr, ptr = _i32(getattr(ir, name))
                keep.append(arr)
                setattr(d, name, ptr)
        h = C.c_void_p()
        rc = lib.mcpb200_create(C.byref(d), flags, C.byref(h))
        if rc != OK:
            raise MCPB200Error(rc, lib.mcpb200_global_error().decode("utf-8", "replace"))
        self._h = h
        self.nx, self.ny, self.ntheta = ir.nx, ir.ny, ir.ntheta

    # -- helpers --------------------------------------------------------------------------------
    def check(self, rc: int):
        if rc != OK:
            raise MCPB200Error(rc, self._lib.mcpb200_last_error(self._h).decode("utf-8", "replace"))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mcpb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def raw(self):
        return self._h

    def info(self) -> dict:
        i = Info()
        self.check(self._lib.mcpb200_get_info(self._h, C.byref(i)))
        return i.asdict()

    def timing(self) -> dict:
        t = Timing()
        self.check(self._lib.mcpb200_get_timing(self._h, C.byref(t)))
        return t.asdict()

    def source(self) -> str:
        p, n = C.c_char_p(), C.c_int64()
        self.check(self._lib.mcpb200_get_source(self._h, C.byref(p), C.byref(n)))
        return p.value.decode("utf-8")

    def set_devices(self, ids):
        arr, ptr = _i32(list(ids))
        self.check(self._lib.mcpb200_set_devices(self._h, ptr, len(arr)))


def default_opts(**overrides) -> SolverOpts:
    o = SolverOpts()
    load_library().mcpb200_default_opts(C.byref(o))
    for k, v in overrides.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown solver option {k!r}")
        setattr(o, k, v)
    return o


def measure_fp64_peak() -> float:
    v = C.c_double()
    lib = load_library()
    rc = lib.mcpb200_measure_fp64_peak(C.byref(v))
    if rc != OK:
        raise MCPB200Error(rc, lib.mcpb200_global_error().decode())
    return v.value


def flush_l2(stream: int = 0):
    lib = load_library()
    rc = lib.mcpb200_flush_l2(C.c_void_p(stream))
    if rc != OK:
        raise MCPB200Error(rc, lib.mcpb200_global_error().decode())
