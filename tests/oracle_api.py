"""TEST INFRASTRUCTURE: ctypes entry points of the CPU oracle (oracle/restate.cpp -> oracle/liboracle795.so) and readers
for the dumps of the unmodified reference binaries (oracle/_ref/ref_raytracer, ref_hits).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.  Nothing
in the product package (raytracer-795_b200/) knows about the oracle: the product path has no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle795.so")
ORACLE_REF_DIR = os.path.join(ROOT, "oracle", "_ref")

_oracle = None


def _rt():
    import sys
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    return ge.load_package()


def load_oracle() -> C.CDLL:
    """The CPU restatement (oracle/restate.cpp).  Never used by the product path."""
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            raise RuntimeError(f"{ORACLE_SO} is missing: run `make -C oracle restate`")
        rt = _rt()
        lib = C.CDLL(ORACLE_SO)
        P = C.POINTER
        lib.oracle_render.argtypes = [P(rt.RtSceneDesc), P(rt.RtCameraDesc), P(rt.RtRenderOpts), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.oracle_trace_primary.argtypes = [P(rt.RtSceneDesc), P(rt.RtCameraDesc), C.c_int, C.c_void_p]
        lib.oracle_ray_queries.argtypes = [C.c_int]
        lib.oracle_ray_queries.restype = C.c_uint64
        lib.oracle_perlin.argtypes = [C.c_float] * 4 + [C.c_int]
        lib.oracle_perlin.restype = C.c_float
        lib.oracle_philox.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        lib.oracle_conductor_fresnel.argtypes = [C.c_float] * 3
        lib.oracle_conductor_fresnel.restype = C.c_float
        lib.oracle_tonemap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
        _oracle = lib
    return _oracle


def oracle_render(desc, cam, opts=None, rows=(0, 0), cols=(0, 0), threads: int = 8) -> np.ndarray:
    lib = load_oracle()
    opts = opts or _rt().default_opts()
    out = np.zeros((cam.ny, cam.nx, 3), dtype=np.float32)
    lib.oracle_render(desc, C.byref(cam), C.byref(opts), rows[0], rows[1], cols[0], cols[1], threads, out.ctypes.data)
    return out


def oracle_trace_primary(desc, cam, threads: int = 8) -> np.ndarray:
    lib = load_oracle()
    out = np.zeros((cam.ny, cam.nx), dtype=_rt().HIT_DTYPE)
    lib.oracle_trace_primary(desc, C.byref(cam), threads, out.ctypes.data)
    return out


def oracle_tonemap(rgb: np.ndarray, key: float = 0.18, burn: float = 1.0, saturation: float = 1.0, gamma: float = 2.2) -> np.ndarray:
    ny, nx, _ = rgb.shape
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    out = np.empty((ny, nx, 3), dtype=np.uint8)
    load_oracle().oracle_tonemap(rgb.ctypes.data, nx, ny, key, burn, saturation, gamma, out.ctypes.data)
    return out


REF_HIT_DTYPE = np.dtype([("full", "<i4"), ("label", "<i4"), ("t", "<f4"), ("point", "<f4", (3,)), ("normal", "<f4", (3,))])


def read_ref_hits(path: str) -> np.ndarray:
    raw = np.fromfile(path, dtype=np.uint8)
    nx, ny = np.frombuffer(raw[:8], dtype="<i4")
    return np.frombuffer(raw[8:], dtype=REF_HIT_DTYPE).reshape(ny, nx).copy()
