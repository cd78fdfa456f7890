"""ctypes front-end of oracle/cellflow_oracle.c.  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs.  The product package (cellflow_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")

MAX_TYPES = 10
MAX_CONN = 16

PARTICLE = np.dtype(
    [("pos", "<f4", 3), ("vel", "<f4", 3), ("acc", "<f4", 3), ("ptype", "<u4"), ("pad", "<f4")]
)
assert PARTICLE.itemsize == 44
EDGE = np.dtype([("i", "<i4"), ("j", "<i4")])
COLOR = np.dtype([("r", "<f4"), ("g", "<f4"), ("b", "<f4")])


class Params(C.Structure):
    """Mirror of cf_params (include/cellflow_b200.h); defaults = SimulationParams.h:15-37."""

    _fields_ = [
        ("radius", C.c_float),
        ("delta_t", C.c_float),
        ("friction", C.c_float),
        ("repulsion", C.c_float),
        ("attraction", C.c_float),
        ("k", C.c_float),
        ("balance", C.c_float),
        ("canvasWidth", C.c_float),
        ("canvasHeight", C.c_float),
        ("canvasDepth", C.c_float),
        ("spawnRegionSize", C.c_float),
        ("numParticleTypes", C.c_int32),
        ("ratioWithLFO", C.c_float),
        ("forceMultiplier", C.c_float),
        ("maxExpectedNeighbors", C.c_int32),
        ("forceRange", C.c_float),
        ("forceBias", C.c_float),
        ("ratio", C.c_float),
        ("lfoA", C.c_float),
        ("lfoS", C.c_float),
        ("forceOffset", C.c_float),
    ]

    DEFAULTS = dict(
        radius=42.07, delta_t=0.18, friction=0.51, repulsion=64.83, attraction=3.06, k=29.45,
        balance=0.79, canvasWidth=8000.0, canvasHeight=8000.0, canvasDepth=8000.0,
        spawnRegionSize=2000.0, numParticleTypes=6, ratioWithLFO=0.0, forceMultiplier=2.33,
        maxExpectedNeighbors=400, forceRange=0.28, forceBias=-0.20, ratio=0.0, lfoA=0.0,
        lfoS=0.1, forceOffset=1.0,
    )

    def __init__(self, **kw):
        super().__init__()
        vals = dict(self.DEFAULTS)
        vals.update(kw)
        for k, v in vals.items():
            setattr(self, k, v)

    def copy(self, **kw):
        p = Params(**{f: getattr(self, f) for f, _ in self._fields_})
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    @property
    def canvas(self):
        return np.array([self.canvasWidth, self.canvasHeight, self.canvasDepth], dtype=np.float32)


def build(force: bool = False) -> str:
    """Compile liboracle.so (and oracle/_ref when the reference tree is present)."""
    src = os.path.join(_HERE, "cellflow_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_ratio_with_lfo.restype = C.c_float
        _lib.orc_ratio_with_lfo.argtypes = [C.POINTER(Params), C.c_float]
        _lib.orc_graph_bruteforce.restype = C.c_int
        _lib.orc_graph_cells.restype = C.c_int
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def max_threads() -> int:
    return int(lib().orc_max_threads())


def particles(n: int) -> np.ndarray:
    return np.zeros(n, dtype=PARTICLE)


def force_table(raw, T, frange, fbias, foffset) -> np.ndarray:
    raw = np.ascontiguousarray(raw, dtype=np.float32)
    out = np.zeros(T * T, dtype=np.float32)
    lib().orc_force_table(_p(raw), C.c_int(T), C.c_float(frange), C.c_float(fbias),
                          C.c_float(foffset), _p(out))
    return out


def default_tables(T: int):
    raw = np.zeros(T * T, dtype=np.float32)
    radio = np.zeros(T, dtype=np.float32)
    lib().orc_default_tables(C.c_int(T), _p(raw), _p(radio))
    return raw, radio


def ratio_with_lfo(params: Params, t: float) -> float:
    return float(lib().orc_ratio_with_lfo(C.byref(params), C.c_float(t)))


def reff_table(params: Params, radio) -> np.ndarray:
    T = params.numParticleTypes
    radio = np.ascontiguousarray(radio, dtype=np.float32)
    out = np.zeros(T * T, dtype=np.float32)
    lib().orc_reff_table(C.byref(params), _p(radio), _p(out))
    return out


def step(pin, cnt_in, params: Params, table, radio, method="cells", threads=1):
    """One Jacobi step.  Returns (particles_out, counts_out, fabs) — fabs[i] is the sum over
    accepted pairs of |forceValue*netForce|, the scale force errors are measured against."""
    n = len(pin)
    pin = np.ascontiguousarray(pin, dtype=PARTICLE)
    cnt_in = np.ascontiguousarray(cnt_in if cnt_in is not None else np.zeros(n), dtype=np.int32)
    table = np.ascontiguousarray(table, dtype=np.float32)
    radio = np.ascontiguousarray(radio, dtype=np.float32)
    out = np.zeros(n, dtype=PARTICLE)
    cnt = np.zeros(n, dtype=np.int32)
    fabs = np.zeros(n, dtype=np.float32)
    fn = lib().orc_step_cells if method == "cells" else lib().orc_step_bruteforce
    fn(_p(pin), _p(cnt_in), C.c_int(n), C.byref(params), _p(table), _p(radio), _p(out), _p(cnt),
       _p(fabs), C.c_int(threads))
    return out, cnt, fabs


def step_norms(pin, cnt_in, params: Params, table, radio, threads=1):
    """Cell-list step that also returns fnet[i] = sum_j |f_ij| (the net pair forces' magnitudes): the
    norm SURVEY.md section 7 names, reported beside the gross-term norm `fabs`."""
    n = len(pin)
    fnet = np.zeros(n, dtype=np.float32)
    lib().orc_set_fnet_output(_p(fnet))
    try:
        out, cnt, fabs = step(pin, cnt_in, params, table, radio, "cells", threads)
    finally:
        lib().orc_set_fnet_output(None)
    return out, cnt, fabs, fnet


def set_sort_candidates(on: bool):
    lib().orc_set_sort_candidates(C.c_int(1 if on else 0))


def step_range(pin, cnt_in, params: Params, table, radio, i0, i1, threads=1):
    """Cell-list step of particles [i0, i1) only (outputs outside the range stay zero)."""
    n = len(pin)
    pin = np.ascontiguousarray(pin, dtype=PARTICLE)
    cnt_in = np.ascontiguousarray(cnt_in if cnt_in is not None else np.zeros(n), dtype=np.int32)
    table = np.ascontiguousarray(table, dtype=np.float32)
    radio = np.ascontiguousarray(radio, dtype=np.float32)
    out = np.zeros(n, dtype=PARTICLE)
    cnt = np.zeros(n, dtype=np.int32)
    fabs = np.zeros(n, dtype=np.float32)
    lib().orc_step_cells_range(_p(pin), _p(cnt_in), C.c_int(n), C.c_int(i0), C.c_int(i1),
                               C.byref(params), _p(table), _p(radio), _p(out), _p(cnt), _p(fabs),
                               C.c_int(threads))
    return out, cnt, fabs


def step_f64(pin, cnt_in, params: Params, table, radio, threads=1):
    n = len(pin)
    pin = np.ascontiguousarray(pin, dtype=PARTICLE)
    cnt_in = np.ascontiguousarray(cnt_in if cnt_in is not None else np.zeros(n), dtype=np.int32)
    table = np.ascontiguousarray(table, dtype=np.float32)
    radio = np.ascontiguousarray(radio, dtype=np.float32)
    force = np.zeros((n, 3), dtype=np.float64)
    pos = np.zeros((n, 3), dtype=np.float64)
    fabs = np.zeros(n, dtype=np.float64)
    lib().orc_step_f64(_p(pin), _p(cnt_in), C.c_int(n), C.byref(params), _p(table), _p(radio),
                       _p(force), _p(pos), _p(fabs), C.c_int(threads))
    return force, pos, fabs


def graph(pin, dist, max_conn, canvas=None, method="cells") -> np.ndarray:
    n = len(pin)
    pin = np.ascontiguousarray(pin, dtype=PARTICLE)
    cap = max(1, n * min(max_conn, MAX_CONN))
    edges = np.zeros(cap, dtype=EDGE)
    if method == "cells":
        canvas = np.ascontiguousarray(canvas, dtype=np.float32)
        ne = lib().orc_graph_cells(_p(pin), C.c_int(n), _p(canvas), C.c_float(dist),
                                   C.c_int(max_conn), _p(edges), C.c_int(cap))
    else:
        ne = lib().orc_graph_bruteforce(_p(pin), C.c_int(n), C.c_float(dist), C.c_int(max_conn),
                                        _p(edges), C.c_int(cap))
    return edges[:ne].copy()


def graph_vertices(pin, edges, colors, num_types) -> np.ndarray:
    pin = np.ascontiguousarray(pin, dtype=PARTICLE)
    edges = np.ascontiguousarray(edges, dtype=EDGE)
    colors = np.ascontiguousarray(colors, dtype=COLOR)
    out = np.zeros((len(edges), 12), dtype=np.float32)
    lib().orc_graph_vertices(_p(pin), _p(edges), C.c_int(len(edges)), _p(colors),
                             C.c_int(num_types), _p(out))
    return out


def move_universe(pin, dx, dy, dz, canvas) -> np.ndarray:
    p = np.ascontiguousarray(pin, dtype=PARTICLE).copy()
    canvas = np.ascontiguousarray(canvas, dtype=np.float32)
    lib().orc_move_universe(_p(p), C.c_int(len(p)), C.c_float(dx), C.c_float(dy), C.c_float(dz),
                            _p(canvas))
    return p


def init_particles(n, T, seed, mode, canvas, id0=0) -> np.ndarray:
    out = np.zeros(n, dtype=PARTICLE)
    canvas = np.ascontiguousarray(canvas, dtype=np.float32)
    lib().orc_init_particles(_p(out), C.c_int(n), C.c_int(id0), C.c_int(T), C.c_uint64(seed),
                             C.c_int(mode), _p(canvas))
    return out


def cell_keys(pin, canvas, dims) -> np.ndarray:
    pin = np.ascontiguousarray(pin, dtype=PARTICLE)
    canvas = np.ascontiguousarray(canvas, dtype=np.float32)
    dims = np.ascontiguousarray(dims, dtype=np.int32)
    keys = np.zeros(len(pin), dtype=np.uint32)
    lib().orc_cell_keys(_p(pin), C.c_int(len(pin)), _p(canvas), _p(dims), _p(keys))
    return keys
