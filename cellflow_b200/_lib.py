"""ctypes binding of include/cellflow_b200.h.  Loads lib/libcellflow_b200.so and nothing else:
if the library is missing this raises — there is no Python or CPU implementation to fall back to."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CELLFLOW_B200_LIB: another build of the same library (kernel A/B experiments, tools/r02_ab.sh)
_LIBPATH = os.environ.get("CELLFLOW_B200_LIB") or os.path.join(_HERE, "lib", "libcellflow_b200.so")

MAX_TYPES = 10
MAX_GRAPH_CONN = 16
INIT_SPAWN_CUBE = 0
INIT_UNIFORM = 1

PARTICLE = np.dtype(
    [("pos", "<f4", 3), ("vel", "<f4", 3), ("acc", "<f4", 3), ("ptype", "<u4"), ("pad", "<f4")]
)
EDGE = np.dtype([("i", "<i4"), ("j", "<i4")])
COLOR = np.dtype([("r", "<f4"), ("g", "<f4"), ("b", "<f4")])
assert PARTICLE.itemsize == 44


class CellFlowError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cellflow_b200 error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    """cf_params: physics subset of the reference SimulationParams (SimulationParams.h:15-37)."""

    _fields_ = [
        ("radius", C.c_float), ("delta_t", C.c_float), ("friction", C.c_float),
        ("repulsion", C.c_float), ("attraction", C.c_float), ("k", C.c_float),
        ("balance", C.c_float), ("canvasWidth", C.c_float), ("canvasHeight", C.c_float),
        ("canvasDepth", C.c_float), ("spawnRegionSize", C.c_float),
        ("numParticleTypes", C.c_int32), ("ratioWithLFO", C.c_float),
        ("forceMultiplier", C.c_float), ("maxExpectedNeighbors", C.c_int32),
        ("forceRange", C.c_float), ("forceBias", C.c_float), ("ratio", C.c_float),
        ("lfoA", C.c_float), ("lfoS", C.c_float), ("forceOffset", C.c_float),
    ]

    def copy(self, **kw):
        p = Params()
        C.memmove(C.byref(p), C.byref(self), C.sizeof(Params))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}

    @property
    def canvas(self):
        return np.array([self.canvasWidth, self.canvasHeight, self.canvasDepth], dtype=np.float32)


class Color(C.Structure):
    _fields_ = [("r", C.c_float), ("g", C.c_float), ("b", C.c_float)]


class Preset(C.Structure):
    """cf_preset: everything CellFlowWidget::loadPreset reads (CellFlowWidget.cpp:1070-1180)."""

    _fields_ = [
        ("particleCount", C.c_int32), ("params", Params), ("pointSize", C.c_float),
        ("depthFadeStart", C.c_float), ("depthFadeEnd", C.c_float),
        ("sizeAttenuationFactor", C.c_float), ("brightnessMin", C.c_float),
        ("focusDistance", C.c_float), ("apertureSize", C.c_float),
        ("enableDepthFade", C.c_int32), ("enableSizeAttenuation", C.c_int32),
        ("enableBrightnessAttenuation", C.c_int32), ("enableDOF", C.c_int32),
        ("invertPan", C.c_int32), ("invertForwardBack", C.c_int32), ("invertRotation", C.c_int32),
        ("effectType", C.c_int32), ("numColors", C.c_int32),
        ("particleColors", Color * MAX_TYPES), ("numRadio", C.c_int32),
        ("radioByType", C.c_float * MAX_TYPES), ("numRawForce", C.c_int32),
        ("rawForceTable", C.c_float * (MAX_TYPES * MAX_TYPES)),
    ]

    @property
    def radio(self):
        return np.array(self.radioByType[: self.numRadio], dtype=np.float32)

    @property
    def raw_force(self):
        return np.array(self.rawForceTable[: self.numRawForce], dtype=np.float32)

    @property
    def colors(self):
        out = np.zeros(MAX_TYPES, dtype=COLOR)
        for i in range(self.numColors):
            c = self.particleColors[i]
            out[i] = (c.r, c.g, c.b)
        return out


class Stats(C.Structure):
    _fields_ = [
        ("ms_total", C.c_double), ("ms_sort", C.c_double), ("ms_force", C.c_double),
        ("ms_integrate", C.c_double), ("ms_exchange", C.c_double), ("ms_graph", C.c_double),
        ("steps", C.c_int64), ("launches", C.c_int64), ("accepted_pairs", C.c_int64),
        ("tested_pairs", C.c_int64), ("grid", C.c_int32 * 3), ("stencil", C.c_int32),
        ("n_owned", C.c_int32), ("n_ghost", C.c_int32), ("force_kernel", C.c_int32), ("graph_kernel", C.c_int32),
        ("ms_graph_total", C.c_double), ("graph_builds", C.c_int64),
        ("ms_exchange_migrants", C.c_double), ("ms_exchange_halo", C.c_double),
        ("exact_tested_pairs", C.c_int64), ("evaluated_pair_lanes", C.c_int64),
        ("ms_step_max", C.c_double), ("ms_exchange_max", C.c_double),
    ]


# every symbol include/cellflow_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "cf_create", "cf_destroy", "cf_set_particle_count", "cf_get_particle_count",
    "cf_set_num_particle_types", "cf_get_num_particle_types", "cf_regenerate_force_table",
    "cf_reference_default_tables", "cf_set_raw_force_table", "cf_get_raw_force_table",
    "cf_update_force_table", "cf_get_force_table", "cf_set_force_table", "cf_set_radio_by_type",
    "cf_set_radio_by_type_value", "cf_get_radio_by_type", "cf_rotate_radio_by_type",
    "cf_init_particles", "cf_upload_particles", "cf_download_particles",
    "cf_upload_neighbor_counts", "cf_download_neighbor_counts", "cf_render_feed", "cf_move_universe",
    "cf_set_params", "cf_get_params", "cf_step", "cf_sync", "cf_step_host", "cf_ratio_with_lfo",
    "cf_build_graph", "cf_get_graph_edge_count", "cf_download_graph_edges", "cf_download_graph_vertices", "cf_graph_vertices_device",
    "cf_default_params", "cf_default_preset", "cf_load_preset", "cf_save_preset",
    "cf_apply_preset", "cf_save_snapshot", "cf_load_snapshot", "cf_comm_init", "cf_comm_mailbox_handle", "cf_comm_connect", "cf_slab_set_bounds", "cf_init_particles_global", "cf_slab_bounds",
    "cf_upload_particles_ids",
    "cf_download_particles_ids", "cf_get_stats", "cf_stats_reset", "cf_download_cell_keys",
    "cf_set_option", "cf_bench_fp32_peak", "cf_bench_flush_l2", "cf_bench_flush_l2_async", "cf_last_error", "cf_version",
]


def lib_path() -> str:
    return _LIBPATH


def build(force: bool = False) -> str:
    """Compile the CUDA library in-tree (nvcc, sm_100a).  Used by __graft_entry__.build()."""
    src = os.path.join(_HERE, "csrc")
    args = ["make", "-s", "-C", src]
    if force:
        args.append("-B")
    subprocess.check_call(args)
    return _LIBPATH


_lib = None


def lib():
    """The loaded C library.  Raises if it has not been built — no fallback exists."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIBPATH):
            raise CellFlowError(
                -2, f"{_LIBPATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). cellflow_b200 has no CPU or PyTorch fallback.")
        L = C.CDLL(_LIBPATH)
        L.cf_last_error.restype = C.c_char_p
        L.cf_version.restype = C.c_char_p
        L.cf_ratio_with_lfo.restype = C.c_float
        L.cf_ratio_with_lfo.argtypes = [C.POINTER(Params), C.c_float]
        L.cf_default_params.restype = None
        L.cf_default_preset.restype = None
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise CellFlowError(rc, lib().cf_last_error().decode(errors="replace"))


def default_params(**kw) -> Params:
    p = Params()
    lib().cf_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def load_preset(path: str) -> Preset:
    pr = Preset()
    lib().cf_default_preset(C.byref(pr))
    rc = lib().cf_load_preset(os.fsencode(path), C.byref(pr))
    if rc != 0:
        raise CellFlowError(rc, f"cannot load preset {path}")
    return pr


def save_preset(path: str, preset: Preset):
    rc = lib().cf_save_preset(os.fsencode(path), C.byref(preset))
    if rc != 0:
        raise CellFlowError(rc, f"cannot save preset {path}")


def reference_default_tables(T: int):
    raw = np.zeros(T * T, np.float32)
    radio = np.zeros(T, np.float32)
    eff = np.zeros(T * T, np.float32)
    check(lib().cf_reference_default_tables(C.c_int(T), raw.ctypes.data_as(C.c_void_p),
                                            radio.ctypes.data_as(C.c_void_p),
                                            eff.ctypes.data_as(C.c_void_p)))
    return raw, radio, eff
