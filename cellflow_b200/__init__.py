"""cellflow_b200 — B200-native particle-life engine with CellFlow's simulation semantics.

The product is `lib/libcellflow_b200.so` (hand-written sm_100a CUDA behind the C ABI of
include/cellflow_b200.h).  This package is the thin Python host side used by the tests and
bench.py: a ctypes binding (`_lib`), a mirror of the reference's `ParticleSimulation` class
(`sim.ParticleSimulation`, reference: cuda-native/include/ParticleSimulation.cuh:10-75) and the
one-process-per-GPU launcher glue (`dist`).  There is no CPU fallback: importing works anywhere,
creating a simulation without the CUDA library or without a GPU raises.
"""
from ._lib import (  # noqa: F401
    CellFlowError, Params, Preset, Stats, PARTICLE, EDGE, COLOR, INIT_SPAWN_CUBE, INIT_UNIFORM,
    lib, lib_path, build, load_preset, save_preset, default_params, reference_default_tables,
)
from .sim import ParticleSimulation  # noqa: F401

__all__ = [
    "CellFlowError", "Params", "Preset", "Stats", "PARTICLE", "EDGE", "COLOR", "ParticleSimulation",
    "INIT_SPAWN_CUBE", "INIT_UNIFORM", "lib", "lib_path", "build", "load_preset", "save_preset",
    "default_params", "reference_default_tables",
]
