"""Python mirror of the reference's `class ParticleSimulation`
(cuda-native/include/ParticleSimulation.cuh:10-75) on top of the C ABI.

Method names, argument meaning and error behaviour follow the reference so that the parity
tests read like tests of the reference class; every call goes straight to
lib/libcellflow_b200.so.  Host buffers are numpy arrays in the reference's 44-byte `Particle`
layout (`PARTICLE` dtype).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import COLOR, EDGE, PARTICLE, Params, Preset, Stats, check


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class ParticleSimulation:
    """ParticleSimulation(particleCount) — reference ctor at ParticleSimulation.cu:427-435."""

    def __init__(self, particleCount: int, numParticleTypes: int = 6, device: int = 0,
                 init: bool = True, seed: int = 0x5EED0000):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        check(self._L.cf_create(C.c_int(particleCount), C.c_int(numParticleTypes), C.c_int(device),
                                C.byref(self._h)))
        self.params = _lib.default_params(numParticleTypes=numParticleTypes)
        if init:
            self.initializeParticles(seed=seed)

    # -- lifetime -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.cf_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- reference API ----------------------------------------------------------------------
    def initializeParticles(self, canvasWidth=None, canvasHeight=None, seed=0x5EED0000,
                            mode=_lib.INIT_SPAWN_CUBE):
        """initializeParticles(), .cu:484-505 (spawn cube hard-coded to 2000 there)."""
        if canvasWidth is not None:
            self.updateCanvasDimensions(canvasWidth, canvasHeight)
        check(self._L.cf_set_params(self._h, C.byref(self.params)))
        check(self._L.cf_init_particles(self._h, C.c_uint64(seed), C.c_int(mode)))

    def updateCanvasDimensions(self, canvasWidth, canvasHeight):
        """.cu:507-511: depth defaults to height."""
        self.params.canvasWidth = canvasWidth
        self.params.canvasHeight = canvasHeight
        self.params.canvasDepth = canvasHeight

    def initializeForceTable(self):
        check(self._L.cf_regenerate_force_table(self._h))

    regenerateForceTable = initializeForceTable  # .cu:581-583

    def updateForceTable(self, forceRange, forceBias, forceOffset):
        check(self._L.cf_update_force_table(self._h, C.c_float(forceRange), C.c_float(forceBias),
                                            C.c_float(forceOffset)))
        self.params.forceRange, self.params.forceBias, self.params.forceOffset = (
            forceRange, forceBias, forceOffset)

    def simulate(self, params: Params | None = None, steps: int = 1, sync: bool = True):
        """simulate(const SimulationParams&), .cu:541-556: params travel with every call."""
        if params is not None:
            self.params = params
        check(self._L.cf_step(self._h, C.byref(self.params), C.c_int(steps)))
        if sync:
            check(self._L.cf_sync(self._h))

    def sync(self):
        check(self._L.cf_sync(self._h))

    def getParticleData(self) -> np.ndarray:
        """getParticleData(std::vector<Particle>&), .cu:558-562; original particle order."""
        n = self.getParticleCount()
        out = np.zeros(n, dtype=PARTICLE)
        check(self._L.cf_download_particles(self._h, _p(out), C.c_int(n)))
        return out

    def setParticleData(self, particles: np.ndarray, counts: np.ndarray | None = None):
        particles = np.ascontiguousarray(particles, dtype=PARTICLE)
        n = len(particles)
        if n != self.getParticleCount():
            check(self._L.cf_upload_particles_ids(self._h, _p(particles), None, None, C.c_int(n)))
        else:
            check(self._L.cf_upload_particles(self._h, _p(particles), C.c_int(n)))
        if counts is not None:
            self.setNeighborCounts(counts)

    def getNeighborCounts(self) -> np.ndarray:
        n = self.getParticleCount()
        out = np.zeros(n, dtype=np.int32)
        check(self._L.cf_download_neighbor_counts(self._h, _p(out), C.c_int(n)))
        return out

    def setNeighborCounts(self, counts):
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        check(self._L.cf_upload_neighbor_counts(self._h, _p(counts), C.c_int(len(counts))))

    def setParticleCount(self, count: int):
        check(self._L.cf_set_particle_count(self._h, C.c_int(count)))

    def getParticleCount(self) -> int:
        return int(self._L.cf_get_particle_count(self._h))

    def setNumParticleTypes(self, types: int):
        check(self._L.cf_set_num_particle_types(self._h, C.c_int(types)))
        self.params.numParticleTypes = types

    def getNumParticleTypes(self) -> int:
        return int(self._L.cf_get_num_particle_types(self._h))

    def getRawForceTableValues(self) -> np.ndarray:
        T = self.getNumParticleTypes()
        out = np.zeros(T * T, dtype=np.float32)
        check(self._L.cf_get_raw_force_table(self._h, _p(out), C.c_int(T * T)))
        return out

    def setRawForceTableValues(self, raw):
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        check(self._L.cf_set_raw_force_table(self._h, _p(raw), C.c_int(len(raw))))

    def getForceTable(self) -> np.ndarray:
        T = self.getNumParticleTypes()
        out = np.zeros(T * T, dtype=np.float32)
        check(self._L.cf_get_force_table(self._h, _p(out), C.c_int(T * T)))
        return out

    def setForceTable(self, eff):
        eff = np.ascontiguousarray(eff, dtype=np.float32)
        check(self._L.cf_set_force_table(self._h, _p(eff), C.c_int(len(eff))))

    def moveUniverse(self, dx, dy, dz=0.0):
        check(self._L.cf_set_params(self._h, C.byref(self.params)))
        check(self._L.cf_move_universe(self._h, C.c_float(dx), C.c_float(dy), C.c_float(dz)))

    def rotateRadioByType(self):
        check(self._L.cf_rotate_radio_by_type(self._h))

    def getRadioByType(self) -> np.ndarray:
        T = self.getNumParticleTypes()
        out = np.zeros(T, dtype=np.float32)
        check(self._L.cf_get_radio_by_type(self._h, _p(out), C.c_int(T)))
        return out

    def setRadioByTypeValue(self, index: int, value: float):
        check(self._L.cf_set_radio_by_type_value(self._h, C.c_int(index), C.c_float(value)))

    def setRadioByType(self, radio):
        radio = np.ascontiguousarray(radio, dtype=np.float32)
        check(self._L.cf_set_radio_by_type(self._h, _p(radio), C.c_int(len(radio))))

    def buildGraphAsync(self, proximityDistance: float, maxConnectionsPerParticle: int):
        """The graph build enqueued without reading anything back (multi-GPU step loops)."""
        check(self._L.cf_set_params(self._h, C.byref(self.params)))
        check(self._L.cf_build_graph(self._h, C.c_float(proximityDistance), C.c_int(maxConnectionsPerParticle), None))

    def graphEdgeCount(self) -> int:
        ne = C.c_int(0)
        check(self._L.cf_get_graph_edge_count(self._h, C.byref(ne)))
        return ne.value

    def generateProximityGraph(self, proximityDistance: float, maxConnectionsPerParticle: int,
                               particleColors=None):
        """generateProximityGraph(...), .cu:625-687.  Returns (edges, vertices): `edges` is the
        (i, j) list, `vertices` the reference VBO content (12 floats per edge) when colours are
        given.  outVertexCount of the reference = 2 * len(edges)."""
        check(self._L.cf_set_params(self._h, C.byref(self.params)))
        ne = C.c_int(0)
        check(self._L.cf_build_graph(self._h, C.c_float(proximityDistance),
                                     C.c_int(maxConnectionsPerParticle), C.byref(ne)))
        edges = np.zeros(ne.value, dtype=EDGE)
        check(self._L.cf_download_graph_edges(self._h, _p(edges), C.c_int(len(edges))))
        verts = None
        if particleColors is not None:
            colors = np.ascontiguousarray(particleColors, dtype=COLOR)
            verts = np.zeros((ne.value, 12), dtype=np.float32)
            check(self._L.cf_download_graph_vertices(self._h, _p(colors), C.c_int(len(colors)),
                                                     _p(verts), C.c_int(ne.value)))
        return edges, verts

    def graphVerticesDevice(self, particleColors, device_dst: int = 0, capacity_edges: int = 0) -> int:
        """The last graph's vertex stream written on the device (zero-copy hand-off to a renderer): into
        `device_dst` (a device pointer the caller owns, e.g. a mapped GL VBO) or into the library's persistent
        buffer.  Returns the device address of the stream.  Asynchronous; sync() orders it."""
        colors = np.ascontiguousarray(particleColors, dtype=COLOR)
        ptr = C.c_void_p()
        check(self._L.cf_graph_vertices_device(self._h, _p(colors), C.c_int(len(colors)), C.c_void_p(device_dst or None),
                                               C.c_int(capacity_edges), C.byref(ptr)))
        return ptr.value or 0

    # -- beyond the reference class ------------------------------------------------------------
    def saveSnapshot(self, path: str):
        """Parameters, tables and the full particle state (pos, vel, acc, type, previous count, id)."""
        import os
        check(self._L.cf_set_params(self._h, C.byref(self.params)))
        check(self._L.cf_save_snapshot(self._h, os.fsencode(path)))

    def loadSnapshot(self, path: str):
        import os
        check(self._L.cf_load_snapshot(self._h, os.fsencode(path)))
        p = Params()
        check(self._L.cf_get_params(self._h, C.byref(p)))
        self.params = p

    def applyPreset(self, preset: Preset):
        """CellFlowWidget::loadPreset applied to the simulation (CellFlowWidget.cpp:1079-1177)."""
        check(self._L.cf_apply_preset(self._h, C.byref(preset)))
        p = Params()
        check(self._L.cf_get_params(self._h, C.byref(p)))
        self.params = p

    def stepHost(self, particles, counts=None, params: Params | None = None):
        """Stateless step with host buffers (H2D, step, D2H inside the call)."""
        if params is not None:
            self.params = params
        particles = np.ascontiguousarray(particles, dtype=PARTICLE)
        n = len(particles)
        cin = np.ascontiguousarray(counts if counts is not None else np.zeros(n), dtype=np.int32)
        out = np.zeros(n, dtype=PARTICLE)
        cout = np.zeros(n, dtype=np.int32)
        check(self._L.cf_step_host(self._h, C.byref(self.params), _p(particles), _p(cin), _p(out),
                                   _p(cout), C.c_int(n)))
        return out, cout

    def stepHostInto(self, pin, cin, pout, cout):
        """Same as stepHost but on caller-owned (e.g. pinned) buffers given as raw addresses."""
        check(self._L.cf_step_host(self._h, C.byref(self.params), C.c_void_p(pin[0]),
                                   C.c_void_p(cin[0]), C.c_void_p(pout[0]), C.c_void_p(cout[0]),
                                   C.c_int(pin[1])))

    def getRenderFeed(self):
        """(xyzt[n,4], type_counts[T]): the widget's per-frame vertex data and type histogram
        (CellFlowWidget.cpp:742-761, 875-886), built on the device."""
        n, T = self.getParticleCount(), self.getNumParticleTypes()
        xyzt = np.zeros((n, 4), dtype=np.float32)
        counts = np.zeros(T, dtype=np.int32)
        check(self._L.cf_render_feed(self._h, _p(xyzt), C.c_int(n), _p(counts), None))
        return xyzt, counts

    # -- multi-GPU slabs (one process per GPU) ----------------------------------------------------
    def commInit(self, rank: int, world: int, capacity: int):
        """Rank `rank` of `world` owns one x slab (uniform split unless setSlabBounds says otherwise).
        world == 1 runs the same slab code path with the rank's own mailbox as both neighbours; world > 1
        needs mailboxHandle() of both ring neighbours passed to connect() (cellflow_b200/dist.py does it)."""
        check(self._L.cf_set_params(self._h, C.byref(self.params)))
        check(self._L.cf_comm_init(self._h, C.c_int(rank), C.c_int(world), C.c_int(capacity)))
        self._slab_capacity = capacity

    def mailboxHandle(self) -> bytes:
        """CUDA IPC handle (64 bytes) of this rank's mailbox, for the two ring neighbours."""
        buf = C.create_string_buffer(64)
        check(self._L.cf_comm_mailbox_handle(self._h, buf))
        return buf.raw

    def connect(self, left_handle: bytes, right_handle: bytes):
        check(self._L.cf_comm_connect(self._h, C.create_string_buffer(left_handle, 64),
                                      C.create_string_buffer(right_handle, 64)))

    def setSlabBounds(self, bounds):
        b = np.ascontiguousarray(bounds, dtype=np.float32)
        check(self._L.cf_set_params(self._h, C.byref(self.params)))
        check(self._L.cf_slab_set_bounds(self._h, _p(b), C.c_int(len(b))))

    def initParticlesGlobal(self, n_total: int, seed: int, mode=_lib.INIT_UNIFORM):
        check(self._L.cf_set_params(self._h, C.byref(self.params)))
        check(self._L.cf_init_particles_global(self._h, C.c_int64(n_total), C.c_uint64(seed), C.c_int(mode)))

    def slabBounds(self):
        lo, hi = C.c_float(0), C.c_float(0)
        check(self._L.cf_set_params(self._h, C.byref(self.params)))
        check(self._L.cf_slab_bounds(self._h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def uploadOwned(self, particles, counts, ids):
        particles = np.ascontiguousarray(particles, dtype=PARTICLE)
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        check(self._L.cf_upload_particles_ids(self._h, _p(particles), _p(counts), _p(ids),
                                              C.c_int(len(particles))))

    def downloadOwned(self):
        """(particles, counts, ids) of the particles this rank currently owns (slot order)."""
        cap = max(getattr(self, "_slab_capacity", 0), self.getParticleCount(), 1)
        out = np.zeros(cap, dtype=PARTICLE)
        counts = np.zeros(cap, dtype=np.int32)
        ids = np.zeros(cap, dtype=np.int32)
        n = C.c_int(0)
        check(self._L.cf_download_particles_ids(self._h, _p(out), _p(counts), _p(ids), C.c_int(cap),
                                                C.byref(n)))
        return out[: n.value], counts[: n.value], ids[: n.value]

    def cellKeys(self):
        n = self.getParticleCount()
        keys = np.zeros(n, dtype=np.uint32)
        ids = np.zeros(n, dtype=np.int32)
        cnt = C.c_int(0)
        check(self._L.cf_set_params(self._h, C.byref(self.params)))
        check(self._L.cf_download_cell_keys(self._h, _p(keys), _p(ids), C.c_int(n), C.byref(cnt)))
        return keys, ids

    def stats(self) -> Stats:
        st = Stats()
        check(self._L.cf_get_stats(self._h, C.byref(st)))
        return st

    def statsReset(self):
        check(self._L.cf_stats_reset(self._h))

    def setOption(self, name: str, value: float):
        check(self._L.cf_set_option(self._h, name.encode(), C.c_double(value)))
