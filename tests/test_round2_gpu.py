"""GPU tests added in round 2: the exact bench workloads against the oracle (VERDICT r01 'parity holes'), the
particle snapshot (restore and step equals), the device vertex hand-off and its staleness check, the
clamped particle types, the table generator's rand() stream, the block counters of the tile kernel."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import cellflow_b200 as cf
from cellflow_b200 import _lib
import oracle as O
import util as U

sys.path.insert(0, U.ROOT)
import bench as B  # noqa: E402  (the workloads under test are bench.py's own definitions)

pytestmark = pytest.mark.gpu
THREADS = max(1, O.max_threads())


def bench_sim(name, **opts):
    """The simulation exactly as bench.py sets it up for workload `name` on one GPU + the oracle's view of it."""
    spec = B.workload_spec(name, 1)
    params, raw, radio = B.cf_setup(spec)
    sim = cf.ParticleSimulation(spec["n_total"], spec["T"], init=False)
    sim.params = params
    sim.setRadioByType(radio)
    sim.setRawForceTableValues(raw)
    sim.updateForceTable(params.forceRange, params.forceBias, params.forceOffset)
    sim.initializeParticles(seed=spec["seed"], mode=spec["mode"])
    for k, v in opts.items():
        sim.setOption(k, v)
    _, op, table, oradio = B.oracle_setup(spec)
    state = O.init_particles(spec["n_total"], spec["T"], spec["seed"], spec["mode"], op.canvas)
    return spec, sim, op, table, oradio, state


@pytest.mark.parametrize("name", ["c3-eater-1M", "c5-settings-2M"])
def test_exact_bench_workload_one_step(name, capsys):
    """One step of the bench's own state (same preset, radii, ratio, seed, count, grid, kernel choice):
    counts exact, forces within 1e-5; c5 also the proximity graph (200, 5) on the stepped state."""
    spec, sim, op, table, radio, state = bench_sim(name)
    assert sim.getParticleData().tobytes() == state.tobytes()       # same spawn on both sides
    counts = np.zeros(len(state), np.int32)
    sim.simulate()
    got, gcnt = sim.getParticleData(), sim.getNeighborCounts()
    want, wcnt, fabs, fnet = O.step_norms(state, counts, op, table, radio, THREADS)
    assert np.array_equal(gcnt, wcnt), f"{(gcnt != wcnt).sum()} of {len(wcnt)} counts differ"
    mult = U.force_multiplier_of(op, wcnt, counts)
    rep = U.force_err_report(got["acc"], want["acc"], fabs, fnet, mult)
    assert rep["gross"] <= U.FORCE_RTOL, rep
    st = sim.stats()
    with capsys.disabled():
        print(f"\n[{name}] kernel {st.force_kernel} grid {list(st.grid)} mean nbrs {gcnt.mean():.1f} force error: "
              f"gross-term norm {rep['gross']:.2e}  sum|f_ij| norm {rep['net']:.2e}  |F|inf norm {rep['finf']:.2e}")
    if spec["graph"]:
        edges, _ = sim.generateProximityGraph(*spec["graph"])
        want_e = O.graph(got, spec["graph"][0], spec["graph"][1], canvas=op.canvas, method="cells")
        assert len(edges) == len(want_e) and U.edge_set(edges) == U.edge_set(want_e)
    sim.close()


def small_sim(n=40000, kernel=3, seed=11):
    p, table, _ = U.config("eater", ratioWithLFO=0.5, canvasWidth=3200.0, canvasHeight=3200.0, canvasDepth=3200.0)
    radio = np.float32([1.0, 0.5, 0.0, 0.0, -0.5, 1.0])
    state, counts = U.random_state(n, 6, seed, p.canvas, "uniform")
    sim = cf.ParticleSimulation(n, 6, init=False)
    sim.params = U.to_lib_params(p)
    sim.setRadioByType(radio)
    sim.setForceTable(table)
    sim.setOption("force_kernel", kernel)
    sim.setParticleData(state, counts)
    return sim, p, table, radio, state, counts


@pytest.mark.parametrize("kernel", [1, 3])
def test_snapshot_restore_and_step_equals(tmp_path, kernel):
    """save after 3 steps, continue 4 more; a fresh handle that loads the file and runs the same 4 steps ends in
    the bit-identical state (particles, counts) — the snapshot carries pos, vel, type, previous count, id, the
    slot order, parameters and tables."""
    sim, p, table, radio, state, counts = small_sim(kernel=kernel)
    sim.simulate(steps=3)
    path = str(tmp_path / "state.cfsnap")
    sim.saveSnapshot(path)
    sim.simulate(steps=4)
    a, ac = sim.getParticleData(), sim.getNeighborCounts()
    sim.close()
    fresh = cf.ParticleSimulation(7, 3, init=False)          # wrong count and types on purpose
    fresh.setOption("force_kernel", kernel)
    fresh.loadSnapshot(path)
    assert fresh.getParticleCount() == len(state) and fresh.getNumParticleTypes() == 6
    assert np.allclose(fresh.getRadioByType(), radio) and fresh.params.radius == p.radius
    fresh.simulate(steps=4)
    b, bc = fresh.getParticleData(), fresh.getNeighborCounts()
    assert np.array_equal(ac, bc)
    assert a.tobytes() == b.tobytes()
    fresh.close()
    with pytest.raises(cf.CellFlowError):
        bad = str(tmp_path / "bad.cfsnap")
        open(bad, "wb").write(b"not a snapshot")
        s2 = cf.ParticleSimulation(4, 6)
        try:
            s2.loadSnapshot(bad)
        finally:
            s2.close()


def test_graph_vertices_device_handoff_and_staleness():
    """cf_graph_vertices_device: the stream written into caller-owned device memory and into the library's
    persistent buffer equals the host download; after a reorder of the particles the request is refused instead
    of returning vertices of the wrong particles."""
    import torch
    sim, p, table, radio, state, counts = small_sim(n=30000)
    colors = np.zeros(10, cf.COLOR)
    colors["r"], colors["g"], colors["b"] = np.arange(10) * 0.1, 0.25, 1.0 - np.arange(10) * 0.1
    edges, verts = sim.generateProximityGraph(180.0, 5, colors)
    ne = len(edges)
    assert ne > 0 and np.array_equal(verts, O.graph_vertices(state, edges, colors, 6))
    # caller-owned device buffer (stands in for the mapped GL VBO of the widget)
    vbo = torch.zeros(ne * 12, dtype=torch.float32, device="cuda")
    addr = sim.graphVerticesDevice(colors, vbo.data_ptr(), ne)
    sim.sync()
    assert addr == vbo.data_ptr()
    assert np.array_equal(vbo.cpu().numpy().reshape(ne, 12), verts)
    # the library's persistent buffer: same content, same address on the next call (no malloc per call)
    a1 = sim.graphVerticesDevice(colors)
    a2 = sim.graphVerticesDevice(colors)
    sim.sync()
    assert a1 == a2 and a1 != 0
    out = torch.zeros(ne * 12, dtype=torch.float32, device="cuda")
    rt = C.CDLL("libcudart.so.12")
    assert rt.cudaMemcpy(C.c_void_p(out.data_ptr()), C.c_void_p(a1), C.c_size_t(ne * 48), C.c_int(3)) == 0
    assert np.array_equal(out.cpu().numpy().reshape(ne, 12), verts)
    # too small a caller buffer is refused
    with pytest.raises(cf.CellFlowError):
        sim.graphVerticesDevice(colors, vbo.data_ptr(), ne - 1)
    # a step moves the particles: the edge slots are stale
    sim.simulate()
    with pytest.raises(cf.CellFlowError):
        sim.graphVerticesDevice(colors)
    # a rebuilt cell list alone (no step) also invalidates them: the slots were reordered
    edges2, verts2 = sim.generateProximityGraph(180.0, 5, colors)
    sim.moveUniverse(10.0, 0.0, 0.0)
    sim.cellKeys()
    with pytest.raises(cf.CellFlowError):
        sim.graphVerticesDevice(colors)
    sim.close()


def test_uploaded_types_are_clamped():
    """ptype == numTypes (the reference's own spawn can produce it: curand_uniform may return 1.0, .cu:62) or
    larger is clamped to T-1 on upload instead of indexing the tables out of bounds."""
    sim, p, table, radio, state, counts = small_sim(n=20000)
    bad = state.copy()
    bad["ptype"][::97] = 6
    bad["ptype"][5::101] = 250
    sim.setParticleData(bad, counts)
    sim.simulate()
    gcnt = sim.getNeighborCounts()
    fixed = bad.copy()
    fixed["ptype"] = np.minimum(bad["ptype"], 5)
    _, wcnt, _ = O.step(fixed, counts, p, table, radio, "cells", THREADS)
    assert np.array_equal(gcnt, wcnt)
    assert np.array_equal(sim.getParticleData()["ptype"], fixed["ptype"])
    sim.close()


def test_table_draws_continue_one_rand_stream():
    """The reference never seeds libc's rand(): construction draws T*T + T values, every later
    regenerateForceTable / setNumParticleTypes continues the same sequence (.cu:513-539, 574-583)."""
    libc = C.CDLL("libc.so.6")
    libc.srand(1)  # the never-seeded state
    draw = lambda k: np.float32([np.float32(np.float32(libc.rand()) / np.float32(2147483647)) * np.float32(2.0) - np.float32(1.0)
                                 for _ in range(k)])
    sim = cf.ParticleSimulation(16, 6)
    raw0, radio0 = draw(36), draw(6)
    assert np.array_equal(sim.getRawForceTableValues()[:36], raw0) and np.array_equal(sim.getRadioByType(), radio0)
    sim.regenerateForceTable()
    assert np.array_equal(sim.getRawForceTableValues()[:36], draw(36))
    sim.setNumParticleTypes(4)
    assert np.array_equal(sim.getRawForceTableValues()[:16], draw(16)) and np.array_equal(sim.getRadioByType(), draw(4))
    sim.close()
    # a second handle starts its own copy of the sequence, like a second process would
    sim2 = cf.ParticleSimulation(16, 6)
    assert np.array_equal(sim2.getRawForceTableValues()[:36], raw0)
    sim2.close()


def test_block_counters_of_the_tile_kernel():
    """Option "count_blocks": the instrumented tile kernel reports how many pairs reached the exact test and how
    many pair-lanes were evaluated; results are those of the plain kernel."""
    sim, p, table, radio, state, counts = small_sim(n=60000, kernel=3)
    sim.simulate()
    plain, pc = sim.getParticleData(), sim.getNeighborCounts()
    sim.setParticleData(state, counts)
    sim.setOption("count_blocks", 1)
    sim.simulate()
    inst, ic = sim.getParticleData(), sim.getNeighborCounts()
    st = sim.stats()
    assert np.array_equal(pc, ic) and plain.tobytes() == inst.tobytes()
    acc = int(ic.sum())
    assert st.force_kernel == 3
    # every accepted pair (and every particle's own slot) sits in an evaluated block; every evaluated block was tested.
    # (Blocks count 128 pair-lanes each, padded lanes included, so they may exceed the stencil's pair count.)
    assert st.evaluated_pair_lanes >= acc + len(state) and st.exact_tested_pairs >= st.evaluated_pair_lanes
    sim.close()


@pytest.mark.parametrize("stage", [0, 1, 2, 3])
def test_alternative_chunk_staging_is_bit_identical(stage):
    """Option "t4_stage": 1 = the j chunks of the type-sorted copy arrive through cp.async.bulk + mbarrier from SoA
    planes; 2 = quad bounding boxes precomputed once per step, positions loaded for surviving quads only; 3 = SoA
    planes through registers (no transposition, mask-free fast path); 0 = every quad stored, prefilter afterwards (the
    default until the middle of round 2).  The default (4) runs the box prefilter on the registers and stores the live
    quads compacted.  1-3 align
    chunk starts down to 4 elements and mask foreign elements.  Same pairs in the same order: counts and forces are
    bit-identical to the default staging, and match the oracle."""
    sim, p, table, radio, state, counts = small_sim(n=60000, kernel=3)
    sim.simulate(steps=2)
    a, ac = sim.getParticleData(), sim.getNeighborCounts()
    sim.setParticleData(state, counts)
    sim.setOption("t4_stage", stage)
    sim.simulate(steps=2)
    b, bc = sim.getParticleData(), sim.getNeighborCounts()
    assert sim.stats().force_kernel == 3
    assert np.array_equal(ac, bc) and a.tobytes() == b.tobytes()
    sim.setParticleData(state, counts)
    sim.simulate()
    got, gcnt = sim.getParticleData(), sim.getNeighborCounts()
    want, wcnt, fabs = O.step(state, counts, p, table, radio, "cells", THREADS)
    assert np.array_equal(gcnt, wcnt)
    assert U.force_rel_err(got["acc"], want["acc"], fabs, U.force_multiplier_of(p, wcnt, counts)).max() <= U.FORCE_RTOL
    sim.close()
    if stage in (0, 2):   # uniform radius (the other instantiation) + a slab-mode run with ghost slots
        p2, table2, radio2 = U.config("pulser", canvasWidth=3000.0, canvasHeight=3000.0, canvasDepth=3000.0)
        st2, c2 = U.random_state(50000, 6, 19, p2.canvas, "uniform")
        for slab in (False, True):
            sim = cf.ParticleSimulation(0 if slab else len(st2), 6, init=False)
            sim.params = U.to_lib_params(p2)
            sim.setRadioByType(radio2)
            sim.setForceTable(table2)
            sim.setOption("force_kernel", 3)
            sim.setOption("t4_stage", stage)
            if slab:
                sim.commInit(0, 1, 70000)
                sim.uploadOwned(st2, c2, np.arange(len(st2), dtype=np.int32))
            else:
                sim.setParticleData(st2, c2)
            sim.simulate()
            if slab:
                pp, cc, ii = sim.downloadOwned()
                gcnt = np.zeros(len(st2), np.int32)
                gcnt[ii] = cc
            else:
                gcnt = sim.getNeighborCounts()
            _, wcnt, _ = O.step(st2, c2, p2, table2, radio2, "cells", THREADS)
            assert np.array_equal(gcnt, wcnt), f"slab={slab}"
            sim.close()
