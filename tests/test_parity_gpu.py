"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the
same inputs.  Bars (BASELINE.json north_star): neighbour counts, cell assignment and graph edge
sets bit-exact; forces within 1e-5 of the summed pair-term magnitude; positions within the
position quantum the reference's own fmodf(pos + W, W) imposes."""
import numpy as np
import pytest

import cellflow_b200 as cf
import oracle as O
import util as U

pytestmark = pytest.mark.gpu

THREADS = max(1, min(8, O.max_threads()))


def make_sim(p, table, radio, state, counts, **opts):
    sim = cf.ParticleSimulation(len(state), p.numParticleTypes, init=False)
    sim.params = U.to_lib_params(p)
    sim.setRadioByType(radio)
    sim.setForceTable(table)
    for k, v in opts.items():
        sim.setOption(k, v)
    sim.setParticleData(state, counts)
    return sim


def check_step(p, table, radio, state, counts, **opts):
    sim = make_sim(p, table, radio, state, counts, **opts)
    sim.simulate()
    got = sim.getParticleData()
    gcnt = sim.getNeighborCounts()
    want, wcnt, fabs = O.step(state, counts, p, table, radio, "cells", THREADS)
    # 1. neighbour counts: bit-exact, no exclusions needed (same rounding points by construction)
    assert np.array_equal(gcnt, wcnt), f"{(gcnt != wcnt).sum()} of {len(wcnt)} counts differ"
    # 2. forces (p.acc, .cu:146): 1e-5 of the summed |pair term|
    mult = U.force_multiplier_of(p, wcnt, counts)
    rel = U.force_rel_err(got["acc"], want["acc"], fabs, mult)
    assert rel.max() <= U.FORCE_RTOL, f"force rel err {rel.max():.3e}"
    # 3. velocity: v' = v*friction + acc*dt -> error is dt * force error
    dv = np.abs(got["vel"].astype(np.float64) - want["vel"]).max(axis=1)
    scale = p.delta_t * np.abs(mult) * fabs.astype(np.float64)
    assert np.all(dv <= U.FORCE_RTOL * scale + 1e-6 * np.abs(want["vel"]).max(axis=1) + 1e-30)
    # 4. position: within dt * velocity error plus one quantum of fmodf(pos + W, W)
    canvas = p.canvas
    dp = U.wrapped_abs_diff(got["pos"], want["pos"], canvas).max(axis=1)
    quantum = float(np.spacing(np.float32(2 * canvas.max())))
    assert np.all(dp <= p.delta_t * p.delta_t * U.FORCE_RTOL * np.abs(mult) * fabs + 1.01 * quantum), dp.max()
    assert np.array_equal(got["ptype"], state["ptype"])
    sim.close()
    return rel.max(), gcnt


CASES = [
    # name, n, mode, overrides, radio
    ("settings", 10000, "cube", {}, None),                      # BASELINE config 1 shape
    ("eater", 20000, "uniform", {}, None),
    ("pulser", 30000, "cube", {}, None),
    ("littlecells", 20000, "blobs", {}, None),
    ("eater", 20000, "uniform", {"ratioWithLFO": 0.5}, [1.0, 0.5, 0.0, 0.0, -0.5, 1.0]),   # config 3 radii
    ("eater", 6000, "uniform", {"ratioWithLFO": 1.0, "radius": 900.0}, [1.0, 0.5, 0.0, 0.0, -0.5, 1.0]),
    ("settings", 5000, "uniform", {"radius": 3000.0}, None),   # R > W/3: one or two cells per axis
    ("settings", 5000, "uniform", {"radius": 0.5}, None),      # nobody interacts
    ("eater", 4000, "uniform", {"canvasWidth": 4000.0, "canvasHeight": 3000.0, "canvasDepth": 1500.0}, None),
    ("pulser", 5000, "cube", {"balance": 1.4, "maxExpectedNeighbors": 50}, None),   # balance > 1
]


@pytest.mark.parametrize("name,n,mode,over,radio", CASES)
@pytest.mark.parametrize("kernel", [1, 2, 3])
def test_step_parity(name, n, mode, over, radio, kernel):
    p, table, r0 = U.config(name, **over)
    radio = np.float32(radio) if radio is not None else r0
    state, counts = U.random_state(n, p.numParticleTypes, 42 + n, p.canvas, mode)
    check_step(p, table, radio, state, counts, force_kernel=kernel)


@pytest.mark.parametrize("kernel", [1, 2, 3])
def test_ten_types_random_radii(kernel):
    """MAX_PARTICLE_TYPES = 10 (SimulationParams.h:64) with a different radius modifier per type and
    a dense state: 180 (run, type) sub-runs per tile in the generation-4 kernel, asymmetric 10x10
    force matrix."""
    rng = np.random.default_rng(2024)
    T = 10
    p, _, _ = U.config("eater", numParticleTypes=T, ratioWithLFO=0.8, canvasWidth=3000.0, canvasHeight=2600.0,
                       canvasDepth=2200.0, radius=180.0)
    table = O.force_table(rng.uniform(-1, 1, T * T).astype(np.float32), T, p.forceRange, p.forceBias, p.forceOffset)
    radio = rng.uniform(-0.6, 1.0, T).astype(np.float32)
    state, counts = U.random_state(60000, T, 7, p.canvas, "uniform")
    check_step(p, table, radio, state, counts, force_kernel=kernel)


def test_auto_kernel_choice_dense_mixed_radii():
    """No force_kernel option: a dense state with per-type radii takes the generation-4 tile kernel
    (type-sorted j copy), and the result still matches the oracle."""
    p, table, _ = U.config("eater", ratioWithLFO=0.5, canvasWidth=2400.0, canvasHeight=2400.0, canvasDepth=2400.0)
    radio = np.float32([1.0, 0.5, 0.0, 0.0, -0.5, 1.0])
    state, counts = U.random_state(40000, 6, 11, p.canvas, "uniform")
    sim = make_sim(p, table, radio, state, counts)
    sim.simulate()
    assert sim.stats().force_kernel == 3
    sim.close()
    check_step(p, table, radio, state, counts)


def test_default_parameters_and_matrix():
    """README headline configuration: SimulationParams.h defaults, default (glibc rand) matrix."""
    p = O.Params()
    raw, radio, eff = cf.reference_default_tables(6)
    state, counts = U.random_state(20000, 6, 5, p.canvas, "cube")
    sim = cf.ParticleSimulation(20000, 6, init=False)  # fresh object: default tables inside
    assert np.array_equal(sim.getForceTable(), eff) and np.array_equal(sim.getRadioByType(), radio)
    sim.close()
    check_step(p, eff, radio, state, counts)
    p2 = O.Params(ratioWithLFO=1.5)  # default radio vector is non-zero: ratio matters
    check_step(p2, eff, radio, state, counts)


@pytest.mark.parametrize("n", [1, 2, 31, 33, 257])
def test_tiny_counts(n):
    p, table, radio = U.config("eater")
    state, counts = U.random_state(n, 6, n, p.canvas, "cube", cube=600.0)
    check_step(p, table, radio, state, counts)


def test_coincident_and_boundary_particles():
    p, table, radio = U.config("settings")
    state, counts = U.random_state(4000, 8, 9, p.canvas, "cube")
    state["pos"][:50] = state["pos"][50]            # coincident: dist = sqrt(1e-4), counted
    top = np.nextafter(p.canvas, np.float32(0))
    state["pos"][100:140] = 0.0                     # on the periodic seam
    state["pos"][140:180] = top
    state["pos"][180:200, 0] = top[0]
    state["pos"][180:200, 1] = 0.0
    check_step(p, table, radio, state, counts)


def test_multi_step_tracks_oracle_per_step():
    """The system is chaotic, so parity is per step: every step restarts the oracle from the
    engine's own state."""
    p, table, radio = U.config("pulser")
    state, counts = U.random_state(8000, 6, 77, p.canvas, "cube")
    sim = make_sim(p, table, radio, state, counts)
    for step in range(6):
        before = sim.getParticleData()
        cprev = sim.getNeighborCounts()
        sim.simulate()
        got, gcnt = sim.getParticleData(), sim.getNeighborCounts()
        want, wcnt, fabs = O.step(before, cprev, p, table, radio, "cells", THREADS)
        assert np.array_equal(gcnt, wcnt), step
        mult = U.force_multiplier_of(p, wcnt, cprev)
        assert U.force_rel_err(got["acc"], want["acc"], fabs, mult).max() <= U.FORCE_RTOL
    sim.close()


def test_clustered_state_switches_to_the_tile_kernel():
    """Blobs: most cells are empty (7 particles per cell on average -> the per-particle kernel on the
    first step) but most particles sit in cells with hundreds.  The occupancy measured by a step
    (sum n_c^2 / n, taken over at the next synchronising call) moves the following steps to the
    generation-4 tile kernel; every step still matches the oracle."""
    p, table, radio = U.config("eater")
    state, counts = U.random_state(60000, 6, 12, p.canvas, "blobs")
    sim = make_sim(p, table, radio, state, counts)
    used = []
    for step in range(3):
        before = sim.getParticleData()
        cprev = sim.getNeighborCounts()
        sim.simulate()
        used.append(sim.stats().force_kernel)
        got, gcnt = sim.getParticleData(), sim.getNeighborCounts()
        want, wcnt, fabs = O.step(before, cprev, p, table, radio, "cells", THREADS)
        assert np.array_equal(gcnt, wcnt), step
        mult = U.force_multiplier_of(p, wcnt, cprev)
        assert U.force_rel_err(got["acc"], want["acc"], fabs, mult).max() <= U.FORCE_RTOL
    assert used[0] == 1 and used[-1] == 3, used
    sim.close()


def test_simulate_n_steps_equals_n_calls():
    p, table, radio = U.config("eater")
    state, counts = U.random_state(5000, 6, 3, p.canvas, "cube")
    a = make_sim(p, table, radio, state, counts)
    b = make_sim(p, table, radio, state, counts)
    a.simulate(steps=5)
    for _ in range(5):
        b.simulate()
    assert a.getParticleData().tobytes() == b.getParticleData().tobytes()   # deterministic
    assert np.array_equal(a.getNeighborCounts(), b.getNeighborCounts())
    a.close(), b.close()


def test_step_host_equals_resident_step():
    p, table, radio = U.config("settings")
    state, counts = U.random_state(6000, 8, 4, p.canvas, "cube")
    a = make_sim(p, table, radio, state, counts)
    a.simulate()
    b = make_sim(p, table, radio, state, counts)
    out, cout = b.stepHost(state, counts)
    assert out.tobytes() == a.getParticleData().tobytes()
    assert np.array_equal(cout, a.getNeighborCounts())
    a.close(), b.close()


def test_cell_assignment_bit_exact():
    p, table, radio = U.config("settings")
    state, counts = U.random_state(50000, 8, 8, p.canvas, "uniform")
    sim = make_sim(p, table, radio, state, counts)
    keys, ids = sim.cellKeys()
    dims = np.int32(list(sim.stats().grid))
    # key = cell*64 + Hilbert index of the 4x4x4 sub-cell
    inv = dims.astype(np.float32) / p.canvas
    y = (state["pos"] * inv).astype(np.float32)
    cell3 = np.minimum(y.astype(np.int32), dims - 1)
    sub = np.clip(np.minimum((y * np.float32(4)).astype(np.int32), 4 * dims - 1) - 4 * cell3, 0, 3).astype(np.uint32)
    want = O.cell_keys(state, p.canvas, dims) * np.uint32(64) + U.hilbert64(sub[:, 0], sub[:, 1], sub[:, 2])
    assert sorted(ids.tolist()) == list(range(len(state)))          # a permutation
    assert np.array_equal(keys, want[ids])                          # key of every slot
    assert np.all(np.diff(keys.astype(np.int64)) >= 0)              # sorted
    same = keys[1:] == keys[:-1]
    assert np.all(ids[1:][same] > ids[:-1][same])                   # stable inside a sub-cell
    sim.close()


@pytest.mark.parametrize("mode,n,dist,mc", [("cube", 20000, 200.0, 5), ("uniform", 30000, 500.0, 3),
                                            ("blobs", 20000, 120.0, 16), ("cube", 5000, 200.0, 20)])
@pytest.mark.parametrize("gkernel", [1, 2])
def test_graph_edge_set_bit_exact(mode, n, dist, mc, gkernel):
    """gkernel 1 = thread per particle, 2 = warp per particle (dense states)."""
    p, table, radio = U.config("eater")
    state, counts = U.random_state(n, 6, 21, p.canvas, mode)
    sim = make_sim(p, table, radio, state, counts, graph_kernel=gkernel)
    colors = np.zeros(10, cf.COLOR)
    colors["r"], colors["g"], colors["b"] = np.arange(10) * 0.1, 0.5, 1.0 - np.arange(10) * 0.1
    edges, verts = sim.generateProximityGraph(dist, mc, colors)
    want = O.graph(state, dist, mc, canvas=p.canvas, method="cells")
    assert len(edges) == len(want) and len(want) > 0
    assert U.edge_set(edges) == U.edge_set(want)
    # reference VBO layout, compared as a set of 12-float records
    wv = O.graph_vertices(state, edges, colors, 6)
    assert np.array_equal(verts, wv)
    sim.close()


def test_graph_dense_cluster_and_ties():
    """A tight cluster (hundreds of same-type candidates per particle, many of them at exactly the
    same distance) through both graph kernels and the automatic choice: the edge set is the rule's."""
    p, table, radio = U.config("eater")
    state, counts = U.random_state(6000, 6, 33, p.canvas, "cube", cube=300.0)
    state["pos"] = np.round(state["pos"] / 25.0).astype(np.float32) * 25.0   # lattice: equal distances
    want = O.graph(state, 200.0, 5, canvas=p.canvas, method="cells")
    for gk in (1, 2):
        sim = make_sim(p, table, radio, state, counts, graph_kernel=gk)
        edges, _ = sim.generateProximityGraph(200.0, 5)
        assert U.edge_set(edges) == U.edge_set(want), f"graph_kernel={gk}"
        sim.close()
    sim = make_sim(p, table, radio, state, counts)   # automatic choice
    for call in range(2):   # the second call has seen the occupancy of the first
        edges, _ = sim.generateProximityGraph(200.0, 5)
        assert U.edge_set(edges) == U.edge_set(want), f"automatic, call {call}"
    sim.close()


def test_graph_after_steps_uses_current_positions():
    p, table, radio = U.config("pulser")
    state, counts = U.random_state(10000, 6, 2, p.canvas, "cube")
    sim = make_sim(p, table, radio, state, counts)
    sim.simulate(steps=3)
    now = sim.getParticleData()
    edges, _ = sim.generateProximityGraph(200.0, 5)
    assert U.edge_set(edges) == U.edge_set(O.graph(now, 200.0, 5, canvas=p.canvas, method="cells"))
    sim.close()


def test_graph_replayed_cuda_graphs_track_the_state():
    """step + graph, ten times: from the third time on both the step and the graph build replay
    captured CUDA graphs (two ping-pong parities each); every build must still be the rule applied
    to the positions of that moment, and a repeated build without a step in between as well."""
    p, table, radio = U.config("pulser")
    state, counts = U.random_state(6000, 6, 4, p.canvas, "cube")
    sim = make_sim(p, table, radio, state, counts)
    for it in range(10):
        sim.simulate()
        edges, _ = sim.generateProximityGraph(200.0, 5)
        now = sim.getParticleData()
        assert U.edge_set(edges) == U.edge_set(O.graph(now, 200.0, 5, canvas=p.canvas, method="cells")), it
        if it % 4 == 3:
            again, _ = sim.generateProximityGraph(200.0, 5)
            assert U.edge_set(again) == U.edge_set(edges)
    sim.close()


def test_init_particles_matches_oracle():
    p, table, radio = U.config("eater")
    for mode in (cf.INIT_SPAWN_CUBE, cf.INIT_UNIFORM):
        sim = cf.ParticleSimulation(12345, 6, init=False)
        sim.params = U.to_lib_params(p)
        sim.initializeParticles(seed=0x5EED0003, mode=mode)
        got = sim.getParticleData()
        want = O.init_particles(12345, 6, 0x5EED0003, mode, p.canvas)
        assert got.tobytes() == want.tobytes()
        sim.close()


def test_move_universe_and_rotate():
    p, table, radio = U.config("eater")
    state, counts = U.random_state(3000, 6, 6, p.canvas)
    sim = make_sim(p, table, np.float32([.1, .2, .3, .4, .5, .6]), state, counts)
    sim.moveUniverse(123.5, -77.25)
    assert np.array_equal(sim.getParticleData()["pos"], O.move_universe(state, 123.5, -77.25, 0.0, p.canvas)["pos"])
    sim.rotateRadioByType()
    assert np.array_equal(sim.getRadioByType(), np.float32([.6, .1, .2, .3, .4, .5]))  # .cu:594-600
    sim.setRadioByTypeValue(99, 1.0)   # out of range: ignored like the reference (.cu:616)
    sim.close()


def test_apply_preset_matches_oracle_tables():
    pr = cf.load_preset(U.PRESETS + "/pulser.json")
    sim = cf.ParticleSimulation(100, 6)
    sim.applyPreset(pr)
    p, table, radio = U.config("pulser")
    assert sim.getParticleCount() == 30000
    assert np.array_equal(sim.getForceTable(), table)
    assert np.array_equal(sim.getRadioByType(), radio)
    sim.simulate(steps=2)
    assert np.isfinite(sim.getParticleData()["pos"]).all()
    sim.close()


def test_large_uniform_matches_oracle_1m():
    """BASELINE scale: 1,000,000 particles, settings.json law, uniform 8000^3 (~189 neighbours)."""
    p, table, radio = U.config("settings")
    n = 1_000_000
    state = O.init_particles(n, 8, 0x5EED0003, 1, p.canvas)
    counts = np.zeros(n, np.int32)
    sim = make_sim(p, table, radio, state, counts)
    sim.simulate()
    got, gcnt = sim.getParticleData(), sim.getNeighborCounts()
    O.set_sort_candidates(False)   # tolerance-based comparison: summation order is free
    try:
        want, wcnt, fabs = O.step(state, counts, p, table, radio, "cells", THREADS)
    finally:
        O.set_sort_candidates(True)
    assert np.array_equal(gcnt, wcnt)
    assert 150 < wcnt.mean() < 230
    mult = U.force_multiplier_of(p, wcnt, counts)
    assert U.force_rel_err(got["acc"], want["acc"], fabs, mult).max() <= U.FORCE_RTOL
    # size-independent property: uniform radius -> symmetric neighbour relation
    assert int(gcnt.astype(np.int64).sum()) % 2 == 0
    sim.close()


def test_newton_third_law_with_symmetric_matrix():
    """With a symmetric force matrix and uniform radius the pair terms cancel exactly in real
    arithmetic: the net force of the whole system vanishes relative to the summed magnitudes."""
    p, table, radio = U.config("eater")
    T = 6
    sym = ((table.reshape(T, T) + table.reshape(T, T).T) * 0.5).astype(np.float32).ravel()
    n = 200_000
    state = O.init_particles(n, T, 11, 0, p.canvas)
    p.maxExpectedNeighbors = 10 ** 9   # density factor ~ 0 -> uniform multiplier
    sim = make_sim(p, sym, radio, state, np.zeros(n, np.int32))
    sim.simulate()
    acc = sim.getParticleData()["acc"].astype(np.float64)
    assert np.abs(acc.sum(0)).max() <= 1e-5 * np.abs(acc).sum(0).max()
    sim.close()


@pytest.mark.parametrize("tag", ["settings", "eater_radii_wrap", "defaults", "pulser_edge"])
@pytest.mark.parametrize("kernel", [1, 2, 3])
def test_against_reference_kernel_golden(tag, kernel):
    """The engine against the REFERENCE'S OWN kernel output (tests/golden, generated on a B200 by
    the unmodified reference .cu one warp at a time): neighbour counts bit-exact, forces 1e-5."""
    import ctypes as C
    import os
    z = np.load(os.path.join(U.GOLDEN, f"step_{tag}.npz"))
    p = O.Params()
    C.memmove(C.byref(p), z["params"].tobytes(), C.sizeof(p))
    state, counts = z["state"].view(O.PARTICLE).reshape(-1), z["counts"]
    ref, ref_cnt = z["out"].view(O.PARTICLE).reshape(-1), z["cnt"]
    sim = make_sim(p, z["table"], z["radio"], state, counts, force_kernel=kernel)
    sim.simulate()
    got, gcnt = sim.getParticleData(), sim.getNeighborCounts()
    assert np.array_equal(gcnt, ref_cnt)
    _, _, fabs = O.step(state, counts, p, z["table"], z["radio"], "cells", THREADS)   # scale only
    mult = U.force_multiplier_of(p, ref_cnt, counts)
    assert U.force_rel_err(got["acc"], ref["acc"], fabs, mult).max() <= U.FORCE_RTOL
    dp = U.wrapped_abs_diff(got["pos"], ref["pos"], p.canvas).max(axis=1)
    quantum = float(np.spacing(np.float32(2 * p.canvas.max())))
    assert np.all(dp <= p.delta_t * p.delta_t * U.FORCE_RTOL * np.abs(mult) * fabs + 1.01 * quantum)
    sim.close()


def test_graph_against_reference_kernel_golden():
    import os
    for tag in ("cube", "blobs"):
        z = np.load(os.path.join(U.GOLDEN, f"graph_{tag}.npz"))
        state = z["state"].view(O.PARTICLE).reshape(-1)
        colors = z["colors"].view(cf.COLOR).reshape(-1)
        p, table, radio = U.config("eater")
        sim = make_sim(p, table, radio, state, np.zeros(len(state), np.int32))
        edges, verts = sim.generateProximityGraph(float(z["dist"]), int(z["max_conn"]), colors)
        rec = verts[np.lexsort(verts.T[::-1])]
        assert np.array_equal(rec, z["records"])      # the reference's VBO content, as a set
        sim.close()


def test_empty_and_degenerate_inputs():
    p, table, radio = U.config("eater")
    sim = cf.ParticleSimulation(0, 6, init=False)       # no particles at all
    sim.params = U.to_lib_params(p)
    sim.simulate(steps=3)
    assert len(sim.getParticleData()) == 0
    edges, _ = sim.generateProximityGraph(200.0, 5)
    assert len(edges) == 0
    sim.close()
    # one type only, everything in one cell, zero velocity
    p1, t1, r1 = U.config("eater", numParticleTypes=1)
    state, counts = U.random_state(500, 1, 3, p1.canvas, "cube", cube=100.0, vel_scale=0.0)
    check_step(p1, np.float32([0.3]), np.float32([0.0]), state, counts)
    # maxConn 0 and distance 0: no edges, no error
    sim = make_sim(p, table, radio, *U.random_state(1000, 6, 1, p.canvas, "cube"))
    assert len(sim.generateProximityGraph(200.0, 0)[0]) == 0
    assert len(sim.generateProximityGraph(0.0, 5)[0]) == 0
    with pytest.raises(cf.CellFlowError):
        sim.setParticleData(np.zeros(10, cf.PARTICLE)[:0], None) if False else sim.setOption("nope", 1)
    sim.close()


def test_render_feed_matches_widget_packing():
    """(x, y, z, float(type)) in particle order + per-type counts, as CellFlowWidget builds them
    on the CPU from getParticleData (CellFlowWidget.cpp:742-761, 875-886)."""
    p, table, radio = U.config("pulser")
    state, counts = U.random_state(7777, 6, 12, p.canvas, "cube")
    sim = make_sim(p, table, radio, state, counts)
    sim.simulate(steps=2)
    now = sim.getParticleData()
    xyzt, tc = sim.getRenderFeed()
    assert np.array_equal(xyzt[:, :3], now["pos"]) and np.array_equal(xyzt[:, 3], now["ptype"].astype(np.float32))
    assert np.array_equal(tc, np.bincount(now["ptype"], minlength=6))
    sim.close()
