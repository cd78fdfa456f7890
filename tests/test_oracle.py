"""CPU tests of the oracle itself: known-answer values derived from the reference formulas
(SURVEY.md section 8a), self-consistency of its two evaluation orders, fp64 error budget."""
import numpy as np
import pytest

import oracle as O
import util as U


def test_default_tables_known_answer():
    # glibc rand() unseeded, T=6 (ParticleSimulation.cu:513-519, 533-539) — SURVEY.md 8a
    raw, radio = O.default_tables(6)
    np.testing.assert_allclose(
        raw[:6], [0.680375457, -0.211234152, 0.566198468, 0.596880078, 0.823294759, -0.604897261],
        rtol=0, atol=1e-8)
    np.testing.assert_allclose(
        radio, [-0.0128340125, 0.945550084, -0.414966404, 0.54271543, 0.0534899235, 0.539827704],
        rtol=0, atol=1e-8)


def test_force_table_known_answers():
    # littlecells / eater / pulser share one raw table and (0.33, -0.185, 0.79)
    rows = {}
    for name in ("littlecells", "eater", "pulser"):
        p, table, _ = U.config(name)
        rows[name] = table
        np.testing.assert_allclose(
            table[:6], [-0.0273239, -0.2714024, -0.3182676, -0.0027768, -0.2616002, -0.0903976],
            atol=2e-7)
        assert abs(table.min() - -0.3518) < 1e-4 and abs(table.max() - 0.0188) < 1e-4
    assert np.array_equal(rows["eater"], rows["pulser"]) and np.array_equal(rows["eater"], rows["littlecells"])
    p, table, _ = U.config("settings")
    assert p.numParticleTypes == 8
    np.testing.assert_allclose(
        table[:8], [-0.1910413, -0.1846015, -0.1326903, -0.1042395, -0.1795201, -0.1827365,
                    -0.2010488, -0.1293678], atol=2e-7)
    assert np.all(np.abs(table) <= 1.0)


def test_force_table_clamp():
    raw = np.array([5.0, -5.0, 0.0, 0.3], np.float32)
    t = O.force_table(raw, 2, 3.0, 0.5, 1.0)
    assert t[0] == 1.0 and t[1] == -1.0
    assert t[2] == np.float32(0.5)


def test_lfo():
    p = O.Params(ratio=0.25, lfoA=0.0, lfoS=3.0)
    assert O.ratio_with_lfo(p, 1.234) == np.float32(0.25)
    p = O.Params(ratio=0.25, lfoA=0.5, lfoS=0.1)
    t = np.float32(2.5)
    want = np.float32(np.float64(np.float32(0.25)) + np.float64(np.float32(0.5)) * np.sin(
        2.0 * np.pi * np.float64(np.float32(0.1)) * np.float64(t)))
    assert O.ratio_with_lfo(p, float(t)) == want


@pytest.mark.parametrize("name,n,mode", [("settings", 3000, "cube"), ("eater", 2500, "uniform"),
                                         ("littlecells", 2000, "blobs")])
def test_bruteforce_equals_cells(name, n, mode):
    p, table, radio = U.config(name)
    state, counts = U.random_state(n, p.numParticleTypes, 7, p.canvas, mode)
    a, ca, fa = O.step(state, counts, p, table, radio, "brute", 4)
    b, cb, fb = O.step(state, counts, p, table, radio, "cells", 4)
    assert a.tobytes() == b.tobytes()
    assert np.array_equal(ca, cb) and np.array_equal(fa, fb)
    assert ca.sum() > 0


def test_nonuniform_radius_and_degenerate_grid():
    # Reff up to 2*radius with radius ~ W/5: fewer than 3 oracle cells per axis
    p, table, radio = U.config("eater", radius=900.0, ratioWithLFO=1.0, canvasWidth=4000.0,
                               canvasHeight=3000.0, canvasDepth=2500.0)
    radio = np.array([1.0, 0.5, 0.0, 0.0, -0.5, 1.0], np.float32)
    state, counts = U.random_state(600, 6, 3, p.canvas)
    a, ca, _ = O.step(state, counts, p, table, radio, "brute", 4)
    b, cb, _ = O.step(state, counts, p, table, radio, "cells", 4)
    assert a.tobytes() == b.tobytes() and np.array_equal(ca, cb)
    reff = O.reff_table(p, radio).reshape(6, 6)
    assert reff[0, 0] == np.float32(1800.0) and reff[4, 4] == np.float32(450.0)
    assert np.array_equal(reff, reff.T)


def test_step_invariants():
    p, table, radio = U.config("pulser")
    state, counts = U.random_state(2000, 6, 11, p.canvas, "cube")
    out, cnt, fabs = O.step(state, counts, p, table, radio, "cells", 4)
    canvas = p.canvas
    assert np.all(out["pos"] >= 0) and np.all(out["pos"] < canvas)
    assert np.array_equal(out["ptype"], state["ptype"])
    # uniform radius -> the neighbour relation is symmetric -> sum of counts is even
    assert cnt.sum() % 2 == 0
    # zero force table -> pure friction drift
    zt = np.zeros_like(table)
    out0, cnt0, _ = O.step(state, counts, p, zt, radio, "cells", 4)
    assert np.array_equal(cnt0, cnt)
    v = (state["vel"] * np.float32(p.friction)).astype(np.float32)
    assert np.array_equal(out0["vel"], v)
    assert np.all(out0["acc"] == 0)


def test_f64_error_budget():
    """fp32 transcription vs fp64 restatement: the law itself is well inside 1e-5."""
    p, table, radio = U.config("settings")
    state, counts = U.random_state(4000, 8, 5, p.canvas, "cube")
    out, cnt, fabs = O.step(state, counts, p, table, radio, "cells", 4)
    F, X, fa = O.step_f64(state, counts, p, table, radio, 4)
    err = np.abs(out["acc"] - F).max(1) / (fa + 1e-30)
    assert cnt.mean() > 30
    assert err.max() < 1e-6
    # sparse neighbourhoods (struct defaults, ~0.8 neighbours): still fine under the gross norm,
    # while relative to the NET pair force the reference's own fp32 arithmetic exceeds 1e-5
    p2 = O.Params()
    raw, radio2 = O.default_tables(6)
    eff = O.force_table(raw, 6, 0.28, -0.20, 1.0)
    s2, c2 = U.random_state(20000, 6, 5, p2.canvas, "cube")
    o2, n2, g2 = O.step(s2, c2, p2, eff, radio2, "cells", 4)
    F2, _, ga2 = O.step_f64(s2, c2, p2, eff, radio2, 4)
    assert (np.abs(o2["acc"] - F2).max(1) / (ga2 + 1e-30)).max() < 1e-6
    d = U.wrapped_abs_diff(out["pos"], np.mod(X, p.canvas), p.canvas)
    assert d.max() < 2e-3  # ulp(pos + W) = 9.8e-4 at W = 8000


@pytest.mark.parametrize("mode,n", [("cube", 3000), ("uniform", 2000), ("blobs", 3000)])
def test_graph_bruteforce_equals_cells(mode, n):
    p, _, _ = U.config("eater")
    state, _ = U.random_state(n, 6, 13, p.canvas, mode)
    for dist, mc in ((200.0, 5), (500.0, 2), (120.0, 16)):
        a = O.graph(state, dist, mc, method="brute")
        b = O.graph(state, dist, mc, canvas=p.canvas, method="cells")
        assert np.array_equal(a, b)
        if len(a):
            assert np.all(a["i"] < a["j"])
            assert np.all(state["ptype"][a["i"]] == state["ptype"][a["j"]])
            assert np.bincount(a["i"]).max() <= mc


def test_graph_rule_details():
    # 1-D line of same-type particles: candidates are the first 2*maxConn by INDEX, then nearest
    n = 12
    s = O.particles(n)
    s["pos"][:, 0] = [0, 50, 10, 40, 20, 30, 5, 45, 15, 35, 25, 60]
    s["pos"][:, 1:] = 100.0
    e = O.graph(s, 100.0, 2, method="brute")
    mine = e[e["i"] == 0]
    # first 4 in-range by index: 1(50),2(10),3(40),4(20) -> nearest two: 2, 4
    assert mine["j"].tolist() == [2, 4]
    # non-wrapped distance: particles at 1 and 7999 are NOT neighbours
    s2 = O.particles(2)
    s2["pos"][0] = (1, 1, 1)
    s2["pos"][1] = (7999, 1, 1)
    assert len(O.graph(s2, 200.0, 5, method="brute")) == 0
    # maxConn above 16 is clamped (the reference overflows nearby[32] there)
    s3 = O.particles(40)
    s3["pos"][:] = 10.0
    assert np.bincount(O.graph(s3, 50.0, 20, method="brute")["i"]).max() == 16


def test_move_universe():
    p, _, _ = U.config("eater")
    state, _ = U.random_state(100, 6, 1, p.canvas)
    out = O.move_universe(state, 123.5, -77.25, 0.0, p.canvas)
    want = np.fmod((state["pos"] + np.float32([123.5, -77.25, 0.0])).astype(np.float32) + p.canvas, p.canvas)
    assert np.array_equal(out["pos"], want.astype(np.float32))


def test_init_particles_shape():
    canvas = np.float32([8000, 8000, 8000])
    a = O.init_particles(20000, 6, 0x5EED0002, 0, canvas)
    assert a["pos"].min() >= 3000 and a["pos"].max() < 5000  # centred 2000^3 cube, .cu:45-55
    assert set(np.unique(a["ptype"])) == set(range(6))
    b = O.init_particles(20000, 6, 0x5EED0002, 1, canvas)
    assert b["pos"].min() >= 0 and b["pos"].max() < 8000 and b["pos"].max() > 7900
    # keyed by id: any sub-range regenerates identically (multi-GPU ranks rely on this)
    c = O.init_particles(100, 6, 0x5EED0002, 1, canvas, id0=5000)
    assert c.tobytes() == b[5000:5100].tobytes()
    small = O.init_particles(1000, 6, 1, 0, np.float32([1000, 8000, 500]))
    assert small["pos"][:, 0].max() < 1000 and small["pos"][:, 2].max() < 500


def test_cell_keys():
    canvas = np.float32([8000, 4000, 2000])
    dims = np.int32([28, 14, 7])
    s, _ = U.random_state(5000, 6, 2, canvas)
    s["pos"][0] = (0, 0, 0)
    s["pos"][1] = np.nextafter(canvas, np.float32(0))
    k = O.cell_keys(s, canvas, dims)
    assert k[0] == 0 and k[1] == 28 * 14 * 7 - 1
    inv = dims.astype(np.float32) / canvas
    c = np.minimum((s["pos"] * inv).astype(np.int32), dims - 1)
    assert np.array_equal(k, ((c[:, 0] * 14 + c[:, 1]) * 7 + c[:, 2]).astype(np.uint32))
