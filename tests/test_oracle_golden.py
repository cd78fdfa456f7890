"""Pins the oracle to the REFERENCE ITSELF: tests/golden/*.npz were produced on a B200 by the
reference's own kernels and host methods (unmodified cuda-native/src/ParticleSimulation.cu
compiled for sm_100a; generator: tools/make_golden.py through oracle/ref_harness.cu).  The step
vectors were computed one warp at a time, which removes the reference's in-place update race
without touching its code.  Runs on CPU: nothing here needs /root/reference or a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle as O
import util as U

G = U.GOLDEN


def load_params(raw):
    p = O.Params()
    assert len(raw) == C.sizeof(p)
    C.memmove(C.byref(p), raw.tobytes(), C.sizeof(p))
    return p


@pytest.mark.parametrize("tag", ["settings", "eater_radii_wrap", "defaults", "pulser_edge"])
def test_step_matches_reference_kernel(tag):
    z = np.load(os.path.join(G, f"step_{tag}.npz"))
    p = load_params(z["params"])
    state, counts = z["state"].view(O.PARTICLE).reshape(-1), z["counts"]
    ref, ref_cnt = z["out"].view(O.PARTICLE).reshape(-1), z["cnt"]
    for method in ("brute", "cells"):
        out, cnt, fabs = O.step(state, counts, p, z["table"], z["radio"], method, 4)
        # neighbour decision: bit-exact against the reference kernel, every particle
        assert np.array_equal(cnt, ref_cnt)
        # forces: only MUFU.EX2 (inside CUDA expf) and the summation order of the rotated launch
        # separate the two -> a few ulp of the summed term magnitude
        mult = U.force_multiplier_of(p, cnt, counts)
        rel = U.force_rel_err(out["acc"], ref["acc"], fabs, mult)
        assert rel.max() < 5e-7, rel.max()
        assert np.array_equal(out["ptype"], ref["ptype"])
        dv = np.abs(out["vel"].astype(np.float64) - ref["vel"])
        assert np.all(dv.max(axis=1) <= 5e-7 * p.delta_t * np.abs(mult) * fabs + 1e-7 * np.abs(ref["vel"]).max(axis=1) + 1e-12)
        dp = U.wrapped_abs_diff(out["pos"], ref["pos"], p.canvas)
        assert dp.max() <= 1.01 * float(np.spacing(np.float32(2 * p.canvas.max())))
    assert ref_cnt.mean() > 15
    # most particles agree to the last bit in position
    assert (out["pos"] == ref["pos"]).all(axis=1).mean() > 0.9


def test_epilogue_is_bit_exact_given_the_reference_force():
    """Feed the reference's own force back through the oracle's integrate law: vel/pos must match
    the reference kernel bit for bit (pins .cu:136-161 incl. fmodf and the fused rounding points)."""
    for tag in ("settings", "defaults", "pulser_edge", "eater_radii_wrap"):
        z = np.load(os.path.join(G, f"step_{tag}.npz"))
        p = load_params(z["params"])
        state, counts = z["state"].view(O.PARTICLE).reshape(-1), z["counts"]
        ref = z["out"].view(O.PARTICLE).reshape(-1)
        f32 = np.float32
        acc = ref["acc"]
        vel = (state["vel"].astype(np.float64) * f32(p.friction) + (acc * f32(p.delta_t)).astype(f32).astype(np.float64)).astype(f32)
        # fma(vel, friction, acc*dt): the double evaluation above is exact before the final rounding
        assert np.array_equal(vel, ref["vel"])
        pos = (vel.astype(np.float64) * f32(p.delta_t) + state["pos"].astype(np.float64)).astype(f32)
        pos = np.fmod((pos + p.canvas).astype(f32), p.canvas).astype(f32)
        assert np.array_equal(pos, ref["pos"])


@pytest.mark.parametrize("tag", ["cube", "blobs"])
def test_graph_matches_reference_kernel(tag):
    z = np.load(os.path.join(G, f"graph_{tag}.npz"))
    state = z["state"].view(O.PARTICLE).reshape(-1)
    dist, mc = float(z["dist"]), int(z["max_conn"])
    colors = z["colors"].view(O.COLOR).reshape(-1)
    for method in ("brute", "cells"):
        edges = O.graph(state, dist, mc, canvas=np.float32([8000, 8000, 8000]), method=method)
        rec = O.graph_vertices(state, edges, colors, 6)
        rec = rec[np.lexsort(rec.T[::-1])]
        assert rec.shape == z["records"].shape
        assert np.array_equal(rec, z["records"])      # the reference's VBO content as a set


def test_tables_match_reference_class():
    z = np.load(os.path.join(G, "tables.npz"))
    for T in (6, 8):
        raw, radio = O.default_tables(T)
        if T == 6:  # T=8 goes through setNumParticleTypes: a second round of rand() draws
            assert np.array_equal(raw, z["raw6"]) and np.array_equal(radio, z["radio6"])
            assert np.array_equal(O.force_table(raw, 6, 0.28, -0.20, 1.0), z["eff6"])
        else:
            assert np.array_equal(O.force_table(z["raw8"], 8, 0.28, -0.20, 1.0), z["eff8"])
    for name, T in (("eater", 6), ("settings", 8)):
        p, table, _ = U.config(name)
        assert np.array_equal(table, z[f"{name}_eff"])


def test_move_matches_reference_kernel():
    z = np.load(os.path.join(G, "move.npz"))
    state = z["state"].view(O.PARTICLE).reshape(-1)
    moved = z["moved"].view(O.PARTICLE).reshape(-1)
    out = O.move_universe(state, 123.5, -77.25, 4000.0, np.float32([8000, 8000, 8000]))
    assert np.array_equal(out["pos"], moved["pos"])
