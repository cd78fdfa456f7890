#!/usr/bin/env python3
"""Generates tests/golden/*.npz by running the REFERENCE's own kernels and host methods
(oracle/_ref/libcellflow_ref.so = the unmodified cuda-native/src/ParticleSimulation.cu compiled
for sm_100a, driven by oracle/ref_harness.cu) on seeded inputs.  Needs a GPU:

    gpurun -- python tools/make_golden.py gpurun_out/golden     # then copy into tests/golden/

The step vectors come from ref_simulate_exact (one warp at a time, so the reference's in-place
update race cannot occur).  Inputs are stored next to the outputs so the fixtures stand alone.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O  # noqa: E402
import util as U  # noqa: E402

out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(out_dir, exist_ok=True)
R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libcellflow_ref.so"))
assert R.ref_device_count() > 0, "needs a GPU"


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def params_array(p):
    return np.frombuffer(bytes(p), dtype=np.uint8).copy()


def ref_step(state, counts, p, table, radio):
    out = np.zeros_like(state)
    cnt = np.zeros(len(state), np.int32)
    rc = R.ref_simulate_exact(vp(state), vp(counts), C.c_int(len(state)), C.byref(p), vp(table), vp(radio),
                              vp(out), vp(cnt))
    assert rc == 0
    return out, cnt


def save_step(tag, state, counts, p, table, radio):
    out, cnt = ref_step(state, counts, p, table, radio)
    np.savez_compressed(os.path.join(out_dir, f"step_{tag}.npz"), state=state, counts=counts,
                        params=params_array(p), table=table, radio=radio, out=out, cnt=cnt)
    print(f"step_{tag}: n={len(state)} mean count {cnt.mean():.1f}")


# 1. settings.json, reference spawn shape, dense
p, table, radio = U.config("settings")
state, counts = U.random_state(2048, 8, 101, p.canvas, "cube", cube=1200.0)
save_step("settings", state, counts, p, table, radio)

# 2. eater.json with per-type radii (ratio matters), small periodic canvas: wrap on every axis
p, table, radio = U.config("eater", ratioWithLFO=0.7, canvasWidth=2400.0, canvasHeight=2000.0, canvasDepth=1800.0)
radio = np.float32([1.0, 0.5, 0.0, 0.0, -0.5, 1.0])
state, counts = U.random_state(1536, 6, 102, p.canvas, "uniform")
save_step("eater_radii_wrap", state, counts, p, table, radio)

# 3. struct defaults + default (glibc rand) tables, balance > 1 branch of the adaptive factor
p = O.Params(balance=1.3, maxExpectedNeighbors=40, ratioWithLFO=0.4)
raw, radio = O.default_tables(6)
table = O.force_table(raw, 6, 0.28, -0.20, 1.0)
state, counts = U.random_state(2048, 6, 103, p.canvas, "cube", cube=330.0)
save_step("defaults", state, counts, p, table, radio)

# 4. pulser.json, coincident particles and particles on the seam
p, table, radio = U.config("pulser")
state, counts = U.random_state(1024, 6, 104, p.canvas, "cube", cube=900.0)
state["pos"][:16] = state["pos"][16]
state["pos"][32:48] = 0.0
state["pos"][48:64] = np.nextafter(p.canvas, np.float32(0))
save_step("pulser_edge", state, counts, p, table, radio)

# 5. proximity graph (reference VBO content, order canonicalised by sorting the edge records)
p, _, _ = U.config("eater")
for tag, mode, n, dist, mc in (("cube", "cube", 4096, 200.0, 5), ("blobs", "blobs", 3000, 120.0, 16)):
    state, _ = U.random_state(n, 6, 105, p.canvas, mode)
    colors = np.zeros(10, O.COLOR)
    colors["r"], colors["g"], colors["b"] = np.arange(10) * 0.1, 0.5, 1.0 - np.arange(10) * 0.1
    verts = np.zeros((n * mc * 2, 6), np.float32)
    nv = C.c_int(0)
    ms = C.c_float(0)
    rc = R.ref_graph(vp(state), C.c_int(n), C.c_int(6), C.c_float(dist), C.c_int(mc), vp(colors), C.c_int(10),
                     vp(verts), C.c_int(len(verts)), C.byref(nv), C.byref(ms))
    assert rc == 0
    rec = verts[: nv.value].reshape(-1, 12)
    rec = rec[np.lexsort(rec.T[::-1])]
    np.savez_compressed(os.path.join(out_dir, f"graph_{tag}.npz"), state=state, dist=np.float32(dist),
                        max_conn=np.int32(mc), colors=colors, records=rec)
    print(f"graph_{tag}: {len(rec)} edges")

# 6. host tables through the reference class itself
tabs = {}
for T in (6, 8):
    raw = np.zeros(T * T, np.float32)
    radio = np.zeros(T, np.float32)
    eff = np.zeros(T * T, np.float32)
    assert R.ref_tables(C.c_int(T), None, C.c_float(0), C.c_float(0), C.c_float(0), vp(raw), vp(radio), vp(eff)) == 0
    tabs[f"raw{T}"], tabs[f"radio{T}"], tabs[f"eff{T}"] = raw, radio, eff
d = U.preset_json("eater")
raw_in = np.float32(d["rawForceTable"])
raw = np.zeros(36, np.float32)
radio = np.zeros(6, np.float32)
eff = np.zeros(36, np.float32)
assert R.ref_tables(C.c_int(6), vp(raw_in), C.c_float(d["forceRange"]), C.c_float(d["forceBias"]),
                    C.c_float(d["forceOffset"]), vp(raw), vp(radio), vp(eff)) == 0
tabs["eater_eff"] = eff
d = U.preset_json("settings")
raw_in = np.float32(d["rawForceTable"])
raw = np.zeros(64, np.float32)
radio = np.zeros(8, np.float32)
eff = np.zeros(64, np.float32)
assert R.ref_tables(C.c_int(8), vp(raw_in), C.c_float(d["forceRange"]), C.c_float(d["forceBias"]),
                    C.c_float(d["forceOffset"]), vp(raw), vp(radio), vp(eff)) == 0
tabs["settings_eff"] = eff
np.savez_compressed(os.path.join(out_dir, "tables.npz"), **tabs)

# 7. moveParticlesKernel
state, _ = U.random_state(512, 6, 107, np.float32([8000, 8000, 8000]))
moved = state.copy()
assert R.ref_move(vp(moved), C.c_int(512), C.c_float(123.5), C.c_float(-77.25), C.c_float(4000.0),
                  C.c_float(8000), C.c_float(8000), C.c_float(8000)) == 0
np.savez_compressed(os.path.join(out_dir, "move.npz"), state=state, moved=moved)
print("golden vectors written to", out_dir)
