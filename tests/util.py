"""Shared helpers of the test-suite: configurations restated from BASELINE.json / SURVEY.md 8(d),
conversion between the oracle's and the library's parameter structs, and the parity norms."""
import ctypes as C
import json
import os

import numpy as np

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRESETS = os.path.join(ROOT, "presets")
GOLDEN = os.path.join(ROOT, "tests", "golden")

FORCE_RTOL = 1e-5  # north_star: per-step forces within 1e-5 relative (of the summed |pair force|)


def preset_json(name):
    with open(os.path.join(PRESETS, name + ".json")) as f:
        return json.load(f)


def oracle_params_from_json(d, **over) -> O.Params:
    """loadPreset's key -> field mapping (CellFlowWidget.cpp:1087-1106), restated in Python."""
    p = O.Params()
    for k in ("radius", "delta_t", "friction", "repulsion", "attraction", "k", "balance",
              "forceMultiplier", "forceRange", "forceBias", "ratio", "lfoA", "lfoS", "forceOffset",
              "canvasWidth", "canvasHeight", "canvasDepth", "spawnRegionSize"):
        if k in d:
            setattr(p, k, d[k])
    p.numParticleTypes = d.get("numParticleTypes", 6)
    p.ratioWithLFO = p.ratio
    for k, v in over.items():
        setattr(p, k, v)
    return p


def config(name, **over):
    """(params, table, radio) of a shipped preset."""
    d = preset_json(name)
    p = oracle_params_from_json(d, **over)
    T = p.numParticleTypes
    radio = np.zeros(T, np.float32)
    r = np.array(d.get("radioByType", []), np.float32)[:T]
    radio[: len(r)] = r
    table = O.force_table(np.array(d["rawForceTable"], np.float32)[: T * T], T, p.forceRange,
                          p.forceBias, p.forceOffset)
    return p, table, radio


def to_lib_params(p):
    """oracle.Params -> cellflow_b200.Params (identical C layout)."""
    import cellflow_b200 as cf
    q = cf.Params()
    assert C.sizeof(q) == C.sizeof(p)
    C.memmove(C.byref(q), C.byref(p), C.sizeof(q))
    return q


def random_state(n, T, seed, canvas, mode="uniform", vel_scale=5.0, cube=2000.0):
    rng = np.random.default_rng(seed)
    p = O.particles(n)
    canvas = np.asarray(canvas, np.float32)
    if mode == "uniform":
        pos = rng.random((n, 3), dtype=np.float32) * canvas
    elif mode == "cube":
        span = np.minimum(cube, canvas)
        pos = (canvas - span) * 0.5 + rng.random((n, 3), dtype=np.float32) * span
    elif mode == "blobs":
        centers = rng.random((8, 3), dtype=np.float32) * canvas
        pos = centers[rng.integers(0, 8, n)] + rng.normal(0, cube / 8, (n, 3)).astype(np.float32)
        pos = np.mod(pos, canvas)
    else:
        raise ValueError(mode)
    pos = np.minimum(pos.astype(np.float32), np.nextafter(canvas, np.float32(0)))
    p["pos"] = pos
    p["vel"] = rng.normal(0, vel_scale, (n, 3)).astype(np.float32)
    p["ptype"] = rng.integers(0, T, n).astype(np.uint32)
    counts = rng.integers(0, 300, n).astype(np.int32)
    return p, counts


def force_rel_err(acc, acc_ref, fabs, mult):
    """max_i |acc_i - ref_i|_inf / (mult_i * sum_j |fv_ij| (|rep e_ij| + |att r_ij|)) — the
    north_star's 1e-5 norm, taken against the summed magnitude of the pair terms.  A net force of
    near-cancelling terms has no meaningful relative error of its own (SURVEY.md section 7 'parity
    norm'): the law's repulsion and attraction terms cancel at a radius inside the cutoff, and a
    particle whose only neighbour sits there has |net| ~ 1e-3 of either term."""
    scale = np.abs(mult) * fabs.astype(np.float64) + 1e-30
    err = np.abs(acc.astype(np.float64) - acc_ref.astype(np.float64)).max(axis=1)
    return err / scale


def force_multiplier_of(p, counts, prev):
    """m_i = forceMultiplier * (1 - (1 - balance) * min(avg/maxExp, 1)), .cu:136-143."""
    avg = (counts.astype(np.float64) + prev.astype(np.float64)) * 0.5
    dens = np.minimum(avg / p.maxExpectedNeighbors, 1.0)
    return p.forceMultiplier * (1.0 - (1.0 - p.balance) * dens)


def wrapped_abs_diff(a, b, canvas):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    return np.minimum(d, np.asarray(canvas, np.float64) - d)


def edge_set(edges):
    return set(zip(edges["i"].tolist(), edges["j"].tolist()))


def hilbert64(sx, sy, sz):
    """Hilbert index of the 4x4x4 sub-cell (sx, sy, sz) — Skilling's axes-to-transpose transform, 2 bits per
    axis, restated independently of the byte table in csrc/cf_device.cuh (cf_hilbert64)."""
    X = [np.asarray(sx, np.int64).copy(), np.asarray(sy, np.int64).copy(), np.asarray(sz, np.int64).copy()]
    bits, n = 2, 3
    M = 1 << (bits - 1)
    Q = M
    while Q > 1:
        P = Q - 1
        for i in range(n):
            m = (X[i] & Q) != 0
            X[0] = np.where(m, X[0] ^ P, X[0])
            t = np.where(m, 0, (X[0] ^ X[i]) & P)
            X[0] ^= t
            X[i] ^= t
        Q >>= 1
    for i in range(1, n):
        X[i] ^= X[i - 1]
    t = np.zeros_like(X[0])
    Q = M
    while Q > 1:
        t = np.where((X[n - 1] & Q) != 0, t ^ (Q - 1), t)
        Q >>= 1
    for i in range(n):
        X[i] ^= t
    h = np.zeros_like(X[0])
    for b in range(bits - 1, -1, -1):
        for i in range(n):
            h = (h << 1) | ((X[i] >> b) & 1)
    return h.astype(np.uint32)


def force_err_report(acc, acc_ref, fabs, fnet, mult):
    """The force error of one step under three norms (max over particles):
       gross   |err|_inf / (m * sum_j |fv| (|rep e| + |att r|))   — the tolerance scale (FORCE_RTOL)
       net     |err|_inf / (m * sum_j |f_ij|)                       — SURVEY.md section 7's norm
       finf    max_i |err_i|_inf / max_i |F_i|_inf                  — against the largest force of the system"""
    err = np.abs(acc.astype(np.float64) - acc_ref.astype(np.float64)).max(axis=1)
    m = np.abs(mult)
    gross = (err / (m * fabs.astype(np.float64) + 1e-30)).max()
    has = fnet > 0
    net = (err[has] / (m[has] * fnet[has].astype(np.float64))).max() if has.any() else 0.0
    finf = err.max() / max(np.abs(acc_ref.astype(np.float64)).max(), 1e-30)
    return {"gross": float(gross), "net": float(net), "finf": float(finf)}
