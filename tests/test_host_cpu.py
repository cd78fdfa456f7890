"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header
declares, preset load/save (CellFlowWidget::loadPreset / savePreset), host-only table code.
No compute entry point is called here — those need a GPU (tests marked `gpu`)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import cellflow_b200 as cf
from cellflow_b200 import _lib
import oracle as O
import util as U
from conftest import have_gpu

HEADER = os.path.join(U.ROOT, "include", "cellflow_b200.h")


def test_library_exports_every_declared_symbol():
    L = cf.lib()
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(cf_[a-z0-9_]+)\s*\(", text))
    assert declared, "no declarations found"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} not exported by {cf.lib_path()}"
    assert b"sm_100a" in L.cf_version()


def test_struct_layouts_match_header():
    assert C.sizeof(cf.Params) == 21 * 4
    assert C.sizeof(O.Params) == C.sizeof(cf.Params)
    assert cf.PARTICLE.itemsize == 44  # reference Particle, SimulationParams.h:6-12
    assert [cf.PARTICLE.fields[k][1] for k in ("pos", "vel", "acc", "ptype", "pad")] == [0, 12, 24, 36, 40]
    p = cf.default_params()
    d = O.Params()
    for f, _ in cf.Params._fields_:
        assert getattr(p, f) == getattr(d, f), f


def test_ctypes_mirror_matches_the_compiled_header(tmp_path):
    """The header is the contract: gcc compiles it as plain C and prints size and field offsets of
    the structs that cross the boundary by pointer; the ctypes mirror must agree byte for byte."""
    import subprocess

    fields = {"cf_stats": [f for f, _ in _lib.Stats._fields_], "cf_params": [f for f, _ in cf.Params._fields_]}
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void) {']
    for st, fs in fields.items():
        prog.append(f'  printf("{st} %zu", sizeof({st}));')
        for f in fs:
            prog.append(f'  printf(" %zu", offsetof({st}, {f}));')
        prog.append('  printf("\\n");')
    prog += ['  printf("cf_particle %zu\\n", sizeof(cf_particle));', '  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-o", str(exe), str(src)])
    out = dict((l.split()[0], [int(x) for x in l.split()[1:]]) for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for st, cls in (("cf_stats", _lib.Stats), ("cf_params", cf.Params)):
        assert out[st][0] == C.sizeof(cls), st
        assert out[st][1:] == [getattr(cls, f).offset for f, _ in cls._fields_], st
    assert out["cf_particle"] == [44]


@pytest.mark.parametrize("name", ["settings", "littlecells", "eater", "pulser"])
def test_load_preset_matches_json(name):
    pr = cf.load_preset(os.path.join(U.PRESETS, name + ".json"))
    d = U.preset_json(name)
    assert pr.particleCount == d["PARTICLE_COUNT"]
    T = d["numParticleTypes"]
    assert pr.params.numParticleTypes == T
    for k in ("radius", "delta_t", "friction", "repulsion", "attraction", "k", "balance",
              "forceMultiplier", "forceRange", "forceBias", "ratio", "lfoA", "lfoS", "forceOffset"):
        assert getattr(pr.params, k) == np.float32(d[k]), k
    # settings.json has no canvas keys -> SimulationParams defaults (8000^3)
    for k in ("canvasWidth", "canvasHeight", "canvasDepth"):
        assert getattr(pr.params, k) == np.float32(d.get(k, 8000.0))
    assert pr.params.ratioWithLFO == pr.params.ratio
    assert pr.params.maxExpectedNeighbors == 400  # never serialised, SimulationParams.h:29
    assert pr.numRadio == T and np.array_equal(pr.radio, np.float32(d["radioByType"][:T]))
    assert pr.numRawForce == T * T
    assert np.array_equal(pr.raw_force, np.float32(d["rawForceTable"][: T * T]))
    assert pr.numColors == min(len(d["particleColors"]), 10)
    assert pr.colors[0]["r"] == np.float32(d["particleColors"][0]["r"])
    assert pr.pointSize == np.float32(d["pointSize"])
    if "enableDepthFade" in d:
        assert bool(pr.enableDepthFade) == d["enableDepthFade"]
        assert pr.effectType == d["effectType"] and bool(pr.invertPan) == d["invertPan"]


def test_preset_round_trip(tmp_path):
    for name in ("settings", "eater"):
        a = cf.load_preset(os.path.join(U.PRESETS, name + ".json"))
        out = tmp_path / (name + "_saved.json")
        cf.save_preset(str(out), a)
        json.load(open(out))  # valid JSON
        b = cf.load_preset(str(out))
        assert bytes(a) == bytes(b)


def test_load_preset_errors(tmp_path):
    with pytest.raises(cf.CellFlowError):
        cf.load_preset(str(tmp_path / "missing.json"))  # loadPreset returns false
    bad = tmp_path / "bad.json"
    bad.write_text("{ \"radius\": ")
    with pytest.raises(cf.CellFlowError):
        cf.load_preset(str(bad))
    odd = tmp_path / "odd.json"  # unknown keys ignored, wrong kinds read as 0 (QJsonValue rules)
    odd.write_text('{"metaball": 3, "radius": "x", "numParticleTypes": 4, "radioByType": [1, 2, 3, 4, 5]}')
    pr = cf.load_preset(str(odd))
    assert pr.params.radius == 0.0 and pr.numRadio == 4 and pr.params.numParticleTypes == 4


def test_reference_default_tables_match_libc():
    """The library replays glibc's unseeded rand() privately; the oracle calls the real one."""
    for T in (6, 8, 10):
        raw, radio, eff = cf.reference_default_tables(T)
        oraw, oradio = O.default_tables(T)
        assert np.array_equal(raw, oraw) and np.array_equal(radio, oradio)
        assert np.array_equal(eff, O.force_table(oraw, T, 0.28, -0.20, 1.0))


def test_lfo_matches_oracle():
    L = cf.lib()
    for (ratio, a, s, t) in ((0.3, 0.0, 1.0, 0.7), (0.3, 0.5, 0.1, 2.5), (-0.42, 1.0, 5.15, 123.456)):
        p = cf.default_params(ratio=ratio, lfoA=a, lfoS=s)
        q = O.Params(ratio=ratio, lfoA=a, lfoS=s)
        assert L.cf_ratio_with_lfo(C.byref(p), C.c_float(t)) == O.ratio_with_lfo(q, t)


@pytest.mark.skipif(have_gpu(), reason="GPU present: creation succeeds")
def test_no_cpu_fallback():
    """Without a device the product fails loudly instead of computing anything on the CPU."""
    with pytest.raises(cf.CellFlowError) as e:
        cf.ParticleSimulation(100)
    assert e.value.code == -2 and "no CPU path" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(U.ROOT, "cellflow_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(root, f), errors="replace").read()
                assert "liboracle" not in text and "import oracle" not in text, f
                assert "cellflow_oracle" not in text or f.endswith((".cuh",)), f


def test_hilbert_table_is_a_space_filling_curve():
    """The 64-byte table of csrc/cf_device.cuh (cf_hilbert64) against the independent restatement in
    tests/util.py: a bijection onto 0..63 whose consecutive sub-cells are face neighbours."""
    import re
    import util as U
    src = open(os.path.join(U.ROOT, "cellflow_b200", "csrc", "cf_device.cuh")).read()
    body = src[src.index("cf_hilbert64("):src.index("cf_sort_key(")]
    words = [int(w, 16) for w in re.findall(r"v = (0x[0-9a-f]{8})u", body)]
    assert len(words) == 16
    table = np.array([(words[i >> 2] >> (8 * (i & 3))) & 255 for i in range(64)], np.uint32)
    g = np.arange(4)
    sx, sy, sz = [a.ravel() for a in np.meshgrid(g, g, g, indexing="ij")]
    want = U.hilbert64(sx, sy, sz)
    assert np.array_equal(table[(sx << 4) | (sy << 2) | sz], want)
    assert sorted(want.tolist()) == list(range(64))
    order = np.argsort(want)
    pts = np.stack([sx, sy, sz], 1)[order]
    assert np.abs(np.diff(pts, axis=0)).sum(1).max() == 1
