"""The C++ host side on a GPU: the headless step loop (preset -> tables -> LFO/simulate/graph
loop through the C ABI) and the ParticleSimulation-shaped shim compiled against the reference's
class interface."""
import json
import os
import subprocess

import numpy as np
import pytest

import cellflow_b200 as cf
import util as U

pytestmark = pytest.mark.gpu
BIN = os.path.join(U.ROOT, "cellflow_b200", "bin")


def test_headless_step_loop(tmp_path):
    exe = os.path.join(BIN, "cellflow_headless")
    if not os.path.exists(exe):
        pytest.skip("cellflow_headless not built")
    saved = tmp_path / "saved.json"
    r = subprocess.run([exe, "--preset", os.path.join(U.PRESETS, "pulser.json"), "--steps", "20", "--graph", "200", "5",
                        "--save", str(saved)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["particles"] == 30000 and out["steps"] == 20
    assert out["mean_neighbours"] > 50 and out["edges_per_step"] > 0 and out["launches"] > 0
    # the same 20 steps through the Python binding give the same neighbour statistics
    pr = cf.load_preset(os.path.join(U.PRESETS, "pulser.json"))
    sim = cf.ParticleSimulation(100, 6)
    sim.applyPreset(pr)
    sim.initializeParticles(seed=0x5EED0000, mode=cf.INIT_SPAWN_CUBE)
    sim.simulate(steps=20)
    assert abs(sim.getNeighborCounts().mean() - out["mean_neighbours"]) < 1e-2
    sim.close()
    again = cf.load_preset(str(saved))       # savePreset round trip keeps the physics
    assert bytes(again.params) == bytes(pr.params) and np.array_equal(again.raw_force, pr.raw_force)


def test_reference_class_shim():
    exe = os.path.join(BIN, "shim_check")
    if not os.path.exists(exe):
        pytest.skip("shim_check not built")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "SHIM_OK particles=4000" in r.stdout and "radio2=0.50" in r.stdout
