"""compute-sanitizer target: small dense cases through every force kernel (1 per-particle, 2 tile gen. 3,
3 tile gen. 4 in both modes) and both graph kernels, four step+graph iterations each (so the CUDA-graph
replays run too).  `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_small.py`."""
import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import cellflow_b200 as cf
import util as U
# small dense cases through every force kernel and both graph kernels, a few steps each
for name, over, radio in (("eater", dict(ratioWithLFO=0.5, canvasWidth=2400.0, canvasHeight=2400.0, canvasDepth=2400.0), [1.0, 0.5, 0.0, 0.0, -0.5, 1.0]),
                          ("pulser", dict(canvasWidth=2000.0, canvasHeight=2000.0, canvasDepth=2000.0), None)):
    p, table, r0 = U.config(name, **over)
    radio = np.float32(radio) if radio is not None else r0
    state, counts = U.random_state(6000, p.numParticleTypes, 5, p.canvas, "uniform")
    for fk in (1, 2, 3):
        for gk in (1, 2):
            sim = cf.ParticleSimulation(len(state), p.numParticleTypes, init=False)
            sim.params = U.to_lib_params(p)
            sim.setRadioByType(radio); sim.setForceTable(table)
            sim.setOption("force_kernel", fk); sim.setOption("graph_kernel", gk)
            sim.setParticleData(state, counts)
            for _ in range(4):
                sim.simulate()
                e, _ = sim.generateProximityGraph(200.0, 5)
            print(name, fk, gk, sim.stats().force_kernel, len(e), flush=True)
            sim.close()
print("done")
