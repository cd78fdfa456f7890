"""Sanity run of the 8-GPU problem of config 5 (16 M particles, 64000 x 8000 x 8000) on ONE GPU:
sizes, sort key width, tile list and graph at 16 M."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import cellflow_b200 as cf
import bench
params, raw, radio, n, seed, mode, graph = bench.workload_setup("c5-settings-2M", 8)   # the 8-GPU problem: 16 M particles, 64000 x 8000 x 8000
print("particles", n, "canvas", params.canvasWidth, params.canvasHeight, params.canvasDepth, flush=True)
T = params.numParticleTypes
sim = cf.ParticleSimulation(n, T, device=0, init=False)
sim.params = params
sim.setRadioByType(radio); sim.setRawForceTableValues(raw)
sim.updateForceTable(params.forceRange, params.forceBias, params.forceOffset)
sim.initializeParticles(seed=seed, mode=mode)
sim.setOption("timing", 2)
for _ in range(3):
    sim.simulate(sync=False)
edges, _ = sim.generateProximityGraph(graph[0], graph[1])
print("graph edges", len(edges), "graph kernel", sim.stats().graph_kernel, flush=True)
sim.sync(); sim.statsReset()
t0 = time.perf_counter()
for _ in range(5):
    sim.simulate(sync=False)
sim.sync()
st = sim.stats()
print("16M on one GPU: ms/step", st.ms_total / st.steps, "wall", (time.perf_counter() - t0) / 5 * 1e3, "kernel", st.force_kernel, "mean nbrs", st.accepted_pairs / n, "grid", list(st.grid), flush=True)
cnt = sim.getNeighborCounts()
print("counts ok", int(cnt.min()), float(cnt.mean()), int(cnt.max()))
sim.close()
