"""Which proximity-graph kernel wins at which density: the reference's spawn cube (2000^3, 6 types, default
parameters, dist 200, maxConn 5) at several particle counts; wall time of buildGraphAsync + graphEdgeCount
(one host synchronisation) per build, median of 7 after 3 warm-ups.  usage: python tools/graph_crossover.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cellflow_b200 as cf

for n in (25_000, 50_000, 100_000, 200_000, 400_000, 800_000):
    row = []
    for gk in (1, 2):
        sim = cf.ParticleSimulation(n, 6, device=0, init=False)
        sim.initializeParticles(seed=0x5EED0002, mode=cf.INIT_SPAWN_CUBE)
        sim.setOption("graph_kernel", gk)
        ts = []
        for it in range(10):
            sim.moveUniverse(0.0, 0.0, 0.0)  # invalidates the sorted state: every build sorts again, as in a step loop
            t0 = time.perf_counter()
            sim.buildGraphAsync(200.0, 5)
            ne = sim.graphEdgeCount()
            ts.append(time.perf_counter() - t0)
        row.append((np.median(ts[3:]) * 1e3, ne, sim.stats().graph_kernel))
        sim.close()
    print(f"n={n:7d}  thread-per-particle {row[0][0]:7.3f} ms ({row[0][1]} edges)  warp-per-particle {row[1][0]:7.3f} ms ({row[1][1]} edges)", flush=True)
