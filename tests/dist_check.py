"""Multi-rank check, run under torch.distributed.run on a box with >= 2 GPUs
(tests/test_slab_gpu.py launches it): the slab-decomposed engine against the CPU oracle and
against itself across world sizes — neighbour counts and graph edge sets must be identical to
the single-domain result, forces within 1e-5, over several steps with real migration."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import cellflow_b200 as cf  # noqa: E402
from cellflow_b200 import dist as cfd  # noqa: E402
import oracle as O  # noqa: E402
import util as U  # noqa: E402


def run_case(name, p, table, radio, state, counts, steps, graph, rank, world, balanced=False, capacity_factor=1.5,
             options=None):
    import torch.distributed as dist
    n = len(state)
    lp = U.to_lib_params(p)
    bounds = None
    if balanced:  # clustered state: equal-count slabs from the x histogram, every slab >= one interaction radius
        hist, _ = np.histogram(state["pos"][:, 0], bins=4096, range=(0.0, float(p.canvasWidth)))
        bounds = cfd.balanced_bounds(hist, p.canvasWidth, world, cfd.interaction_radius(lp, radio) * 1.002)
    sim, rank, world = cfd.make_slab_sim(lp, None, radio, n, None, cf.INIT_UNIFORM, force_table=table, bounds=bounds,
                                         capacity_factor=capacity_factor)
    for k, v in (options or {}).items():
        sim.setOption(k, v)
    mine, mcounts, ids = cfd.partition(state, counts, p.canvasWidth, rank, world, bounds)
    sim.uploadOwned(mine, mcounts, ids)
    want_state, want_counts = state, counts
    first_ids = set(ids.tolist())
    for step in range(steps):
        sim.simulate()
        pp, cc, ii = sim.downloadOwned()
        got, gcnt = cfd.gather_particles(pp, cc, ii, n)
        all_edges = None
        if graph:
            edges, _ = sim.generateProximityGraph(*graph)
            all_edges = [None] * world if rank == 0 else None
            dist.gather_object(edges.tobytes(), all_edges, dst=0)
        else:
            sim.cellKeys()  # consumes the pending migrant exchange like the graph build does
        if rank == 0:
            want, wcnt, fabs = O.step(want_state, want_counts, p, table, radio, "cells", 8)
            assert np.array_equal(gcnt, wcnt), f"{name} step {step}: counts differ on {(gcnt != wcnt).sum()} particles"
            mult = U.force_multiplier_of(p, wcnt, want_counts)
            rel = U.force_rel_err(got["acc"], want["acc"], fabs, mult).max()
            assert rel <= U.FORCE_RTOL, (name, rel)
            if graph:
                es = set()
                for b in all_edges:
                    es |= U.edge_set(np.frombuffer(b, cf.EDGE))
                assert es == U.edge_set(O.graph(got, graph[0], graph[1], canvas=p.canvas, method="cells")), f"{name} step {step}: edges"
            want_state, want_counts = got, gcnt
        if balanced and step == steps // 2:
            # re-balance in the middle of the run: new bounds from the current x histogram, particles redistributed
            _, owned_now = cfd.rebalance(sim, radio)
        # every rank continues from its own device state (no re-upload): real migration
    # migration is applied by the cell-list build that follows a step (the graph build / cellKeys above)
    pp, _, ii = sim.downloadOwned()
    total_migrated = cfd.all_reduce_sum(float(len(set(ii.tolist()) - first_ids)))
    lo, hi = sim.slabBounds()
    assert np.all((pp["pos"][:, 0] >= lo) & (pp["pos"][:, 0] < hi)), f"{name}: a rank holds a particle outside its slab"
    st = sim.stats()
    owned = [int.from_bytes(b, "little") for b in cfd.all_gather_bytes(int(st.n_owned).to_bytes(8, "little"))]
    balance = max(owned) * world / max(sum(owned), 1)
    if balanced:
        assert balance <= 1.3, f"{name}: owned particles per rank {owned} (max/mean {balance:.2f})"
    # the cell grid must be the same on every rank (it once depended on the rank's own count)
    grids = cfd.all_gather_bytes(bytes(np.int32(list(st.grid)[1:]).tobytes()))
    assert len(set(grids)) == 1, f"{name}: ranks disagree on the (y, z) grid: {[np.frombuffer(g, np.int32).tolist() for g in grids]}"
    if rank == 0:
        print(f"DIST_CASE_OK {name} world={world} n={n} steps={steps} migrated={int(total_migrated)} "
              f"ghosts_rank0={st.n_ghost} grid={list(st.grid)} force_kernel={st.force_kernel} "
              f"owned_max_over_mean={balance:.3f}", flush=True)
    sim.close()
    cfd.barrier()
    return int(total_migrated)


def main():
    rank, world = cfd.init_process_group("nccl")
    # A. dense, uniform radius: large dt so that every step migrates particles; graph on
    p, table, radio = U.config("pulser", delta_t=0.9)
    state, counts = U.random_state(120_000, 6, 23, p.canvas, "uniform", vel_scale=60.0)
    mig = run_case("pulser-120k", p, table, radio, state, counts, 4, (200.0, 5), rank, world)
    # B. per-type radii (type-sorted j copy over owned + ghost slots), clustered cube, graph on
    p, table, _ = U.config("eater", ratioWithLFO=0.5, delta_t=0.5)
    radio = np.float32([1.0, 0.5, 0.0, 0.0, -0.5, 1.0])
    state, counts = U.random_state(90_000, 6, 29, p.canvas, "uniform", vel_scale=80.0)
    run_case("eater-radii-90k", p, table, radio, state, counts, 3, (200.0, 5), rank, world)
    # C. small radius, low density: the grid is bound by cells-per-particle, not by the radius — the regime
    #    in which a grid derived from the rank's own count differs between ranks (default preset, 100 k)
    p = O.Params()
    raw, radio = O.default_tables(6)
    table = O.force_table(raw, 6, p.forceRange, p.forceBias, p.forceOffset)
    state, counts = U.random_state(100_000, 6, 31, p.canvas, "uniform", vel_scale=300.0)
    #    (no graph here: the rule's 200 exceeds the one-cell ghost layer of this fine grid, which slab mode refuses)
    run_case("default-100k-sparse", p, table, radio, state, counts, 3, None, rank, world)
    # D. the reference's own spawn rule (centred 2000-wide cube, ParticleSimulation.cu:45-55) and a blob state: the
    #    uniform split would put everything on one or two ranks; equal-count slabs + a re-balance mid-run
    p = O.Params()
    raw, radio = O.default_tables(6)
    table = O.force_table(raw, 6, p.forceRange, p.forceBias, p.forceOffset)
    state = O.init_particles(100_000, 6, 0x5EED0002, cf.INIT_SPAWN_CUBE, p.canvas)
    run_case("spawn-cube-100k", p, table, radio, state, np.zeros(len(state), np.int32), 4, None, rank, world, balanced=True)
    p, table, radio = U.config("pulser", delta_t=0.3)   # (dense blobs accelerate hard: a step must stay below one slab width)
    state, counts = U.random_state(80_000, 6, 37, p.canvas, "blobs", vel_scale=40.0)
    if world * cfd.interaction_radius(U.to_lib_params(p), radio) * 1.002 <= p.canvasWidth:
        run_case("blobs-80k", p, table, radio, state, counts, 4, (200.0, 5), rank, world, balanced=True)
    # E. ranks that own nothing: every particle starts in the first half of rank 0's slab (uniform bounds, capacity for
    #    all of them on one rank); the empty ranks still take part in every exchange
    p, table, radio = U.config("pulser", delta_t=0.5)
    state, counts = U.random_state(40_000, 6, 41, p.canvas, "uniform", vel_scale=30.0)
    state["pos"][:, 0] *= np.float32(0.5 / world)
    run_case("one-rank-owns-all-40k", p, table, radio, state, counts, 3, (200.0, 5), rank, world,
             capacity_factor=1.2 * world)
    # F. a graph distance larger than the interaction radius: option "slab_min_layer_width" widens the x layers (and
    #    the ghost layers with them) so that the rule's 200 fits — the sparse case C with the graph on
    p = O.Params()
    raw, radio = O.default_tables(6)
    table = O.force_table(raw, 6, p.forceRange, p.forceBias, p.forceOffset)
    state, counts = U.random_state(100_000, 6, 43, p.canvas, "uniform", vel_scale=300.0)
    run_case("default-100k-graph-wider-layers", p, table, radio, state, counts, 2, (200.0, 5), rank, world,
             options={"slab_min_layer_width": 200.5})
    if rank == 0:
        assert mig > 0, "test did not exercise migration"
        print(f"DIST_CHECK_OK world={world} migrated={mig}", flush=True)


if __name__ == "__main__":
    main()
