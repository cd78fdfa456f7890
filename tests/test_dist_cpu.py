"""CPU tests of the one-process-per-GPU host logic (cellflow_b200/dist.py) on the gloo backend,
world_size 2: rendezvous from the environment, byte broadcast / all-gather (the mailbox-handle path), slab
ownership / partition / gather, max-over-ranks reductions.  The GPU step itself is replaced by
the oracle here, so what is checked is that slabs + migration of ownership + gather reproduce
the single-domain result."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import util as U
from cellflow_b200 import dist as cfd


def test_slab_bounds_and_owner():
    W = 8000.0
    for world in (1, 2, 3, 4, 8):
        b = [cfd.slab_bound(W, r, world) for r in range(world + 1)]
        assert b[0] == 0 and b[-1] == np.float32(W) and all(x < y for x, y in zip(b, b[1:]))
        x = np.float32([0.0, 1e-3, W / world, np.nextafter(np.float32(W / world), np.float32(0)),
                        np.nextafter(np.float32(W), np.float32(0))])
        o = cfd.slab_owner(x, W, world)
        assert o[0] == 0 and o[-1] == world - 1
        if world > 1:
            assert o[2] == 1 and o[3] == 0          # bound belongs to the right-hand slab
    rng = np.random.default_rng(0)
    x = (rng.random(100000) * W).astype(np.float32)
    o = cfd.slab_owner(x, W, 8)
    assert o.min() == 0 and o.max() == 7
    for r in range(8):
        sel = x[o == r]
        assert sel.min() >= cfd.slab_bound(W, r, 8) and sel.max() < cfd.slab_bound(W, r + 1, 8)


WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    sys.path[:0] = [{root!r}, {root!r} + "/oracle", {root!r} + "/tests"]
    import oracle as O, util as U
    from cellflow_b200 import dist as cfd
    import torch.distributed as dist

    rank, world = cfd.init_process_group("gloo")
    assert world == 2 and dist.get_backend() == "gloo"
    # 1. byte broadcast, and the all-gather that carries the mailbox IPC handles to the ring neighbours
    blob = bytes(range(128)) if rank == 0 else None
    assert cfd.broadcast_bytes(blob, 128, 0) == bytes(range(128))
    handles = cfd.all_gather_bytes(bytes([rank]) * 64)
    assert handles == [bytes([0]) * 64, bytes([1]) * 64]
    assert handles[(rank - 1) % world] == handles[(rank + 1) % world] == bytes([1 - rank]) * 64
    # 2. reductions used by bench.py (max over ranks of the step time, sums of counters)
    assert cfd.all_reduce_max(1.0 + rank) == 2.0 and cfd.all_reduce_sum(1.0 + rank) == 3.0
    # 3. slabs: partition -> step -> ownership moves with x -> gather == single-domain oracle
    p, table, radio = U.config("pulser", delta_t=0.9)
    n = 6000
    state, counts = U.random_state(n, 6, 5, p.canvas, "uniform", vel_scale=80.0)
    mine, mcounts, ids = cfd.partition(state, counts, p.canvasWidth, rank, world)
    assert len(ids) > 0 and np.all(cfd.slab_owner(mine["pos"][:, 0], p.canvasWidth, world) == rank)
    cur, cur_counts = state, counts
    for step in range(3):
        want, wcnt, _ = O.step(cur, cur_counts, p, table, radio, "cells", 2)
        # each rank advances only what it owns (stand-in for the GPU step), then ownership is
        # re-derived from the new x, exactly what the migration exchange implements
        mine_new, mine_cnt = want[ids], wcnt[ids]
        full, fcnt = cfd.gather_particles(mine_new, mine_cnt, ids, n)
        box = [full, fcnt]
        dist.broadcast_object_list(box, src=0)
        full, fcnt = box
        assert full.tobytes() == want.tobytes() and np.array_equal(fcnt, wcnt)
        _, _, new_ids = cfd.partition(full, fcnt, p.canvasWidth, rank, world)
        moved = len(set(new_ids.tolist()) ^ set(ids.tolist()))
        ids, cur, cur_counts = new_ids, full, fcnt
    total = cfd.all_reduce_sum(float(len(ids)))
    assert total == n
    cfd.barrier()
    if rank == 0:
        print("GLOO_OK moved_last_step", moved)
''')


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=U.ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "GLOO_OK" in r.stdout


def test_balanced_bounds_equalise_a_clustered_state():
    """Slab bounds from the x histogram (cellflow_b200/dist.py): the reference's spawn cube (2000 wide in an 8000 box)
    split over 8 ranks — equal counts, every slab at least one interaction radius wide, ends pinned to the box."""
    from cellflow_b200 import dist as cfd
    rng = np.random.default_rng(3)
    x = (rng.random(200_000) * 2000 + 3000).astype(np.float32)
    hist, _ = np.histogram(x, bins=4096, range=(0.0, 8000.0))
    b = cfd.balanced_bounds(hist, 8000.0, 8, 42.2)
    assert b[0] == 0 and b[-1] == 8000 and np.all(np.diff(b) >= 42.2)
    owned = np.bincount(cfd.slab_owner(x, 8000.0, 8, b), minlength=8)
    assert owned.max() * 8 / owned.sum() <= 1.02
    # a radius that does not leave room for equal counts: the widths win, nothing overlaps, all particles are owned
    b = cfd.balanced_bounds(hist, 8000.0, 8, 600.0)
    assert np.all(np.diff(b) >= 600.0) and b[-1] == 8000
    assert np.bincount(cfd.slab_owner(x, 8000.0, 8, b), minlength=8).sum() == len(x)
    # uniform state: the uniform split comes back (to bin resolution)
    xu = (rng.random(400_000) * 8000).astype(np.float32)
    hu, _ = np.histogram(xu, bins=4096, range=(0.0, 8000.0))
    bu = cfd.balanced_bounds(hu, 8000.0, 4, 300.0)
    assert np.allclose(bu, [0, 2000, 4000, 6000, 8000], atol=30)
    with pytest.raises(AssertionError):
        cfd.balanced_bounds(hist, 8000.0, 8, 1200.0)
