"""One-process-per-GPU glue for the slab-decomposed engine.

torch.distributed is used for plumbing only: rendezvous (RANK / WORLD_SIZE / MASTER_* from the
environment, as torch.distributed.run sets them), exchanging the CUDA IPC handles of the ranks'
mailboxes once at start-up, gathering results and reducing timings.  The per-step halo / migration
exchange is done inside the C library's own kernels, which store straight into the neighbours'
mailboxes over NVLink (csrc/kernels_slab.cuh, csrc/slab_host.inl): no collective call per step.

Everything here that does not touch a GPU (slab bounds, partition / gather of particle arrays,
max-over-ranks reductions, id broadcast) also runs on the `gloo` backend; tests/test_dist_cpu.py
covers it with world_size 2 on CPU.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time

import numpy as np

from . import _lib
from ._lib import PARTICLE


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend: str | None = None):
    """Initialises torch.distributed from the environment (idempotent).  Returns (rank, world)."""
    import torch
    import torch.distributed as dist

    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def slab_bound(width: float, r: int, world: int) -> np.float32:
    """x bound r of `world` slabs over [0, width): the same expression the C library evaluates
    (csrc/slab_host.inl: slab_bound), so host-side partitions agree bit for bit."""
    return np.float32(np.float64(np.float32(width)) * np.float64(r) / np.float64(world))


def slab_bounds_array(width: float, world: int, bounds=None) -> np.ndarray:
    """The world + 1 slab bounds: the uniform split (same expression as the C library) or explicit ones."""
    if bounds is not None:
        b = np.asarray(bounds, dtype=np.float32)
        assert len(b) == world + 1
        return b
    return np.array([slab_bound(width, r, world) for r in range(world + 1)], dtype=np.float32)


def slab_owner(x: np.ndarray, width: float, world: int, bounds=None) -> np.ndarray:
    """Owning rank of every x coordinate (x in [0, width))."""
    b = slab_bounds_array(width, world, bounds)
    owner = np.searchsorted(b, np.asarray(x, dtype=np.float32), side="right") - 1
    return np.clip(owner, 0, world - 1).astype(np.int32)


def partition(particles: np.ndarray, counts: np.ndarray, width: float, rank: int, world: int, bounds=None):
    """(particles, counts, ids) owned by `rank`: the subset whose x lies in its slab, ids =
    original indices."""
    owner = slab_owner(particles["pos"][:, 0], width, world, bounds)
    ids = np.nonzero(owner == rank)[0].astype(np.int32)
    return particles[ids], np.asarray(counts, np.int32)[ids], ids


def balanced_bounds(hist: np.ndarray, width: float, world: int, min_width: float) -> np.ndarray:
    """Slab bounds with (nearly) equal particle counts from a histogram of x over [0, width) — for clustered
    states, where the uniform split would put almost everything on one or two ranks (the reference's own spawn
    rule fills a centred 2000-wide cube, ParticleSimulation.cu:45-55).  Every slab stays at least `min_width`
    (one interaction radius) wide; bounds sit on histogram bin edges, bounds[0] = 0, bounds[world] = width."""
    hist = np.asarray(hist, dtype=np.float64)
    nb = len(hist)
    assert world * min_width <= width * (1 + 1e-6), "the box is too narrow for that many slabs of this radius"
    edges = np.linspace(0.0, float(width), nb + 1)
    cdf = np.concatenate([[0.0], np.cumsum(hist)])
    total = cdf[-1]
    b = np.zeros(world + 1, dtype=np.float64)
    b[world] = float(width)
    for r in range(1, world):
        want = total * r / world
        k = int(np.searchsorted(cdf, want, side="left"))
        x = edges[min(max(k, 0), nb)]
        x = max(x, b[r - 1] + min_width)                       # wide enough itself ...
        x = min(x, float(width) - (world - r) * min_width)     # ... and room for the slabs to its right
        b[r] = x
    out = b.astype(np.float32)
    out[0] = 0.0
    out[world] = np.float32(width)
    for r in range(world):  # float32 rounding must not eat the minimum width
        if out[r + 1] - out[r] < np.float32(min_width):
            out[r + 1] = np.nextafter(np.float32(out[r] + np.float32(min_width)), np.float32(np.inf))
    out[world] = np.float32(width)
    return out


def interaction_radius(params, radio) -> float:
    """R_max = radius * (1 + max(0, max_t radio_t * ratioWithLFO)) (ParticleSimulation.cu:108-110)."""
    a = 1.0 + max(0.0, float(np.max(np.asarray(radio, np.float64) * float(params.ratioWithLFO))))
    return float(params.radius) * a


def broadcast_bytes(data: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """Broadcast a byte string from `src` (works on gloo and nccl)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        return data
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def all_gather_bytes(data: bytes) -> list:
    """Every rank's byte string, in rank order (works on gloo and nccl)."""
    import torch.distributed as dist

    if not dist.is_initialized():
        return [data]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, data)
    return out


def connect_ring(sim, rank: int, world: int):
    """Exchange the mailbox handles and map the two ring neighbours' mailboxes into this rank."""
    if world == 1:
        return
    handles = all_gather_bytes(sim.mailboxHandle())
    sim.connect(handles[(rank - 1) % world], handles[(rank + 1) % world])
    barrier()  # every rank has mapped its neighbours before anybody starts stepping


def all_reduce_max(value: float) -> float:
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_reduce_sum(value: float) -> float:
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    import torch.distributed as dist

    if dist.is_initialized():
        dist.barrier()


def gather_particles(particles: np.ndarray, counts: np.ndarray, ids: np.ndarray, n_total: int):
    """All ranks contribute what they own; rank 0 returns (particles, counts) in original id
    order (None elsewhere)."""
    import torch.distributed as dist

    if not dist.is_initialized():
        out = np.zeros(n_total, PARTICLE)
        cnt = np.zeros(n_total, np.int32)
        out[ids] = particles
        cnt[ids] = counts
        return out, cnt
    payload = (particles.tobytes(), np.asarray(counts, np.int32).tobytes(), np.asarray(ids, np.int32).tobytes())
    gathered = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    if dist.get_rank() != 0:
        return None, None
    out = np.zeros(n_total, PARTICLE)
    cnt = np.zeros(n_total, np.int32)
    seen = 0
    for pb, cb, ib in gathered:
        i = np.frombuffer(ib, np.int32)
        out[i] = np.frombuffer(pb, PARTICLE)
        cnt[i] = np.frombuffer(cb, np.int32)
        seen += len(i)
    assert seen == n_total, f"ranks own {seen} particles in total, expected {n_total}"
    return out, cnt


def all_reduce_array(a: np.ndarray) -> np.ndarray:
    """Element-wise sum over the ranks of a float64 array."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        return np.asarray(a, np.float64)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(np.asarray(a, np.float64), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def rebalance(sim, radio, bins: int = 4096, margin: float = 1.002):
    """Moves the slab bounds to equal particle counts (all ranks call this together, between steps): global x
    histogram -> balanced_bounds -> every particle goes to the rank that owns its x under the new bounds ->
    cf_slab_set_bounds + upload.  Host-mediated (download, object all-gather, upload): a rare operation.
    Returns (bounds, owned count of this rank)."""
    import torch.distributed as dist

    rank, world, _ = env_rank()
    width = float(sim.params.canvasWidth)
    pp, cc, ii = sim.downloadOwned()
    hist, _ = np.histogram(pp["pos"][:, 0], bins=bins, range=(0.0, width))
    hist = all_reduce_array(hist)
    bounds = balanced_bounds(hist, width, world, interaction_radius(sim.params, radio) * margin)
    owner = slab_owner(pp["pos"][:, 0], width, world, bounds)
    if dist.is_initialized():
        outgoing = [(pp[owner == r].tobytes(), cc[owner == r].tobytes(), ii[owner == r].tobytes()) for r in range(world)]
        incoming = [None] * world
        gathered = [None] * world
        dist.all_gather_object(gathered, outgoing)
        incoming = [g[rank] for g in gathered]
        pp = np.concatenate([np.frombuffer(b[0], PARTICLE) for b in incoming])
        cc = np.concatenate([np.frombuffer(b[1], np.int32) for b in incoming])
        ii = np.concatenate([np.frombuffer(b[2], np.int32) for b in incoming])
        order = np.argsort(ii, kind="stable")   # deterministic upload order
        pp, cc, ii = pp[order], cc[order], ii[order]
    sim.setSlabBounds(bounds)
    sim.uploadOwned(pp, cc, ii)
    return bounds, len(ii)


def make_slab_sim(params, raw, radio, n_total, seed, mode, capacity_factor=1.5, force_table=None, bounds=None):
    """Creates this rank's simulation, joins the ring and spawns the global initial condition.  `bounds`: slab
    bounds for clustered states (balanced_bounds), the same array on every rank."""
    import torch

    from .sim import ParticleSimulation

    rank, world = init_process_group()
    _, _, local = env_rank()
    T = params.numParticleTypes
    sim = ParticleSimulation(0, T, device=local, init=False)
    sim.params = params
    sim.setRadioByType(radio)
    if force_table is not None:
        sim.setForceTable(force_table)
    else:
        sim.setRawForceTableValues(raw)
        sim.updateForceTable(params.forceRange, params.forceBias, params.forceOffset)
    capacity = int(n_total / world * capacity_factor) + 1024
    sim.setOption("global_particle_count", n_total)   # the cell grid derives from rank-invariant numbers only
    sim.commInit(rank, world, capacity)
    connect_ring(sim, rank, world)
    if bounds is not None:
        sim.setSlabBounds(bounds)
    if seed is not None:
        sim.initParticlesGlobal(n_total, seed, mode)
    return sim, rank, world
