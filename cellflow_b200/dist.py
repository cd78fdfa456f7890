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


def slab_owner(x: np.ndarray, width: float, world: int) -> np.ndarray:
    """Owning rank of every x coordinate (x in [0, width))."""
    bounds = np.array([slab_bound(width, r, world) for r in range(world + 1)], dtype=np.float32)
    owner = np.searchsorted(bounds, np.asarray(x, dtype=np.float32), side="right") - 1
    return np.clip(owner, 0, world - 1).astype(np.int32)


def partition(particles: np.ndarray, counts: np.ndarray, width: float, rank: int, world: int):
    """(particles, counts, ids) owned by `rank`: the subset whose x lies in its slab, ids =
    original indices."""
    owner = slab_owner(particles["pos"][:, 0], width, world)
    ids = np.nonzero(owner == rank)[0].astype(np.int32)
    return particles[ids], np.asarray(counts, np.int32)[ids], ids


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


def make_slab_sim(params, raw, radio, n_total, seed, mode, capacity_factor=1.5, force_table=None):
    """Creates this rank's simulation, joins the ring and spawns the global initial condition."""
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
    if seed is not None:
        sim.initParticlesGlobal(n_total, seed, mode)
    return sim, rank, world
