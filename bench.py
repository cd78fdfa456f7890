#!/usr/bin/env python3
"""bench.py — particle-steps/s of the particle-life step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

Prints ONE JSON line (rank 0).  A "step" is one pass of the hot path (cell-list build -> pair
force -> fused integrate [-> proximity graph when the workload says so]) over all particles.

  value     whole-job particle-steps/s with the state resident in HBM, timed on the device with
            CUDA events on the engine's stream, per step, L2 flushed between timed steps
  e2e       the same metric through the stateless C-ABI call cf_step_host with pinned HOST
            buffers: H2D(particles, counts) -> step -> D2H(particles, counts) every step
  roofline  the pair-force kernel against the FP32 FMA peak (measured live)
  cpu_baseline  the oracle (CPU port of the reference law, cell list, OpenMP) on a bounded sample

`--impl reference` times the reference's own kernel (oracle/_ref: the unmodified
cuda-native/src/ParticleSimulation.cu compiled for sm_100a) on the same workload; the reference
has no CPU implementation, its implementation of this path IS a CUDA kernel.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-steps/s"

# Workloads = BASELINE.json configs restated as synthetic inputs.
WORKLOADS = {
    # config 3 (the metric's "1M particles on one B200"): eater.json law, non-zero radius
    # modifiers and ratio 0.5 so that "high ratio" actually widens the neighbourhoods
    "c3-eater-1M": dict(config=3, preset="eater", n_per_gpu=1_000_000, init="uniform",
                        radio=[1.0, 0.5, 0.0, 0.0, -0.5, 1.0], ratio=0.5, graph=None),
    "c3-pulser-1M": dict(config=3, preset="pulser", n_per_gpu=1_000_000, init="uniform",
                         radio=None, ratio=None, graph=None),
    # config 2 (README headline): defaults, default force matrix, graph on, reference spawn cube
    "c2-default-100k": dict(config=2, preset=None, n_per_gpu=100_000, init="spawn",
                            radio=None, ratio=None, graph=(200.0, 5)),
    "c2-default-100k-uniform": dict(config=2, preset=None, n_per_gpu=100_000, init="uniform",
                                    radio=None, ratio=None, graph=(200.0, 5)),
    # config 5 (16M = 2M per GPU weak scaling): settings.json law, 8 types, graph on
    "c5-settings-2M": dict(config=5, preset="settings", n_per_gpu=2_000_000, init="uniform",
                           radio=None, ratio=None, graph=(200.0, 5)),
    # config 1 shape on the GPU (parity-test size; CPU-only in BASELINE.json)
    "c1-settings-10k": dict(config=1, preset="settings", n_per_gpu=10_000, init="spawn",
                            radio=None, ratio=None, graph=None),
}
DEFAULT_WORKLOAD = "c3-eater-1M"


def workload_setup(name, n_gpus):
    """(lib Params, raw table or None, effective table, radio, n_total, seed, mode, graph)."""
    import cellflow_b200 as cf

    w = WORKLOADS[name]
    if w["preset"]:
        pr = cf.load_preset(os.path.join(ROOT, "presets", w["preset"] + ".json"))
        params = pr.params.copy()
        T = params.numParticleTypes
        radio = np.zeros(T, np.float32)
        radio[: pr.numRadio] = pr.radio
        raw = pr.raw_force
    else:
        params = cf.default_params()
        T = 6
        raw, radio, _ = cf.reference_default_tables(T)
    if w["radio"] is not None:
        radio = np.float32(w["radio"])
    if w["ratio"] is not None:
        params.ratio = w["ratio"]
    params.ratioWithLFO = params.ratio  # lfoA = 0 in every shipped preset
    # weak scaling: one 8000-wide block per GPU along x
    params.canvasWidth = params.canvasWidth * n_gpus
    n_total = w["n_per_gpu"] * n_gpus
    seed = 0x5EED0000 + w["config"]
    mode = cf.INIT_UNIFORM if w["init"] == "uniform" else cf.INIT_SPAWN_CUBE
    return params, raw, radio, n_total, seed, mode, w["graph"]


# ------------------------------------------------------------------------------------------------
# clocks (pynvml; nvidia-smi is not guaranteed in the image)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                r = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
GRAPH_PRIME_STEPS = 8


def run_ours(args):
    import torch
    import cellflow_b200 as cf
    from cellflow_b200 import _lib

    rank, world, local = dist_env()
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    if world > 1:
        from cellflow_b200 import dist as cfdist
        return cfdist.bench_multi(args, WORKLOADS, workload_setup)

    dev = local
    torch.cuda.set_device(dev)
    L = cf.lib()
    params, raw, radio, n, seed, mode, graph = workload_setup(args.workload, 1)
    T = params.numParticleTypes
    sim = cf.ParticleSimulation(n, T, device=dev, init=False)
    sim.params = params
    sim.setRadioByType(radio)
    sim.setRawForceTableValues(raw)
    sim.updateForceTable(params.forceRange, params.forceBias, params.forceOffset)
    sim.initializeParticles(seed=seed, mode=mode)
    if args.force_kernel:
        sim.setOption("force_kernel", args.force_kernel)
    if args.graph_kernel:
        sim.setOption("graph_kernel", args.graph_kernel)
    if args.no_graphs:
        sim.setOption("cuda_graphs", 0)
    for kv in args.opt:
        k, v = kv.split("=")
        sim.setOption(k, float(v))
    sim.setOption("timing", 2)   # whole-step events; the step itself replays a CUDA graph

    def one_step():
        sim.simulate(sync=False)
        if graph:
            sim.generateProximityGraph(graph[0], graph[1])

    # setup, untimed: the engine runs a step directly the first time it meets a parameter set (buffers
    # are sized there), captures it into a CUDA graph the second time and replays it from the third;
    # the two ping-pong parities are separate graphs.  Prime them so that neither the W warm-up steps
    # nor the K timed steps ever contain a capture, whatever W is.
    for _ in range(GRAPH_PRIME_STEPS):
        one_step()
    sim.sync()
    for _ in range(args.warmup):
        one_step()
    sim.sync()
    sim.statsReset()
    L2_FLUSH = 256 << 20
    with ClockSampler(dev) as clocks:
        torch.cuda.synchronize()
        wall0 = time.perf_counter()
        for _ in range(args.steps):
            _lib.check(L.cf_bench_flush_l2(C.c_int(dev), C.c_size_t(L2_FLUSH)))
            one_step()
            sim.sync()
        torch.cuda.synchronize()
        wall = time.perf_counter() - wall0
        st = sim.stats()
        graph_ms = st.ms_graph_total  # the library accumulates the device time of every graph build
    step_ms = st.ms_total / max(st.steps, 1) + graph_ms / args.steps
    value = n / (step_ms * 1e-3)
    launches = int(st.launches)
    # per-phase breakdown: a few extra steps launched kernel by kernel with events between phases
    sim.setOption("timing", 1)
    sim.statsReset()
    for _ in range(5):
        _lib.check(L.cf_bench_flush_l2(C.c_int(dev), C.c_size_t(L2_FLUSH)))
        one_step()
        sim.sync()
    st = sim.stats()
    force_ms = st.ms_force / max(st.steps, 1)
    sort_ms = st.ms_sort / max(st.steps, 1)
    integ_ms = st.ms_integrate / max(st.steps, 1)
    accepted, tested = int(st.accepted_pairs), int(st.tested_pairs)
    sim.setOption("timing", 2)

    # ---- roofline of the dominant kernel (pair force): 37 flop per accepted ordered pair -----
    tf = C.c_double(0)
    mhz = C.c_double(0)
    _lib.check(L.cf_bench_fp32_peak(C.c_int(dev), C.byref(tf), C.byref(mhz)))
    achieved = 37.0 * accepted / (force_ms * 1e-3) * 1e-12 if force_ms > 0 else 0.0
    traffic = None  # dram bytes of the force kernel per launch, from the committed ncu capture
    pipes = None    # pipe utilisation of the same capture (what the FP32-bound kernel actually loads)
    try:
        with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")) as f:
            t = json.load(f).get(args.workload)
        if t:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
            pipes = t.get("pipes_pct_of_peak_sustained_active")
    except Exception:
        pass
    peaks, peak_src = measured_peaks()
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    roofline = {
        "kernel": "pair_force", "bound": "fp32", "achieved": round(achieved, 3), "peak": round(tf.value, 2),
        "unit": "TFLOP/s", "frac": round(achieved / tf.value, 4) if tf.value else None, "traffic": traffic,
        "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)",
        "algorithmic_bytes_per_launch": 32 * n,  # read pos4 16 B + write frc4 16 B per particle
        "ncu_pipes_pct": pipes,  # from the committed ncu --set full capture of this workload (profiles/)
        "peak_source": "FFMA microbenchmark measured in this run (no FP32 figure in MEASURED_PEAKS.json)",
        "flops_per_accepted_pair": 37, "accepted_pairs_per_step": accepted, "tested_pairs_per_step": tested,
        "pair_tests_per_s": round(tested / (force_ms * 1e-3), 1) if force_ms > 0 else None,
        "tested_pairs_note": "pairs covered by the 27-cell stencil; the generation-4 kernel's box prefilter "
                             "decides about half of them without the exact per-pair test",
        "kernel_ms": round(force_ms, 4),
        "hbm_kernels": {
            "integrate": {"bytes_per_particle": 96, "ms": round(integ_ms, 4),
                          "achieved_GBs": round(96.0 * n / (integ_ms * 1e-3) * 1e-9, 1) if integ_ms > 0 else None},
            "cell_list_build": {"ms": round(sort_ms, 4)},
            "peak_GBs": hbm, "peak_source": peak_src},
    }

    # ---- e2e: stateless C-ABI call with pinned host buffers ---------------------------------
    pin_p = torch.empty(n * 44, dtype=torch.uint8).pin_memory()
    pin_c = torch.zeros(n, dtype=torch.int32).pin_memory()
    pout_p = torch.empty(n * 44, dtype=torch.uint8).pin_memory()
    pout_c = torch.zeros(n, dtype=torch.int32).pin_memory()
    cur = sim.getParticleData()
    pin_p.numpy()[:] = cur.view(np.uint8)
    pin_c.numpy()[:] = sim.getNeighborCounts()
    e2e_steps = max(3, min(args.steps, 10))
    bufs = [(pin_p, pin_c), (pout_p, pout_c)]

    def host_step(src, dst):
        _lib.check(L.cf_step_host(sim._h, C.byref(sim.params), C.c_void_p(src[0].data_ptr()),
                                  C.c_void_p(src[1].data_ptr()), C.c_void_p(dst[0].data_ptr()),
                                  C.c_void_p(dst[1].data_ptr()), C.c_int(n)))
        if graph:
            sim.generateProximityGraph(graph[0], graph[1])

    host_step(bufs[0], bufs[1])  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        host_step(bufs[(k + 1) & 1], bufs[k & 1])
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e = {"value": round(n / e2e_s, 1), "unit": METRIC, "h2d_bytes_per_step": n * 48,
           "d2h_bytes_per_step": n * 48 + (8 if graph else 0), "ms_per_step": round(e2e_s * 1e3, 4),
           "steps": e2e_steps, "api": "cf_step_host (pinned host AoS in/out)"}

    cpu = cpu_baseline(args.workload, budget_s=args.cpu_seconds) if not args.no_cpu else None

    out = {
        "metric": METRIC, "value": round(value, 1), "unit": METRIC, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(step_ms, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "particles": n, "types": T,
                   "canvas": [params.canvasWidth, params.canvasHeight, params.canvasDepth],
                   "radius": params.radius, "ratio": params.ratioWithLFO,
                   "mean_neighbours": round(accepted / n, 1), "grid": list(st.grid),
                   "graph": list(graph) if graph else None,
                   "l2": f"flushed between timed steps ({L2_FLUSH >> 20} MiB overwrite)",
                   "force_kernel": args.force_kernel or "auto", "force_kernel_used": int(st.force_kernel),
                   "setup_steps_before_warmup": GRAPH_PRIME_STEPS, "parallelism": "1 GPU"},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary(),
        "roofline": roofline, "cpu_baseline": cpu,
        "phases_ms": {"cell_list_build": round(sort_ms, 4), "pair_force": round(force_ms, 4),
                      "integrate": round(integ_ms, 4), "graph": round(graph_ms / args.steps, 4)},
        "wall_s_timed_region": round(wall, 3),
    }
    sim.close()
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port on the box's host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_baseline(workload, budget_s=15.0, threads=None):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O  # bench.py's cpu_baseline leg: one of the places allowed to use oracle/
    import cellflow_b200 as cf

    params, raw, radio, n, seed, mode, graph = workload_setup(workload, 1)
    T = params.numParticleTypes
    op = O.Params()
    C.memmove(C.byref(op), C.byref(params), C.sizeof(op))
    table = O.force_table(raw, T, params.forceRange, params.forceBias, params.forceOffset)
    state = O.init_particles(n, T, seed, mode, op.canvas)
    threads = threads or O.max_threads()
    O.set_sort_candidates(False)
    # calibrate on a small slice, then size the sample to the budget
    m0 = min(n, 2048)
    t0 = time.perf_counter()
    O.step_range(state, None, op, table, radio, 0, m0, threads)
    t_cal = time.perf_counter() - t0  # includes the grid build over all n particles
    t1 = time.perf_counter()
    O.step_range(state, None, op, table, radio, 0, m0, threads)
    per = max((time.perf_counter() - t1) / m0, 1e-9)
    m = int(min(n, max(m0, budget_s / per)))
    t2 = time.perf_counter()
    _, cnt, _ = O.step_range(state, None, op, table, radio, 0, m, threads)
    dt = time.perf_counter() - t2
    O.set_sort_candidates(True)
    return {"value": round(m / dt, 1), "unit": METRIC, "cores": threads, "kind": "port",
            "sample": f"oracle cell-list step (OpenMP, {threads} threads) of {m} of {n} particles "
                      f"against all {n}, 1 step, {dt:.1f} s; the reference's own O(N^2) loop would "
                      f"test {n} pairs per particle instead of ~{int(cnt[:m].mean() * 27 / 4.19)}",
            "calibration_s": round(t_cal, 2)}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CUDA kernel (oracle/_ref), same workload
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import cellflow_b200 as cf

    params, raw, radio, n, seed, mode, graph = workload_setup(args.workload, max(1, args.gpus))
    T = params.numParticleTypes
    op = O.Params()
    C.memmove(C.byref(op), C.byref(params), C.sizeof(op))
    table = O.force_table(raw, T, params.forceRange, params.forceBias, params.forceOffset)
    state = O.init_particles(n, T, seed, mode, op.canvas)
    base = {"impl": "reference", "metric": METRIC, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "particles": n, "types": T,
                       "canvas": [params.canvasWidth, params.canvasHeight, params.canvasDepth],
                       "radius": params.radius, "ratio": params.ratioWithLFO}}
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libcellflow_ref.so")
    use_cuda = os.path.exists(ref_so) and not args.ref_cpu
    if use_cuda:
        R = C.CDLL(ref_so)
        use_cuda = R.ref_device_count() > 0
    if use_cuda:
        # the reference's kernel costs n^2 pair tests per step; bound the run to ~3 minutes
        ms = C.c_float(0)
        counts = np.zeros(n, np.int32)
        probe_n = min(n, 100_000)
        rc = R.ref_simulate_racy(state[:probe_n].ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p),
                                 C.c_int(probe_n), C.byref(op), table.ctypes.data_as(C.c_void_p),
                                 radio.ctypes.data_as(C.c_void_p), C.c_int(1), C.c_int(1), None, None, C.byref(ms))
        if rc != 0:
            raise SystemExit("reference harness failed")
        est_ms = ms.value * (n / probe_n) ** 2
        steps, warm = args.steps, args.warmup
        budget_ms = 180e3
        if est_ms * (steps + warm) > budget_ms:
            warm = 1 if est_ms * 2 <= budget_ms else 0
            steps = max(1, int(budget_ms / est_ms) - warm)
        rc = R.ref_simulate_racy(state.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p), C.c_int(n),
                                 C.byref(op), table.ctypes.data_as(C.c_void_p), radio.ctypes.data_as(C.c_void_p),
                                 C.c_int(warm), C.c_int(steps), None, None, C.byref(ms))
        if rc != 0:
            raise SystemExit("reference harness failed")
        g_ms = 0.0
        if graph:
            nv = C.c_int(0)
            gms = C.c_float(0)
            colors = np.zeros(10, O.COLOR)
            verts = np.zeros((n * graph[1] * 2, 6), np.float32)
            R.ref_graph(state.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_int(T), C.c_float(graph[0]),
                        C.c_int(graph[1]), colors.ctypes.data_as(C.c_void_p), C.c_int(10),
                        verts.ctypes.data_as(C.c_void_p), C.c_int(len(verts)), C.byref(nv), C.byref(gms))
            g_ms = gms.value
        step_ms = ms.value + g_ms
        value = n / (step_ms * 1e-3)
        base.update({
            "value": round(value, 1), "ms_per_step": round(step_ms, 4), "steps_timed": steps, "warmup_done": warm,
            "cpu_baseline": {"value": round(value, 1), "unit": METRIC, "cores": 0, "kind": "reference",
                             "sample": f"reference simulateParticlesKernel (unmodified .cu, sm_100a) on 1 B200, "
                                       f"{steps} full step(s) of {n} particles (n^2 = {n * n:.2e} pair tests each)"
                                       + (", + generateProximityGraphKernel" if graph else "")
                                       + "; kernel time only (CUDA events), no D2H, no GL",
                             "device": "cuda:0"},
            "e2e": {"value": round(value, 1), "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
    else:
        cpu = cpu_baseline(args.workload, budget_s=max(10.0, args.cpu_seconds))
        base.update({"value": cpu["value"], "ms_per_step": round(n / cpu["value"] * 1e3, 3), "cpu_baseline": cpu,
                     "e2e": {"value": cpu["value"], "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--force-kernel", type=int, default=0, help="0 auto, 1 per-particle, 2 tile (generation 3), 3 tile (generation 4)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--graph-kernel", type=int, default=0, help="0 auto, 1 thread per particle, 2 warp per particle")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (experiments)")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel individually")
    ap.add_argument("--ref-cpu", action="store_true", help="reference arm on the CPU oracle port instead")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
