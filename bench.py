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
  extra     the other BASELINE.json configurations, measured the same way in the same run (shorter):
            N = 1: c3-pulser-1M, c2-default-100k (the metric's "ms/step at 100 k", graph on),
                   c5-settings-2M (graph on), c4-pulser-4M, c1-settings-10k (+ the single-thread O(N^2)
                   CPU transcription BASELINE config 1 names)
            N > 1: c5-settings-2M weak scaling (2 M per GPU, graph on), c4-pulser-4M strong scaling

`--impl reference` times the reference's own kernel (oracle/_ref: the unmodified
cuda-native/src/ParticleSimulation.cu compiled for sm_100a) on the same workload; the reference
has no CPU implementation, its implementation of this path IS a CUDA kernel.  That
arm imports nothing of the product (presets via json, tables via oracle/).
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
                        radio=[1.0, 0.5, 0.0, 0.0, -0.5, 1.0], ratio=0.5, graph=None, scaling="weak"),
    "c3-pulser-1M": dict(config=3, preset="pulser", n_per_gpu=1_000_000, init="uniform",
                         radio=None, ratio=None, graph=None, scaling="weak"),
    # config 2 (README headline): defaults, default force matrix, graph on, reference spawn cube
    "c2-default-100k": dict(config=2, preset=None, n_per_gpu=100_000, init="spawn",
                            radio=None, ratio=None, graph=(200.0, 5), scaling="weak"),
    "c2-default-100k-uniform": dict(config=2, preset=None, n_per_gpu=100_000, init="uniform",
                                    radio=None, ratio=None, graph=(200.0, 5), scaling="weak"),
    # config 4 (4M strong scaling): pulser.json, ratio 0, the SAME 4 M particles in the SAME 8000^3 box on
    # 1/2/4/8 GPUs (slabs 8000/4000/2000/1000 wide)
    "c4-pulser-4M": dict(config=4, preset="pulser", n_total=4_000_000, init="uniform",
                         radio=None, ratio=None, graph=None, scaling="strong"),
    # config 5 (16M = 2M per GPU weak scaling): settings.json law, 8 types, graph on
    "c5-settings-2M": dict(config=5, preset="settings", n_per_gpu=2_000_000, init="uniform",
                           radio=None, ratio=None, graph=(200.0, 5), scaling="weak"),
    # config 1 shape on the GPU (parity-test size; CPU-only in BASELINE.json)
    "c1-settings-10k": dict(config=1, preset="settings", n_per_gpu=10_000, init="spawn",
                            radio=None, ratio=None, graph=None, scaling="weak"),
}
DEFAULT_WORKLOAD = "c3-eater-1M"
L2_FLUSH = 256 << 20

# SimulationParams.h:15-37 defaults (what a fresh reference object holds); cf_params field order
PARAM_FIELDS = ["radius", "delta_t", "friction", "repulsion", "attraction", "k", "balance", "canvasWidth",
                "canvasHeight", "canvasDepth", "spawnRegionSize", "numParticleTypes", "ratioWithLFO",
                "forceMultiplier", "maxExpectedNeighbors", "forceRange", "forceBias", "ratio", "lfoA", "lfoS",
                "forceOffset"]
PARAM_DEFAULTS = dict(radius=42.07, delta_t=0.18, friction=0.51, repulsion=64.83, attraction=3.06, k=29.45,
                      balance=0.79, canvasWidth=8000.0, canvasHeight=8000.0, canvasDepth=8000.0,
                      spawnRegionSize=2000.0, numParticleTypes=6, ratioWithLFO=0.0, forceMultiplier=2.33,
                      maxExpectedNeighbors=400, forceRange=0.28, forceBias=-0.20, ratio=0.0, lfoA=0.0, lfoS=0.1,
                      forceOffset=1.0)


def workload_spec(name, n_gpus):
    """Plain-Python description of a workload (no product or oracle import): parameters as a dict in the
    reference's names (loadPreset's key -> field mapping, CellFlowWidget.cpp:1087-1106), raw force table and
    radioByType (None = the reference's default rand() tables), counts, seed, init mode, graph."""
    w = WORKLOADS[name]
    params = dict(PARAM_DEFAULTS)
    raw = radio = None
    if w["preset"]:
        with open(os.path.join(ROOT, "presets", w["preset"] + ".json")) as f:
            d = json.load(f)
        for k in ("radius", "delta_t", "friction", "repulsion", "attraction", "k", "balance", "forceMultiplier",
                  "forceRange", "forceBias", "ratio", "lfoA", "lfoS", "forceOffset", "canvasWidth", "canvasHeight",
                  "canvasDepth", "spawnRegionSize"):
            if k in d:
                params[k] = float(np.float32(d[k]))
        T = int(d.get("numParticleTypes", 6))
        params["numParticleTypes"] = T
        radio = np.zeros(T, np.float32)
        r = np.array(d.get("radioByType", []), np.float32)[:T]
        radio[: len(r)] = r
        raw = np.array(d["rawForceTable"], np.float32)[: T * T]
    T = params["numParticleTypes"]
    if w["radio"] is not None:
        radio = np.float32(w["radio"])
    if w["ratio"] is not None:
        params["ratio"] = w["ratio"]
    params["ratioWithLFO"] = params["ratio"]  # lfoA = 0 in every shipped preset
    if w["scaling"] == "weak":
        # one 8000-wide block per GPU along x
        params["canvasWidth"] = params["canvasWidth"] * n_gpus
        n_total = w["n_per_gpu"] * n_gpus
    else:
        n_total = w["n_total"]
    return dict(name=name, params=params, T=T, raw=raw, radio=radio, n_total=n_total,
                seed=0x5EED0000 + w["config"], mode=1 if w["init"] == "uniform" else 0, graph=w["graph"],
                scaling=w["scaling"], n_gpus=n_gpus)


def config_dict(spec, parallelism):
    """The `config` object of the JSON line: the same keys from both arms."""
    p = spec["params"]
    return {"workload": spec["name"], "particles": spec["n_total"], "types": spec["T"],
            "canvas": [p["canvasWidth"], p["canvasHeight"], p["canvasDepth"]], "radius": p["radius"],
            "ratio": p["ratioWithLFO"], "graph": list(spec["graph"]) if spec["graph"] else None,
            "l2": f"flushed between timed steps ({L2_FLUSH >> 20} MiB overwrite)", "parallelism": parallelism}


def fill_params(obj, spec):
    for k in PARAM_FIELDS:
        setattr(obj, k, spec["params"][k])
    return obj


def cf_setup(spec):
    """(cf.Params, raw, radio) for the product arm."""
    import cellflow_b200 as cf

    params = fill_params(cf.Params(), spec)
    raw, radio = spec["raw"], spec["radio"]
    if raw is None:
        raw, radio0, _ = cf.reference_default_tables(spec["T"])
        radio = radio0 if radio is None else radio
    return params, raw, radio


def oracle_setup(spec):
    """(O.Params, effective table, radio) for the CPU / reference legs (oracle/ only)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O

    op = fill_params(O.Params(), spec)
    raw, radio = spec["raw"], spec["radio"]
    if raw is None:
        raw, radio0 = O.default_tables(spec["T"])
        radio = radio0 if radio is None else radio
    table = O.force_table(raw, spec["T"], op.forceRange, op.forceBias, op.forceOffset)
    return O, op, table, np.ascontiguousarray(radio, np.float32)


# ------------------------------------------------------------------------------------------------
# clocks (pynvml; nvidia-smi is not guaranteed in the image)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake"}

    def __init__(self, index, period=0.05):
        self.period = float(os.environ.get("CF_BENCH_CLOCK_PERIOD", period))
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
            self._stop.wait(self.period)

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
# our arm, one GPU
# ------------------------------------------------------------------------------------------------
GRAPH_PRIME_STEPS = 8


def make_sim(spec, dev, args):
    import cellflow_b200 as cf

    params, raw, radio = cf_setup(spec)
    sim = cf.ParticleSimulation(spec["n_total"], spec["T"], device=dev, init=False)
    sim.params = params
    sim.setRadioByType(radio)
    sim.setRawForceTableValues(raw)
    sim.updateForceTable(params.forceRange, params.forceBias, params.forceOffset)
    sim.initializeParticles(seed=spec["seed"], mode=spec["mode"])
    if args.force_kernel:
        sim.setOption("force_kernel", args.force_kernel)
    if args.graph_kernel:
        sim.setOption("graph_kernel", args.graph_kernel)
    if args.no_graphs:
        sim.setOption("cuda_graphs", 0)
    for kv in args.opt:
        k, v = kv.split("=")
        sim.setOption(k, float(v))
    return sim


def measure_single(spec, dev, args, steps, warmup, with_phases=True, first_counts=False):
    """Device-timed steps of one workload on one GPU.  Returns (sim, result dict)."""
    import cellflow_b200 as cf
    from cellflow_b200 import _lib

    L = cf.lib()
    sim = make_sim(spec, dev, args)
    graph = spec["graph"]
    n = spec["n_total"]

    def one_step():
        sim.simulate(sync=False)
        if graph:
            sim.buildGraphAsync(graph[0], graph[1])

    counts0 = None
    if first_counts:  # neighbour counts of the very first step, for the parity check against the CPU leg
        sim.simulate()
        counts0 = sim.getNeighborCounts()
    sim.setOption("timing", 2)   # whole-step events; the step itself replays a CUDA graph
    # setup, untimed: the engine runs a step directly the first time it meets a parameter set (buffers
    # are sized there), captures it into a CUDA graph the second time and replays it from the third;
    # the two ping-pong parities are separate graphs.  Prime them so that neither the W warm-up steps
    # nor the K timed steps ever contain a capture, whatever W is.
    for _ in range(GRAPH_PRIME_STEPS):
        one_step()
    sim.sync()
    for _ in range(warmup):
        one_step()
    sim.sync()
    sim.statsReset()
    import torch
    with ClockSampler(dev) as clocks:
        torch.cuda.synchronize()
        wall0 = time.perf_counter()
        for _ in range(steps):
            _lib.check(L.cf_bench_flush_l2(C.c_int(dev), C.c_size_t(L2_FLUSH)))
            one_step()
            sim.sync()
        torch.cuda.synchronize()
        wall = time.perf_counter() - wall0
        st = sim.stats()
        graph_ms = st.ms_graph_total  # the library accumulates the device time of every graph build
    step_ms = st.ms_total / max(st.steps, 1) + graph_ms / steps
    res = {"value": n / (step_ms * 1e-3), "ms_per_step": step_ms, "launches": int(st.launches),
           "graph_ms": graph_ms / steps, "wall": wall, "clocks": clocks.summary(), "counts0": counts0,
           "n_edges": sim.graphEdgeCount() if graph else None}
    if with_phases:
        # per-phase breakdown: a few extra steps launched kernel by kernel with events between phases
        sim.setOption("timing", 1)
        sim.statsReset()
        for _ in range(5):
            _lib.check(L.cf_bench_flush_l2(C.c_int(dev), C.c_size_t(L2_FLUSH)))
            one_step()
            sim.sync()
        st = sim.stats()
        k = max(st.steps, 1)
        res.update(force_ms=st.ms_force / k, sort_ms=st.ms_sort / k, integ_ms=st.ms_integrate / k,
                   accepted=int(st.accepted_pairs), tested=int(st.tested_pairs), grid=list(st.grid),
                   force_kernel=int(st.force_kernel))
        sim.setOption("timing", 2)
    return sim, res


def block_counters(sim, spec, dev):
    """Exact-tested pairs and evaluated pair-lanes of the tile kernel (instrumented variant, one extra step)."""
    try:
        sim.setOption("count_blocks", 1)
    except Exception:
        return None
    sim.setOption("timing", 1)
    sim.statsReset()
    sim.simulate()
    st = sim.stats()
    sim.setOption("count_blocks", 0)
    sim.setOption("timing", 2)
    if st.exact_tested_pairs <= 0:
        return None
    acc = max(int(st.accepted_pairs), 1)
    return {"exact_tested_pairs_per_step": int(st.exact_tested_pairs),
            "evaluated_pair_lanes_per_step": int(st.evaluated_pair_lanes),
            "exact_tests_per_accepted_pair": round(st.exact_tested_pairs / acc, 3),
            "evaluated_lanes_per_accepted_pair": round(st.evaluated_pair_lanes / acc, 3)}


def extra_single(name, dev, args, steps=8, warmup=3):
    spec = workload_spec(name, 1)
    sim, r = measure_single(spec, dev, args, steps, warmup)
    out = {"value": round(r["value"], 1), "unit": METRIC, "ms_per_step": round(r["ms_per_step"], 4), "steps": steps,
           "phases_ms": {"cell_list_build": round(r["sort_ms"], 4), "pair_force": round(r["force_ms"], 4),
                         "integrate": round(r["integ_ms"], 4), "graph": round(r["graph_ms"], 4)},
           "particles": spec["n_total"], "mean_neighbours": round(r["accepted"] / spec["n_total"], 1),
           "force_kernel_used": r["force_kernel"], "graph_edges": r["n_edges"]}
    sim.close()
    return out


def run_ours(args):
    import torch
    import cellflow_b200 as cf
    from cellflow_b200 import _lib

    rank, world, local = dist_env()
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    if world > 1:
        return run_multi(args)

    dev = local
    torch.cuda.set_device(dev)
    L = cf.lib()
    spec = workload_spec(args.workload, 1)
    n, T, graph = spec["n_total"], spec["T"], spec["graph"]
    sim, r = measure_single(spec, dev, args, args.steps, args.warmup, first_counts=not args.no_cpu)
    force_ms, sort_ms, integ_ms = r["force_ms"], r["sort_ms"], r["integ_ms"]
    accepted, tested = r["accepted"], r["tested"]

    # ---- roofline of the dominant kernel (pair force): 37 flop per accepted ordered pair -----
    tf = C.c_double(0)
    mhz = C.c_double(0)
    _lib.check(L.cf_bench_fp32_peak(C.c_int(dev), C.byref(tf), C.byref(mhz)))
    achieved = 37.0 * accepted / (force_ms * 1e-3) * 1e-12 if force_ms > 0 else 0.0
    traffic = None  # dram bytes of the force kernel per launch, from the committed ncu capture
    pipes = None    # pipe utilisation of the same capture (what the FP32-bound kernel actually loads)
    for prof in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", prof)) as f:
                t = json.load(f).get(args.workload)
            if t:
                traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
                pipes = t.get("pipes_pct_of_peak_sustained_active")
                break
        except Exception:
            pass
    peaks, peak_src = measured_peaks()
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    # cell-list build: key + 2-3 radix passes + reorder + bounds, algorithmic bytes per particle
    key_bits = max(1, int(np.ceil(np.log2(max(2, int(np.prod(r["grid"])) * 64)))))
    passes = -(-key_bits // 10)
    sort_bytes = 24 + passes * 20 + 76
    roofline = {
        "kernel": "pair_force", "bound": "fp32", "achieved": round(achieved, 3), "peak": round(tf.value, 2),
        "unit": "TFLOP/s", "frac": round(achieved / tf.value, 4) if tf.value else None, "traffic": traffic,
        "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)",
        "algorithmic_bytes_per_launch": 32 * n,  # read pos4 16 B + write frc4 16 B per particle
        "ncu_pipes_pct": pipes,  # from the committed ncu --set full capture of this workload (profiles/)
        "peak_source": "FFMA microbenchmark measured in this run (no FP32 figure in MEASURED_PEAKS.json)",
        "flops_per_accepted_pair": 37, "accepted_pairs_per_step": accepted, "tested_pairs_per_step": tested,
        "pair_tests_per_s": round(tested / (force_ms * 1e-3), 1) if force_ms > 0 else None,
        "tested_pairs_note": "pairs covered by the 27-cell stencil; block_counters says how many of them reach the "
                             "exact per-pair test and how many pair-lanes are evaluated",
        "block_counters": block_counters(sim, spec, dev),
        "kernel_ms": round(force_ms, 4),
        "hbm_kernels": {
            "integrate": {"bytes_per_particle": 96, "ms": round(integ_ms, 4),
                          "achieved_GBs": round(96.0 * n / (integ_ms * 1e-3) * 1e-9, 1) if integ_ms > 0 else None,
                          "note": "frc4 / pos4 were just touched by the force kernel and sit in L2: an L2 figure, not "
                                  "HBM (profiles/: dram__bytes of the integrate kernel)"},
            "cell_list_build": {"ms": round(sort_ms, 4), "bytes_per_particle": sort_bytes, "radix_passes": passes,
                                "achieved_GBs": round(sort_bytes * n / (sort_ms * 1e-3) * 1e-9, 1) if sort_ms > 0 else None,
                                "frac": round(sort_bytes * n / (sort_ms * 1e-3) * 1e-9 / hbm, 4) if sort_ms > 0 else None},
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
    ne = C.c_int(0)

    def host_step(src, dst):
        _lib.check(L.cf_step_host(sim._h, C.byref(sim.params), C.c_void_p(src[0].data_ptr()),
                                  C.c_void_p(src[1].data_ptr()), C.c_void_p(dst[0].data_ptr()),
                                  C.c_void_p(dst[1].data_ptr()), C.c_int(n)))
        if graph:  # what the reference's generateProximityGraph hands back: the vertex count (edges stay on the device)
            _lib.check(L.cf_build_graph(sim._h, C.c_float(graph[0]), C.c_int(graph[1]), C.byref(ne)))

    host_step(bufs[0], bufs[1])  # warm-up
    host_step(bufs[1], bufs[0])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        host_step(bufs[k & 1], bufs[(k + 1) & 1])
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e = {"value": round(n / e2e_s, 1), "unit": METRIC, "h2d_bytes_per_step": n * 48,
           "d2h_bytes_per_step": n * 48 + (4 if graph else 0), "ms_per_step": round(e2e_s * 1e3, 4),
           "steps": e2e_steps, "api": "cf_step_host (pinned host AoS in/out)"
                                      + (" + cf_build_graph (edge count read back)" if graph else "")}

    cpu = cpu_baseline(spec, budget_s=args.cpu_seconds) if not args.no_cpu else None
    parity = None
    if cpu is not None and r["counts0"] is not None:
        m = cpu.pop("_m")
        want = cpu.pop("_counts")
        parity = {"parity_checked": bool(np.array_equal(r["counts0"][:m], want)), "particles_compared": int(m),
                  "what": "neighbour counts of the first step from the spawn state, GPU vs the oracle's cpu_baseline sample"}
    sim.close()

    extra = {}
    if not args.no_extra and args.workload == DEFAULT_WORKLOAD:
        for name in ("c3-pulser-1M", "c2-default-100k", "c5-settings-2M", "c4-pulser-4M", "c1-settings-10k"):
            try:
                extra[name] = extra_single(name, dev, args)
            except Exception as e:  # an extra must never take the headline down
                extra[name] = {"error": str(e)[:200]}
        if not args.no_cpu:
            try:
                extra["c1-settings-10k"]["cpu_reference_single_thread"] = cpu_config1()
            except Exception as e:
                extra["c1-settings-10k"]["cpu_reference_single_thread"] = {"error": str(e)[:200]}

    out = {
        "metric": METRIC, "value": round(r["value"], 1), "unit": METRIC, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(spec, "1 GPU"),
        "details": {"mean_neighbours": round(accepted / n, 1), "grid": r["grid"],
                    "force_kernel": args.force_kernel or "auto", "force_kernel_used": r["force_kernel"],
                    "setup_steps_before_warmup": GRAPH_PRIME_STEPS},
        "e2e": e2e, "gpu_launches": r["launches"], "clocks": r["clocks"],
        "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        "phases_ms": {"cell_list_build": round(sort_ms, 4), "pair_force": round(force_ms, 4),
                      "integrate": round(integ_ms, 4), "graph": round(r["graph_ms"], 4)},
        "extra": extra,
        "wall_s_timed_region": round(r["wall"], 3),
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm, N GPUs (one rank per GPU, x slabs, peer-to-peer mailbox exchange)
# ------------------------------------------------------------------------------------------------
def measure_multi(spec, args, steps, warmup, with_e2e):
    """One workload on all ranks: free-running step loop (no host synchronisation inside the timed region),
    device-timed per rank and step with CUDA events, max over ranks."""
    import torch
    from cellflow_b200 import _lib
    from cellflow_b200 import dist as cfd

    rank, world, local = dist_env()
    L = _lib.lib()
    params, raw, radio = cf_setup(spec)
    n_total, graph = spec["n_total"], spec["graph"]
    sim, rank, world = cfd.make_slab_sim(params, raw, radio, n_total, spec["seed"], spec["mode"])
    if args.force_kernel:
        sim.setOption("force_kernel", args.force_kernel)
    for kv in args.opt:
        k, v = kv.split("=")
        sim.setOption(k, float(v))
    sim.setOption("timing", 1)

    def one_step():
        sim.simulate(sync=False)
        if graph:
            sim.buildGraphAsync(graph[0], graph[1])

    for _ in range(warmup):  # (the flush allocates its scratch on first use: not inside the timed region)
        _lib.check(L.cf_bench_flush_l2_async(sim._h, C.c_size_t(L2_FLUSH)))
        one_step()
    sim.sync()
    sim.statsReset()
    with ClockSampler(local) as clocks:  # (NVML start-up takes tens of ms and differs per rank: before the barrier)
        torch.cuda.synchronize()
        cfd.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            _lib.check(L.cf_bench_flush_l2_async(sim._h, C.c_size_t(L2_FLUSH)))
            one_step()
        host_enqueue_ms = (time.perf_counter() - t0) / steps * 1e3  # host time to enqueue a step (must stay below the device time)
        sim.sync()
        torch.cuda.synchronize()
        cfd.barrier()
        wall = time.perf_counter() - t0
        st = sim.stats()
        graph_ms = st.ms_graph_total  # accumulated inside the library: no per-step stats call (it synchronises)
    k = max(st.steps, 1)
    my_ms = st.ms_total / k + graph_ms / steps
    res = dict(
        step_ms=cfd.all_reduce_max(my_ms), owned=cfd.all_reduce_sum(float(st.n_owned)),
        owned_max=cfd.all_reduce_max(float(st.n_owned)),
        accepted=cfd.all_reduce_sum(float(st.accepted_pairs)), exch_ms=cfd.all_reduce_max(st.ms_exchange / k),
        exch_mig_ms=cfd.all_reduce_max(st.ms_exchange_migrants / k), exch_halo_ms=cfd.all_reduce_max(st.ms_exchange_halo / k),
        force_ms=cfd.all_reduce_max(st.ms_force / k), sort_ms=cfd.all_reduce_max(st.ms_sort / k),
        integ_ms=cfd.all_reduce_max(st.ms_integrate / k), graph_ms=cfd.all_reduce_max(graph_ms / steps),
        launches=cfd.all_reduce_sum(float(st.launches)), ghosts=cfd.all_reduce_sum(float(st.n_ghost)),
        wall=wall, clocks=clocks.summary(), wall_ms_per_step=cfd.all_reduce_max(wall / steps * 1e3),
        per_rank=[json.loads(b.decode()) for b in cfd.all_gather_bytes(json.dumps(
            {"step": round(my_ms, 4), "force": round(st.ms_force / k, 4), "sort": round(st.ms_sort / k, 4),
             "wait_migrants": round(st.ms_exchange_migrants / k, 4), "wait_halo": round(st.ms_exchange_halo / k, 4),
             "integrate": round(st.ms_integrate / k, 4), "graph": round(graph_ms / steps, 4),
             "host_enqueue": round(host_enqueue_ms, 4), "step_max": round(st.ms_step_max, 4),
             "exchange_max_step": round(st.ms_exchange_max, 4),
             "sm_mhz": clocks.summary()["sm_mhz"], "reasons": clocks.summary()["reasons"]}).encode())])
    if with_e2e:
        # e2e: every rank round-trips what it owns through PINNED host memory each step
        # (D2H particles+counts+ids -> H2D the same -> step), raw C-ABI calls on the pinned buffers
        cap = int(st.n_owned * 1.2) + 4096
        pin_p = torch.empty(cap * 44, dtype=torch.uint8).pin_memory()
        pin_c = torch.zeros(cap, dtype=torch.int32).pin_memory()
        pin_i = torch.zeros(cap, dtype=torch.int32).pin_memory()
        cnt = C.c_int(0)

        def round_trip():
            _lib.check(L.cf_download_particles_ids(sim._h, C.c_void_p(pin_p.data_ptr()), C.c_void_p(pin_c.data_ptr()),
                                                   C.c_void_p(pin_i.data_ptr()), C.c_int(cap), C.byref(cnt)))
            _lib.check(L.cf_upload_particles_ids(sim._h, C.c_void_p(pin_p.data_ptr()), C.c_void_p(pin_c.data_ptr()),
                                                 C.c_void_p(pin_i.data_ptr()), cnt))
            return cnt.value

        e2e_steps = max(3, min(steps, 8))
        round_trip()
        one_step()
        sim.sync()
        cfd.barrier()
        t0 = time.perf_counter()
        h2d = d2h = 0
        for _ in range(e2e_steps):
            m = round_trip()
            one_step()
            sim.sync()
            h2d += m * 52
            d2h += m * 52
        cfd.barrier()
        res.update(e2e_s=cfd.all_reduce_max((time.perf_counter() - t0) / e2e_steps),
                   h2d=cfd.all_reduce_sum(h2d / e2e_steps), d2h=cfd.all_reduce_sum(d2h / e2e_steps), e2e_steps=e2e_steps)
    sim.close()
    cfd.barrier()
    return res


def run_multi(args):
    import torch
    from cellflow_b200 import _lib
    from cellflow_b200 import dist as cfd

    rank, world = cfd.init_process_group("nccl")
    _, _, local = dist_env()
    torch.cuda.set_device(local)
    L = _lib.lib()
    spec = workload_spec(args.workload, world)
    r = measure_multi(spec, args, args.steps, args.warmup, with_e2e=True)
    tf = C.c_double(0)
    mhz = C.c_double(0)
    _lib.check(L.cf_bench_fp32_peak(C.c_int(local), C.byref(tf), C.byref(mhz)))
    extra = {}
    if not args.no_extra and args.workload == DEFAULT_WORKLOAD:
        for name in ("c5-settings-2M", "c4-pulser-4M"):
            try:
                sp = workload_spec(name, world)
                x = measure_multi(sp, args, max(5, min(args.steps, 10)), 3, with_e2e=False)
                extra[name] = {"value": round(sp["n_total"] / (x["step_ms"] * 1e-3), 1), "unit": METRIC,
                               "scaling": sp["scaling"], "particles": sp["n_total"], "n_gpus": world,
                               "ms_per_step": round(x["step_ms"], 4),
                               "phases_ms": {"cell_list_build_max": round(x["sort_ms"], 4), "pair_force_max": round(x["force_ms"], 4),
                                             "integrate_max": round(x["integ_ms"], 4), "graph_max": round(x["graph_ms"], 4),
                                             "exchange_max": round(x["exch_ms"], 4)},
                               "owned_max_over_mean": round(x["owned_max"] * world / max(x["owned"], 1), 3),
                               "per_rank_ms": x["per_rank"],
                               "mean_neighbours": round(x["accepted"] / max(x["owned"], 1), 1)}
            except Exception as e:
                extra[name] = {"error": str(e)[:300]}
    if rank == 0:
        n_total = spec["n_total"]
        step_ms, force_ms, accepted = r["step_ms"], r["force_ms"], r["accepted"]
        out = {
            "metric": METRIC, "value": round(n_total / (step_ms * 1e-3), 1), "unit": METRIC, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(step_ms, 4),
            "higher_is_better": True, "scaling": spec["scaling"], "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": config_dict(spec, f"{world} x-slabs, peer-to-peer mailbox halo+migration exchange per step "
                                        "(NVLink stores from the producing kernels, no collective, no host sync)"),
            "details": {"particles_per_gpu": n_total // world, "mean_neighbours": round(accepted / max(r["owned"], 1), 1),
                        "owned_max_over_mean": round(r["owned_max"] * world / max(r["owned"], 1), 3),
                        "ghost_particles": int(r["ghosts"]), "timing": "CUDA events per rank and step, max over ranks; "
                        "free-running step loop (one host synchronisation after the K timed steps)",
                        "wall_ms_per_step": round(r["wall_ms_per_step"], 4), "per_rank_ms": r["per_rank"]},
            "e2e": {"value": round(n_total / r["e2e_s"], 1), "unit": METRIC, "h2d_bytes_per_step": int(r["h2d"]),
                    "d2h_bytes_per_step": int(r["d2h"]), "ms_per_step": round(r["e2e_s"] * 1e3, 4), "steps": r["e2e_steps"],
                    "api": "cf_download_particles_ids -> cf_upload_particles_ids -> cf_step per rank, pinned host buffers"},
            "gpu_launches": int(r["launches"]), "clocks": r["clocks"],
            "roofline": {"kernel": "pair_force", "bound": "fp32",
                         "achieved": round(37.0 * accepted / (force_ms * 1e-3) * 1e-12, 3) if force_ms > 0 else None,
                         "peak": round(tf.value * world, 2), "unit": "TFLOP/s",
                         "frac": round(37.0 * accepted / (force_ms * 1e-3) * 1e-12 / (tf.value * world), 4)
                         if force_ms > 0 else None, "traffic": None,
                         "peak_source": "FFMA microbenchmark on rank 0 x n_gpus"},
            "phases_ms": {"cell_list_build_max": round(r["sort_ms"], 4), "pair_force_max": round(force_ms, 4),
                          "integrate_max": round(r["integ_ms"], 4), "exchange_max": round(r["exch_ms"], 4),
                          "exchange_migrants_max": round(r["exch_mig_ms"], 4), "exchange_halo_max": round(r["exch_halo_ms"], 4),
                          "note": "exchange_* are parts of cell_list_build (wait for + unpack of the two mailbox messages)"},
            "extra": extra,
            "cpu_baseline": None, "wall_s_timed_region": round(r["wall"], 3),
        }
        print(json.dumps(out), flush=True)
    cfd.barrier()


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port on the box's host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_baseline(spec, budget_s=15.0, threads=None):
    O, op, table, radio = oracle_setup(spec)  # bench.py's cpu_baseline leg: one of the places allowed to use oracle/
    n, T = spec["n_total"], spec["T"]
    state = O.init_particles(n, T, spec["seed"], spec["mode"], op.canvas)
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
    if dt < 0.4 * budget_s and m < n:  # the small calibration slice over-estimated the cost: one larger sample
        m = int(min(n, m * 0.8 * budget_s / max(dt, 1e-3)))
        t2 = time.perf_counter()
        _, cnt, _ = O.step_range(state, None, op, table, radio, 0, m, threads)
        dt = time.perf_counter() - t2
    O.set_sort_candidates(True)
    return {"value": round(m / dt, 1), "unit": METRIC, "cores": threads, "kind": "port",
            "sample": f"oracle cell-list step (OpenMP, {threads} threads) of {m} of {n} particles "
                      f"against all {n}, 1 step, {dt:.1f} s; the reference's own O(N^2) loop would "
                      f"test {n} pairs per particle instead of ~{int(cnt[:m].mean() * 27 / 4.19)}",
            "calibration_s": round(t_cal, 2), "_m": m, "_counts": cnt[:m].copy()}


def cpu_config1(budget_steps=3):
    """BASELINE config 1 as written: the single-threaded transcription of the reference's O(N^2) loop
    (ParticleSimulation.cu:86-133), settings.json, 10 k particles from the reference spawn cube; `budget_steps`
    of the 1,000 steps are timed (cost per step does not depend on the step index: 1e8 pair tests each)."""
    spec = workload_spec("c1-settings-10k", 1)
    O, op, table, radio = oracle_setup(spec)
    n = spec["n_total"]
    state = O.init_particles(n, spec["T"], spec["seed"], spec["mode"], op.canvas)
    counts = np.zeros(n, np.int32)
    t0 = time.perf_counter()
    for _ in range(budget_steps):
        state, counts, _ = O.step(state, counts, op, table, radio, "bruteforce", 1)
    dt = time.perf_counter() - t0
    return {"value": round(n * budget_steps / dt, 1), "unit": METRIC, "cores": 1, "kind": "port",
            "ms_per_step": round(dt / budget_steps * 1e3, 2),
            "sample": f"{budget_steps} of the 1,000 steps of config 1 (settings.json, {n} particles, all-pairs loop, "
                      f"1 thread): {n * n:.1e} pair tests per step"}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CUDA kernel (oracle/_ref), same workload.  Nothing of the product is
# imported here: presets are parsed with json (workload_spec), tables come from oracle/.
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return
    # the reference is a single-GPU program: at --gpus N > 1 it runs the single-GPU workload and says so
    spec = workload_spec(args.workload, 1)
    O, op, table, radio = oracle_setup(spec)
    n, T, graph = spec["n_total"], spec["T"], spec["graph"]
    state = O.init_particles(n, T, spec["seed"], spec["mode"], op.canvas)
    base = {"impl": "reference", "metric": METRIC, "unit": METRIC, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(spec, "1 GPU")}
    if args.gpus > 1:
        base["note"] = (f"launched with --gpus {args.gpus}: the reference has no multi-GPU path, this is its single-GPU "
                        f"run of the per-GPU workload ({n} particles)")
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libcellflow_ref.so")
    use_cuda = os.path.exists(ref_so) and not args.ref_cpu
    if use_cuda:
        R = C.CDLL(ref_so)
        use_cuda = R.ref_device_count() > 0
    if use_cuda:
        # the reference's kernel costs n^2 pair tests per step; bound the run to ~2 minutes
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
        budget_ms = 120e3
        if est_ms * (steps + warm) > budget_ms:
            warm = max(1, min(warm, int(0.15 * budget_ms / est_ms)))
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
            "value": round(value, 1), "ms_per_step": round(step_ms, 4), "steps": steps, "warmup": warm,
            "steps_requested": args.steps, "warmup_requested": args.warmup,
            "cpu_baseline": {"value": round(value, 1), "unit": METRIC, "cores": 0, "kind": "reference",
                             "sample": f"reference simulateParticlesKernel (unmodified .cu, sm_100a) on 1 B200, "
                                       f"{steps} full step(s) of {n} particles (n^2 = {n * n:.2e} pair tests each)"
                                       + (", + generateProximityGraphKernel" if graph else "")
                                       + "; kernel time only (CUDA events), no D2H, no GL",
                             "device": "cuda:0"},
            "e2e": {"value": round(value, 1), "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
    else:
        cpu = cpu_baseline(spec, budget_s=max(10.0, args.cpu_seconds))
        cpu.pop("_m", None)
        cpu.pop("_counts", None)
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
    ap.add_argument("--no-extra", action="store_true", help="headline workload only (no `extra` configurations)")
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
