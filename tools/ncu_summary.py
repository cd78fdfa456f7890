"""profiles/r02_ncu_traffic.json from the raw exports of tools/r02_profile_final.sh
(`ncu -i <rep> --page raw --csv`): usage  python tools/ncu_summary.py <workload>=<raw.csv> ...  > profiles/r02_ncu_traffic.json"""
import csv, json, sys


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


out = {}
for arg in sys.argv[1:]:
    wl, path = arg.split("=", 1)
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}

    def g(name, scale=None):
        v, u = m[name]
        x = num(v)
        if scale and x is not None:
            x *= scale.get(u, 1.0)
        return x

    byte_scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
    time_scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    out[wl] = {
        "kernel": m["Kernel Name"][0][:60],
        "capture": "ncu --set full --clock-control none --import-source on -k regex:force_ -s 3 -c 1, python bench.py --steps 2 "
                   "--warmup 3 --no-cpu --no-extra --workload " + wl,
        "duration_ms": g("gpu__time_duration.sum", time_scale),
        "dram_bytes_read": g("dram__bytes_read.sum", byte_scale),
        "dram_bytes_write": g("dram__bytes_write.sum", byte_scale),
        "warp_instructions": g("smsp__inst_executed.sum"),
        "registers_per_thread": g("launch__registers_per_thread"),
        "threads_per_warp_instruction": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "warps_active_pct_of_peak": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "smsp_cycles_active_over_elapsed": round(g("smsp__cycles_active.avg") / g("smsp__cycles_elapsed.max"), 4)
        if "smsp__cycles_elapsed.max" in m else None,
        "l2_hit_rate_pct": g("lts__t_sector_hit_rate.pct"),
        "pipes_pct_of_peak_sustained_active": {
            "fma": g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),  # cycles: a packed FFMA2 holds the pipe twice
            "alu": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "xu": g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
            "lsu": g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
            "issue_active": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        },
        "raw_export": "profiles/r02_ncu_force_%s_raw.csv" % wl,
    }
json.dump(out, sys.stdout, indent=1)
print()
