"""Turn the raw ncu artefacts of a GPU round (gpurun_out/<tag>_*) into small tracked summaries under profiles/.
    python tools/summarize_ncu.py <tag> [name ...]
A full capture gpurun_out/<tag>_prof_<name>.ncu-rep may come with gpurun_out/<tag>_prof_<name>.shape.json (written by
tools/profile_target.py: precision, rows, steps of the profiled launch); the summary then gets a companion
profiles/<tag>_ncu_<name>.meta.json (kernel, shape, DRAM bytes, tensor-pipe %, duration, commit), which bench.py reads for
`roofline.traffic`.
"""
import collections
import csv
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.environ.get("GPB_SUMMARY_DIR", os.path.join(ROOT, "profiles"))

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
]


def launches(tag):
    path = os.path.join(OUT, f"{tag}_launches.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(PROF, f"{tag}_launches_summary.csv"), "w") as f:
        f.write("kernel,launches,total_us,mean_us,share_pct\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k[:110]}\",{n},{t:.1f},{t / n:.1f},{100 * t / tot:.2f}\n")
    print("wrote launches summary,", len(agg), "kernels")


def full(tag, name):
    rep = os.path.join(OUT, f"{tag}_prof_{name}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(PROF, f"{tag}_ncu_{name}.csv"), "w") as f:
        f.write("kernel,metric,unit,value\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            kn = d["Kernel Name"][:80]
            for k in KEYS:
                if k in d:
                    f.write(f"\"{kn}\",{k},{units[hdr.index(k)]},{d[k]}\n")
            for k in hdr:
                if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") or \
                        (k.startswith("smsp__average_warp_latency_issue_stalled") and k.endswith(".ratio")):
                    f.write(f"\"{kn}\",{k},{units[hdr.index(k)]},{d[k]}\n")
    print("wrote full-capture summary for", name, len(rows) - 2, "launches")
    shape_path = os.path.join(OUT, f"{tag}_prof_{name}.shape.json")
    if os.path.exists(shape_path) and len(rows) > 2:
        d = dict(zip(hdr, rows[-1]))                      # the last captured launch

        def val(k, scale_units=None):
            if k not in d:
                return None
            v = float(d[k].replace(",", ""))
            u = units[hdr.index(k)]
            if scale_units:
                v *= scale_units.get(u, 1.0)
            return v
        byte_units = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        time_units = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
        meta = json.load(open(shape_path))
        try:
            commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
        except Exception:
            commit = None
        meta.update({"file": f"profiles/{tag}_ncu_{name}.csv", "kernel": d["Kernel Name"], "commit": commit or meta.get("commit"),
                     "order": int(time.time()),
                     "dram_bytes": int((val("dram__bytes_read.sum", byte_units) or 0) + (val("dram__bytes_write.sum", byte_units) or 0)),
                     "tensor_pipe_active_pct": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                     "duration_ms": val("gpu__time_duration.sum", time_units)})
        json.dump(meta, open(os.path.join(PROF, f"{tag}_ncu_{name}.meta.json"), "w"), indent=1)
        print("wrote", f"{tag}_ncu_{name}.meta.json", meta)


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    for n in (sys.argv[2:] or ("sampler", "encoder", "tc_sampler", "tc_ode_sampler", "tc_sampler_sat", "energy_rank_pool")):
        full(tag, n)
