#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small, tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv profiles/launches_r1.csv
    python profiles/summarize.py full gpurun_out/prof_r1.ncu-rep profiles/ncu_fused_brgemm_r1.json [--dominant]

`launches`: per-kernel launch count / total / mean of gpu__time_duration.sum and each kernel's SHARE of the step
(ncu's times are cold-cache and serialised: the share is what should agree with bench.py, not the absolute).
`full`: the metrics of the recipe in /opt/skills/guides/B200_PROFILING.md from an `ncu --set full` report, averaged
over the captured launches; with --dominant also writes profiles/dominant_kernel.json (bench.py reads
dram_bytes_per_launch from it for roofline.traffic).
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__cluster_size", "launch__block_size", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if r and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name, val = r[4], float(r[-1])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += val
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ns", "mean_ns", "share_of_captured_time"])
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([name, n, round(t), round(t / n), round(t / total, 4)])
    print(open(dst).read())


def full(src, dst, dominant):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = {"source": src, "launches_captured": len(data), "kernel": data[0][hdr.index("Kernel Name")], "metrics": {}}
    for i, h in enumerate(hdr):
        if h in WANT:
            vals = [float(r[i].replace(",", "")) for r in data if r[i] not in ("", "n/a")]
            if vals:
                out["metrics"][h] = {"mean": sum(vals) / len(vals), "unit": units[i], "per_launch": vals}

    def mean(k, scale=1.0):
        m = out["metrics"].get(k)
        return m["mean"] * scale if m else None

    def to_bytes(k):
        m = out["metrics"].get(k)
        if not m:
            return None
        u = m["unit"].lower()
        f = 1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1.0
        return m["mean"] * f

    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    out["dram_bytes_per_launch"] = (rd or 0) + (wr or 0)
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({k: (v["mean"], v["unit"]) for k, v in out["metrics"].items()}, indent=1))
    print("dram bytes per launch:", out["dram_bytes_per_launch"])
    if dominant:
        json.dump({"kernel": out["kernel"], "dram_bytes_per_launch": out["dram_bytes_per_launch"], "from": dst},
                  open("profiles/dominant_kernel.json", "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], "--dominant" in sys.argv)
