#!/usr/bin/env python
"""Turn the ncu outputs of tools/gpu_artifacts.sh (gpurun_out/) into the tracked summaries under profiles/.

usage: python tools/make_profiles.py [tag]      (default tag r01)
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")

SHORT = {"iir_overlap4_kernel": "iir_overlap", "iir_overlap_kernel": "iir_overlap", "logmel_power_kernel": "logmel_power",
         "logmel_finalize_kernel": "logmel_finalize", "gather_kernel": "gather", "trim_index_hop4_kernel": "trim_index",
         "trim_index_hop_kernel": "trim_index", "trim_index_kernel": "trim_index", "fbank_kernel": "fbank",
         "logmel_init_stats_kernel": "logmel_init_stats", "fbank_zero_pad_kernel": "fbank_zero_pad",
         "trim_frame_power_kernel": "trim_power", "pcm16_decode_kernel": "pcm16_decode"}


def short(name):
    base = name.split("(")[0].replace("void ", "").replace("hmfe::", "").split("<")[0].strip()
    return SHORT.get(base, base)


def launches():
    src = os.path.join(G, f"launches_{TAG}.csv")
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hi], rows[hi + 1:]
    kn, mv, mn = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    gs, bs = h.index("Grid Size"), h.index("Block Size")
    agg = collections.OrderedDict()
    lines = ["id,kernel,grid,block,gpu_time_us"]
    for r in data:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        us = float(r[mv].replace(",", "")) / 1e3
        k = short(r[kn])
        agg.setdefault(k, []).append(us)
        lines.append(f"{r[0]},{k},{r[gs].replace(',', ' ')},{r[bs].replace(',', ' ')},{us:.2f}")
    tot = sum(sum(v) for v in agg.values())
    head = [f"# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e` (c2, 5272 clips),",
            "# --metrics gpu__time_duration.sum --clock-control none -k regex:hmfe; replayed one kernel at a time, cold",
            "# caches: use the SHARES, not the absolute times.  kernel: launches, mean us, share of the summed time"]
    for k, v in agg.items():
        head.append(f"#   {k:22s} n={len(v):3d}  mean={sum(v) / len(v):9.1f} us  share={100 * sum(v) / tot:5.1f} %")
    open(os.path.join(OUT, f"{TAG}_launches_c2.csv"), "w").write("\n".join(head + lines) + "\n")
    print("\n".join(head))


def summarize(rep, dst):
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_ncu.py"), rep, dst], check=True, stdout=subprocess.DEVNULL)


def traffic():
    out = {}
    for wl, clips in (("c2", 5272), ("c1", 1000), ("c3", 1000)):
        rep = os.path.join(G, f"prof_{TAG}_{wl}.ncu-rep")
        if not os.path.exists(rep):
            continue
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        h, units = rows[0], rows[1]
        for d in rows[2:]:
            def val(m):
                i = h.index(m)
                x = float(d[i].replace(",", ""))
                u = units[i].lower()
                return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            k = short(d[h.index("Kernel Name")])
            out.setdefault(wl, {})[k] = {
                "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                "clips_per_launch": clips, "source": f"ncu --set full --clock-control none, one launch ({os.path.basename(rep)})"}
    json.dump(out, open(os.path.join(OUT, f"{TAG}_traffic.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1)[:1500])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    launches()
    for wl in ("c2", "c1", "c3"):
        rep = os.path.join(G, f"prof_{TAG}_{wl}.ncu-rep")
        if os.path.exists(rep):
            summarize(rep, os.path.join(OUT, f"{TAG}_ncu_full_{wl}.txt"))
    traffic()
