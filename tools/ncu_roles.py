"""Stall-reason breakdown of an ncu --set full capture per ADDRESS RANGE of a kernel (warp-specialised kernels: one range
per role).  Ranges are split at instructions matching the given regexes, or printed as N equal blocks.

    python tools/ncu_roles.py rep.ncu-rep kernel-substring [block_instructions]
"""
import csv, subprocess, sys, collections

rep, want = sys.argv[1], sys.argv[2]
block = int(sys.argv[3]) if len(sys.argv) > 3 else 400
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, cur, data = None, None, []
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = r[1]
    elif r and r[0] == "Address":
        hdr = r
    elif hdr and cur and want in cur and len(r) == len(hdr):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"{len(data)} instructions; columns: inst = warp instructions executed; samples by stall reason")
tot_s = sum(int(r[ix['# Samples']]) for r in data) or 1
for b0 in range(0, len(data), block):
    blk = data[b0 : b0 + block]
    inst = sum(int(r[ix["Instructions Executed"]]) for r in blk)
    samp = sum(int(r[ix["# Samples"]]) for r in blk)
    agg = collections.Counter()
    for r in blk:
        for k in reasons:
            agg[k[6:]] += int(r[ix[k]])
    top = ", ".join(f"{k}={v}" for k, v in agg.most_common(6) if v)
    ops = collections.Counter(r[ix["Source"]].split()[0] for r in blk if int(r[ix["# Samples"]]) > 0.02 * samp)
    print(f"[{b0:5d},{b0 + len(blk):5d}) inst {inst/1e6:7.2f}M samples {samp:6d} ({100*samp/tot_s:4.1f}%)  {top}   hot: {dict(ops)}")
