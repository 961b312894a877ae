"""Stall reasons, opcode mix and hottest instructions of ONE address range (role) of a warp-specialised kernel in an
ncu --set full capture.  The range is [first instruction matching START, first instruction matching END after it).

    python tools/ncu_role.py rep.ncu-rep START_REGEX END_REGEX [items]
    e.g. FFT role of logmel_tc_kernel:  USETMAXREG.TRY_ALLOC  "BAR.SYNC"
"""
import collections, csv, re, subprocess, sys

rep, start_re, end_re = sys.argv[1], sys.argv[2], sys.argv[3]
items = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
hdr, data = None, []
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Address":
        hdr = r
    elif hdr and len(r) == len(hdr):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
src = [r[ix["Source"]] for r in data]
start = next(i for i, s in enumerate(src) if re.search(start_re, s))
end = next(i for i in range(start + 1, len(src)) if re.search(end_re, src[i]))
rows = data[start:end]
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg, ops = collections.Counter(), collections.Counter()
inst = samp = 0
for r in rows:
    n = int(r[ix["Instructions Executed"]])
    inst += n
    samp += int(r[ix["# Samples"]])
    op = r[ix["Source"]].split()
    ops[(op[1] if op[0].startswith("@") else op[0]).split(".")[0]] += n
    for k in reasons:
        agg[k[6:]] += int(r[ix[k]])
tot_s = sum(int(r[ix["# Samples"]]) for r in data)
print(f"range [{start}, {end}) of {len(data)} instructions: {inst / items:.0f} warp instructions per item, {samp} of {tot_s} samples")
print("stalls: " + ", ".join(f"{k}={100 * v / samp:.0f}%" for k, v in agg.most_common(12)))
print("opcodes per item: " + ", ".join(f"{k}={v / items:.0f}" for k, v in ops.most_common(30)))
print("hottest instructions:")
for r in sorted(rows, key=lambda r: -int(r[ix["# Samples"]]))[:25]:
    top = max(reasons, key=lambda k: int(r[ix[k]]))
    print(f"  {100 * int(r[ix['# Samples']]) / samp:5.1f}%  {top[6:]:12s} {r[ix['Source']][:90]}")
