"""Shared-memory wavefronts per SASS instruction class of one kernel in an ncu --set full capture (source page):
which loads / stores carry the wavefronts, and which of them are excessive (bank conflicts).

    python tools/ncu_smem.py rep.ncu-rep kernel-substring [items]     (items: divide the totals by this count)
"""
import collections, csv, subprocess, sys

rep, want = sys.argv[1], sys.argv[2]
items = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
hdr, cur, data = None, None, []
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = r[1]
    elif r and r[0] == "Address":
        hdr = r
    elif hdr and len(r) == len(hdr) and (cur is None or want in cur):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
tot = [0, 0, 0, 0]
for r in data:
    w = int(r[ix["L1 Wavefronts Shared"]] or 0)
    if not w:
        continue
    op = r[ix["Source"]].split()
    op = op[1] if op[0].startswith("@") else op[0]
    vals = (int(r[ix["Instructions Executed"]]), w, int(r[ix["L1 Wavefronts Shared Ideal"]] or 0), int(r[ix["L1 Wavefronts Shared Excessive"]] or 0))
    for i, v in enumerate(vals):
        agg[op][i] += v
        tot[i] += v
print(f"{'opcode':28s} {'warp inst':>12s} {'wavefronts':>12s} {'ideal':>12s} {'excessive':>12s}   (per item: / {items:g})")
for op, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{op:28s} " + " ".join(f"{x / items:12.1f}" for x in v))
print(f"{'total':28s} " + " ".join(f"{x / items:12.1f}" for x in tot))
