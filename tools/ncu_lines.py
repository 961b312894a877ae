"""Per-source-line instruction and stall-sample shares of the kernels in an .ncu-rep captured with
`ncu --set full --import-source on` (kernels compiled with -lineinfo).

    python tools/ncu_lines.py gpurun_out/prof_r01_c3.ncu-rep [kernel-name-substring] [top-n]
"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                         text=True).stdout
    cur_file, cur_fn = None, None
    agg = collections.defaultdict(lambda: collections.OrderedDict())
    for r in csv.reader(out.splitlines()):
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif len(r) >= 2 and r[0] == "Function Name":
            cur_fn = r[1]
        elif len(r) >= 8 and r[0].isdigit():
            try:
                inst, samp = int(r[7]), int(r[6])
            except ValueError:
                continue
            a = agg[cur_fn].setdefault((cur_file, int(r[0]), r[1].strip()[:100]), [0, 0])
            a[0] += inst
            a[1] += samp
    for fn, lines in agg.items():
        if want not in fn:
            continue
        tot = sum(v[0] for v in lines.values()) or 1
        tots = sum(v[1] for v in lines.values()) or 1
        print(f"== {fn[:110]}\n   warp instructions {tot}, stall samples {tots}")
        for k, v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
            print(f"   {k[0]:20s}:{k[1]:4d}  inst {100 * v[0] / tot:5.1f}%  samples {100 * v[1] / tots:5.1f}%  {k[2]}")


if __name__ == "__main__":
    main()
