#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into a compact text summary for profiles/.

usage: python tools/summarize_ncu.py gpurun_out/prof_x.ncu-rep profiles/r01_x.txt [items_per_launch]
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",  # load/store data pipe: shared + global wavefronts
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    lines = [f"# summary of {rep} (ncu --set full --clock-control none; replayed, cold cache: use shares, not absolutes)"]
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    for d in data:
        lines.append("")
        lines.append("kernel: " + d[hdr.index("Kernel Name")])
        for m in METRICS:
            if m in hdr:
                lines.append(f"  {m:70s} {d[hdr.index(m)]} {units[hdr.index(m)]}")
    src = ncu_csv(rep, "source", ["--print-source", "sass"])
    if len(src) > 2:
        h = src[1]
        i_s, i_e, i_n = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        ops, samp, stalls = collections.Counter(), collections.Counter(), collections.Counter()
        for r in src[2:]:
            if len(r) <= i_e:
                continue
            m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[i_s])
            if not m:
                continue
            try:
                e, s = int(r[i_e]), int(r[i_n])
            except ValueError:
                continue
            ops[m.group(1)] += e
            samp[m.group(1)] += s
            for i in stall_cols:
                try:
                    stalls[h[i]] += int(r[i])
                except ValueError:
                    pass
        tot, ts = sum(ops.values()), max(1, sum(samp.values()))
        lines.append("")
        lines.append(f"warp-level instructions executed (all captured launches): {tot}")
        lines.append("opcode        share_of_instr  share_of_samples")
        for op, c in ops.most_common(16):
            lines.append(f"  {op:10s}  {100.0 * c / tot:6.2f}%        {100.0 * samp[op] / ts:6.2f}%")
        lines.append("stall reasons (pc samples): " + ", ".join(f"{k[6:]}={v}" for k, v in stalls.most_common(9)))
        sass = " ".join(r[i_s] for r in src[2:] if len(r) > i_s)
        lines.append("Blackwell packed-FP32 SASS present: " + ", ".join(k for k in ("FFMA2", "FADD2", "FMUL2", "DFMA") if k in sass))
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
