#!/usr/bin/env python
"""Per-kernel SASS opcode inventory of libhmfe.so (cuobjdump -sass) for profiles/: which kernels contain the
Blackwell-specific instructions (tcgen05 = UTC*MMA / LDTM / STTM / UTCBAR, bulk async copies = UBLKCP / UTMALDG,
packed FP32 = FFMA2 / FADD2 / FMUL2, cp.async = LDGSTS, mbarrier = SYNCS, multimem) and how many of each.

usage: python tools/sass_opcodes.py [lib.so] > profiles/r02_sass_opcodes.txt
"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "heart_murmur_detection_b200/libhmfe.so"
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTCMMA", "LDTM", "STTM", "UTCBAR", "UTCATOM", "UBLKCP", "UTMALDG",
         "UTMASTG", "UBLKPF", "SYNCS", "LDGSTS", "FFMA2", "FADD2", "FMUL2", "DFMA", "DADD", "DMUL", "F2F", "HMMA", "MUFU",
         "REDUX", "CREDUX", "ATOMG", "RED", "MULTIMEM", "USETMAXREG", "ELECT", "STL", "LDL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    op_re = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)")
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = op_re.match(line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["__total__"] += 1
            if m.group(1) in ("LDG", "STG", "LDS", "STS") and ".128" in m.group(2):
                cur[m.group(1) + ".128"] += 1
    names = demangle(list(kernels))
    print(f"# SASS opcode inventory of {LIB} (cuobjdump -sass, sm_100a); static instruction counts per kernel")
    print("# tcgen05: UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit); bulk async copy: UBLKCP")
    print("# (cp.async.bulk), UTMALDG (cp.async.bulk.tensor); mbarrier: SYNCS; cp.async: LDGSTS; packed FP32: FFMA2 / FADD2 / FMUL2")
    tot = collections.Counter()
    for k, c in kernels.items():
        hits = [(w, c[w]) for w in WATCH if c[w]]
        for w, n in hits:
            tot[w] += n
        wide = " ".join(f"{w}={c[w]}" for w in ("LDG.128", "STG.128", "LDS.128", "STS.128") if c[w])
        print(f"\n{names.get(k, k)}\n  instructions {c['__total__']}  " + " ".join(f"{w}={n}" for w, n in hits) + ("  | " + wide if wide else ""))
    print("\n# whole library: " + " ".join(f"{w}={tot[w]}" for w in WATCH if tot[w]))


if __name__ == "__main__":
    main()
