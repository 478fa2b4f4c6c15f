#!/usr/bin/env python3
"""Per-role instruction and stall accounting of the warp-specialised K1-TC kernel from an ncu source page:
   ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_roles.py src.csv <pixels in the launch>
Roles are recognised by the SETMAXREG instructions that open each warpgroup's code region (producers, consumers,
MMA issuer + converters); the MMA issuer warp and the converter warps share the last region."""
import csv
import sys
from collections import Counter, defaultdict


def main():
    path, npx = sys.argv[1], float(sys.argv[2])
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    ins = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        ins.append(r)
    # region boundaries: USETMAXREG occurrences
    marks = [i for i, r in enumerate(ins) if "SETMAXREG" in r[col["Source"]]]
    # each setmaxnreg is a small retry loop: keep the first of each cluster
    starts = [m for k, m in enumerate(marks) if k == 0 or m - marks[k - 1] > 8]
    names = ["prologue", "producers (x-pass, split, tcgen05.st)", "consumers (tcgen05.ld, fold, store)", "MMA issuer (1 warp) + converters (3 warps: cp.async, luma)"]
    bounds = [0] + starts + [len(ins)]
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    total_exec = 0
    out = []
    for k in range(len(bounds) - 1):
        seg = ins[bounds[k]:bounds[k + 1]]
        name = names[k] if k < len(names) else "region %d" % k
        parts = [(name, seg)]   # the MMA issuer warp and the three converter warps share one setmaxnreg region
        for nm, sg in parts:
            ex = sum(float(r[col["Instructions Executed"]] or 0) for r in sg)
            samples = sum(float(r[col["# Samples"]] or 0) for r in sg)
            st = Counter()
            for r in sg:
                for c in stall_cols:
                    v = float(r[col[c]] or 0)
                    if v:
                        st[c] += v
            ops = Counter()
            for r in sg:
                op = r[col["Source"]].split()
                op = [t for t in op if not t.startswith("@")]
                if op:
                    ops[op[0].split(".")[0]] += float(r[col["Instructions Executed"]] or 0)
            out.append((nm, ex, samples, st, ops))
            total_exec += ex
    print("total warp instructions %.0f = %.2f lane-instr/px" % (total_exec, total_exec * 32 / npx))
    for nm, ex, samples, st, ops in out:
        print("\n%-62s %12.0f warp-instr  %6.2f lane-instr/px  %5.1f %%  samples %d" % (nm, ex, ex * 32 / npx, 100 * ex / total_exec, samples))
        print("   top opcodes: " + ", ".join("%s %.2f" % (o, n * 32 / npx) for o, n in ops.most_common(14)))
        tot = sum(st.values()) or 1
        print("   stalls: " + ", ".join("%s %.0f%%" % (c[6:], 100 * v / tot) for c, v in st.most_common(6)))


if __name__ == "__main__":
    main()
