#!/usr/bin/env python3
"""Opcode histogram per kernel of dct_carver_b200/libdctc.so (cuobjdump -sass): the mnemonics that prove the
Blackwell-native paths (UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UBLKCP = TMA, LDGSTS = cp.async,
SYNCS = mbarrier, UCGABAR = cluster barrier) plus the arithmetic mix.  Usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT",
       "STAS", "ATOMG", "BAR", "SHFL", "FFMA2", "FADD2", "FMUL2", "FFMA", "FMNMX3", "FMNMX", "F2FP", "FHADD", "IDP", "I2FP",
       "LDS", "STS", "LDG", "STG", "LDL", "STL"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dct_carver_b200", "libdctc.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*", "", name)
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    print("# SASS opcode counts per kernel, %s (static instruction counts, cuobjdump -sass, sm_100a)" % os.path.basename(lib))
    print("# key: UTCHMMA tcgen05.mma | UTCBAR tcgen05.commit | LDTM/STTM tcgen05.ld/st | UTMALDG/UBLKCP TMA bulk copies | LDGSTS cp.async |")
    print("#      SYNCS mbarrier | UCGABAR cluster barrier | STAS st.async (DSMEM) | FFMA2/FADD2/FMUL2 packed FP32x2\n")
    for name, c in kernels.items():
        total = sum(c.values())
        parts = ["%s %d" % (k, c[k]) for k in KEY if c.get(k)]
        print("%-70s total %6d | %s" % (name[:70], total, ", ".join(parts)))


if __name__ == "__main__":
    main()
