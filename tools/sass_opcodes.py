"""Opcode census of the built library (cuobjdump -sass), per kernel and in total: the evidence that the hot kernels are
Blackwell-native (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UBLKCP = TMA, FFMA2 = packed fp32).

    python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "motif_b200", "lib", "libmotif_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF", "SYNCS", "FFMA2", "FADD2", "FMUL2", "HMMA",
         "MUFU.SIN", "MUFU.EX2", "LDCU.64", "LDGSTS", "REDG", "RED.", "ATOMG", "LDG.E.128", "STG.E.128", "LDS.128", "STS.128", "REDUX", "ELECT", "UCGABAR", "NANOSLEEP"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
    total = collections.Counter()
    for c in per.values():
        total.update(c)

    def watch(c):
        out = []
        for w in WATCH:
            n = sum(v for k, v in c.items() if k.startswith(w) or (w.endswith(".") and k.startswith(w[:-1] + ".")))
            if n:
                out.append(f"{w}={n}")
        return " ".join(out)

    print(f"cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (static instruction counts)")
    print(f"TOTAL instructions {sum(total.values())}: {watch(total)}\n")
    for name, c in per.items():
        short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
        print(f"{short}: {sum(c.values())} instr; {watch(c)}")


if __name__ == "__main__":
    sys.exit(main())
