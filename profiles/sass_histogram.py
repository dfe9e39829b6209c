"""Opcode histogram per kernel of libe2e_asr_b200.so (cuobjdump -sass): the evidence for which tensor-core / TMA / cluster
instructions each kernel really contains.  Usage: python profiles/sass_histogram.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "e2e_asr_b200", "libe2e_asr_b200.so")
# mnemonic prefixes worth reporting (B200_PROFILING.md): tcgen05 = UTCHMMA / UTCQMMA, TMEM loads = LDTM, TMA = UTMALDG /
# UBLKCP, legacy tensor pipe = HMMA / IMMA / DMMA, cluster barriers = UCGABAR / BAR / SYNCS, fp64 = DFMA / DADD / DMUL
WATCH = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "IMMA", "DMMA", "DFMA",
         "DADD", "DMUL", "MUFU", "SYNCS", "UCGABAR", "BAR", "ATOM", "RED", "LDGSTS", "SHFL", "F2FP", "FFMA", "LDS", "STS", "LDG",
         "STG", "MEMBAR", "FENCE", "ERRBAR", "CCTL")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*", "", name)
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    print("# cuobjdump -sass e2e_asr_b200/libe2e_asr_b200.so: instructions per kernel (static counts), opcode families of")
    print("# interest first with their full mnemonics; built by profiles/sass_histogram.py")
    for name, c in kernels.items():
        total = sum(c.values())
        fam = collections.Counter()
        for op, n in c.items():
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    fam[w] += n
        print("\n== %s  (%d instructions)" % (name, total))
        print("   families: " + ", ".join("%s %d" % (k, v) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])))
        tc = sorted(((op, n) for op, n in c.items() if op.split(".")[0] in
                     ("UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UBLKCP", "HMMA", "DMMA", "DFMA", "UCGABAR", "SYNCS", "UTCBAR")),
                    key=lambda kv: -kv[1])
        if tc:
            print("   detail:   " + ", ".join("%s x%d" % kv for kv in tc[:14]))


if __name__ == "__main__":
    main()
