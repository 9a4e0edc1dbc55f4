#!/usr/bin/env python
"""Opcode counts per kernel from `cuobjdump -sass` of libb200wm.so (static counts, CPU only).
    python scripts/sass_digest.py > profiles/r02_sass_digest.txt
What to look for: UBLKCP (bulk-copy TMA, cp.async.bulk), SYNCS (mbarrier), FFMA2 / FMUL2 / FADD2 (packed FP32, sm_100),
VIADDMNMX (DPX 16-bit saturating add), IDP (dp4a), PRMT; and the absence of local memory (LDL / STL) on the hot kernels."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "video-fingerprinting_b200", "lib", "libb200wm.so")
WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "VIADDMNMX", "IDP", "PRMT", "I2F", "F2I", "MUFU",
         "DFMA", "DMUL", "LDL", "STL", "LDG", "STG", "LDS", "STS", "ATOM", "ATOMS", "RED", "BAR", "CALL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            kernels[name][m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS opcode digest of {os.path.relpath(LIB, ROOT)} (sm_100a), static instruction counts per kernel")
    print("# columns: total | " + " ".join(WATCH))
    for (mangled, c), nice in zip(kernels.items(), demangle):
        nice = re.sub(r"\(.*", "", nice).replace("b200wm::", "")
        if sum(c.values()) < 40:
            continue
        print(f"{nice}\n    total {sum(c.values())} | " + " ".join(f"{k} {c[k]}" for k in WATCH if c[k]))


if __name__ == "__main__":
    sys.exit(main())
