#!/usr/bin/env python
"""Opcode histogram per kernel of libsfmmatch.so (cuobjdump -sass): the ISA proof that the hot path is hand-written
tcgen05 / TMEM / TMA code (UTC*MMA, LDTM/STTM, UTMALDG/UBLKCP) and XOR+POPC (LOP3/POPC).  Runs without a GPU.

    python tools/sass_histogram.py > profiles/sass_r02.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sfm_danpipeline_b200", "csrc", "libsfmmatch.so")
KEY = ["UTCIMMA", "UTCHMMA", "UTCOMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "ELECT", "POPC", "LOP3", "VIMNMX", "VIMNMX3",
       "IMAD", "FFMA", "REDUX", "ATOMS", "ATOMG", "HMMA", "IMMA", "LDS", "LDG", "STG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    names = list(kernels)
    try:
        dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, dem))
    except Exception:
        pass
    total = collections.Counter()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a)  -- opcode counts per kernel, selected mnemonics first\n")
    for k, c in kernels.items():
        total.update(c)
        name = demangle.get(k, k)
        name = re.sub(r"\(.*", "", name)
        sel = "  ".join(f"{op}={c[op]}" for op in KEY if c[op])
        print(f"{name}\n    instructions={sum(c.values())}  {sel}")
    print("\n# whole library")
    print("    " + "  ".join(f"{op}={total[op]}" for op in KEY if total[op]))
    print("    library GEMM / legacy tensor opcodes (HMMA, IMMA, HGMMA): " + str(total["HMMA"] + total["IMMA"] + total["HGMMA"]))


if __name__ == "__main__":
    sys.exit(main())
