#!/usr/bin/env python
"""SASS instruction-class table of every kernel in libtetrad_b200.so (cuobjdump -sass), as evidence of which hardware
paths a kernel uses: UTC*MMA (tcgen05.mma), UTCBAR (tcgen05.commit), LDTM / STTM (tcgen05.ld / st), UBLKCP (cp.async.bulk,
the TMA unit), IMMA (legacy mma.sync), LDGSTS (cp.async), SYNCS (mbarrier), POPC, DFMA ...

    python profiles/tools/sass_table.py > profiles/r02_sass_table.md
"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
LIB = ROOT / "tetrad_b200" / "libtetrad_b200.so"
KEYS = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "UTMALDG", "IMMA", "LDGSTS", "SYNCS",
        "ATOMS", "POPC", "LOP3", "SHF", "DFMA", "DMUL", "LDS", "STS", "LDG", "SHFL", "VOTE", "BRA", "NANOSLEEP"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    kernels[cur][k] += 1
                    break
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("| kernel | instructions | " + " | ".join(KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for mangled, name in zip(kernels, names):
        if "cub::" in name or "pipe_kernel" in name or "mix_kernel" in name:
            continue
        short = re.sub(r"\(.*", "", name).replace("tq::", "")
        c = kernels[mangled]
        print(f"| `{short}` | {c['total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")


if __name__ == "__main__":
    main()
