"""Opcode histogram of every kernel in libheston_b200.so (static SASS from `cuobjdump -sass`, no GPU needed).

    python profiles/sass_histogram.py [path/to/lib.so] > profiles/rNN_sass_histogram.txt

Evidence kept under profiles/: the cubins are sm_100a only, the unfused transform kernel stages its slices with
the bulk-async (TMA) engine (UBLKCP), and what the FP64 / integer / memory instruction mix of each kernel is.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pde_b200", "csrc", "libheston_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
print(f"# {os.path.relpath(lib, ROOT)}: cubin architectures {arch}")
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
FP64 = {"DFMA", "DADD", "DMUL", "DSETP", "MUFU"}
MEM = {"LDS", "STS", "LDG", "STG", "LDL", "STL", "LDC", "LDCU", "ATOMS", "ATOMG", "RED", "UBLKCP", "SYNCS"}
for k, h in hist.items():
    tot = sum(h.values())
    if tot == 0:
        continue
    f = sum(v for o, v in h.items() if o in FP64)
    mm = sum(v for o, v in h.items() if o in MEM)
    print(f"\n## {k}: {tot} SASS instructions, FP64 {100 * f / tot:.1f} %, memory {100 * mm / tot:.1f} %"
          + (", UBLKCP (bulk-async / TMA copy) present" if h.get("UBLKCP") else ""))
    print("   " + "  ".join(f"{o} {v}" for o, v in h.most_common(18)))
