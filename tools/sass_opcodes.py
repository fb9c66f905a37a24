#!/usr/bin/env python
"""Opcode histogram per kernel of libsmm_b200.so (cuobjdump -sass), for profiles/: which Blackwell / Hopper+ mechanisms the
built library really contains.    python tools/sass_opcodes.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sparse_matrix_math_b200", "libsmm_b200.so")
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_opcodes.txt")
NOTE = {
    "UBLKCP": "TMA bulk copy (cp.async.bulk)", "SYNCS": "mbarrier (arrive / try_wait)", "UCGABAR_ARV": "cluster barrier arrive", "UCGABAR_WAIT": "cluster barrier wait",
    "LDGSTS": "cp.async (global -> shared)", "LDGDEPBAR": "cp.async group commit", "DEPBAR": "cp.async wait / scoreboard wait", "NANOSLEEP": "nanosleep back-off",
    "MEMBAR": "memory fence", "CCTL": "cache control", "ATOMG": "global atomic", "REDG": "global reduction", "ATOMS": "shared atomic", "MUFU": "special function (rcp of the IEEE division)",
    "SHFL": "warp shuffle", "VOTE": "warp vote", "VOTEU": "warp vote (uniform)", "REDUX": "warp reduce (redux.sync)", "MATCH": "warp match", "WARPSYNC": "__syncwarp", "BAR": "CTA barrier",
    "ERRBAR": "error barrier (fence of the cluster sync)", "CGAERRBAR": "cluster error barrier", "S2UR": "special register (cluster rank ...)", "LDS": "shared load", "STS": "shared store",
    "LDG": "global load", "STG": "global store", "ST": "generic store (st.shared::cluster lands here)", "LD": "generic load", "HMMA": "tensor core", "UTCMMA": "tcgen05 mma",
}
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        kern = re.sub(r"\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
        if m.group(1) in ("SYNCS", "UBLKCP", "LDG", "ST", "LD"):
            hist[kern][m.group(1) + m.group(2)] += 1
with open(OUT, "w") as f:
    f.write(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): opcode counts per kernel; * = mechanism worth noting\n")
    f.write("# tensor-core opcodes (HMMA / UTCMMA ...) are absent by design: nothing on this path is a dense contraction\n\n")
    for k, c in hist.items():
        total = sum(v for op, v in c.items() if "." not in op)
        f.write(f"{k}   [{total} instructions]\n")
        notable = ("UBLKCP", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "LDGSTS", "LDGDEPBAR", "NANOSLEEP", "REDUX", "ATOMS", "HMMA", "UTCMMA")
        special = [(op, v) for op, v in c.items() if op.split(".")[0] in notable]
        for op, v in sorted(special, key=lambda t: -t[1]):
            f.write(f"    * {op:38s} {v:5d}   {NOTE[op.split('.')[0]]}\n")
        common = ", ".join(f"{op} {v}" for op, v in c.most_common(14) if "." not in op)
        f.write(f"      most frequent: {common}\n\n")
print(open(OUT).read()[:3000])
