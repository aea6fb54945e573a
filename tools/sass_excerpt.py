"""SASS evidence for the built library (run here, no GPU): per kernel the registers / shared memory / spills (cuobjdump
-res-usage) and the counts of the mnemonics that show how it was built — UBLKCP (TMA bulk copy), SYNCS (mbarrier), STG.E.128 /
LDG.E.128 (128-bit global access), LDS.128, MUFU.* (special-function unit), FFMA, VOTE / SHFL / MATCH (warp ballot,
compaction), ATOM / RED, FMNMX3.  Writes profiles/r02_sass_excerpt.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "opentk-pathtracer_b200", "libptb200.so")
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\n\s+(REG:\d+.*)", res):
    usage[m.group(1)] = m.group(2).strip()
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, cur = collections.defaultdict(collections.Counter), None
pat = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = pat.match(line)
    if m and cur:
        op = m.group(1)
        counts[cur]["total"] += 1
        for key in ("UBLKCP", "SYNCS", "STG.E.128", "LDG.E.128", "LDS.128", "MUFU", "FFMA", "FMNMX3", "VOTE", "SHFL", "ATOM", "RED", "BREV", "FLO", "LDL", "STL"):
            if op.startswith(key) or (key in ("STG.E.128", "LDG.E.128") and op.startswith(key.split(".")[0]) and ".128" in op):
                counts[cur][key] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
out = ["SASS excerpt of opentk-pathtracer_b200/libptb200.so (sm_100a only; tools/sass_excerpt.py)", ""]
for mangled, name in sorted(zip(counts, demangle), key=lambda x: x[1]):
    c = counts[mangled]
    out.append(name)
    out.append("    " + usage.get(mangled, "?"))
    out.append("    " + "  ".join(f"{k}={c[k]}" for k in ("total", "UBLKCP", "SYNCS", "STG.E.128", "LDG.E.128", "LDS.128", "MUFU", "FFMA", "FMNMX3", "VOTE", "SHFL", "BREV", "FLO", "ATOM", "RED", "LDL", "STL") if c[k]))
open(os.path.join(ROOT, "profiles", "r02_sass_excerpt.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
