"""Per-kernel SASS summary of libhrm_b200.so (cuobjdump -sass): instruction count and the mnemonics that show how a
kernel uses the machine (bulk-async copies + mbarrier, 128/256-bit loads, shared-memory atomics, DPX min/max, warp
reductions).  Output: profiles/r2_sass_summary.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "hashreadmapper_b200", "libhrm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UBLKCP", "SYNCS", "LDG.E.128", ".256", "LDG.E.64", "STG.E.128", "ATOMS.OR", "ATOMS.CAS", "ATOMS.ADD", "ATOMS.POPC",
        "ATOMG", "RED.", "VIMNMX3", "VIADDMNMX", "VIMNMX", "REDUX", "VOTE", "SHFL", "POPC", "MATCH", "LDS.128", "STS.128",
        "BAR.SYNC", "WARPSYNC", "IMAD.WIDE", "LOP3", "SHF"]
kern = None
stats = collections.OrderedDict()
arch = ""
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        stats[kern] = collections.Counter()
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch = m.group(1)
    if kern and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        body = line.split("*/", 1)[1]
        stats[kern]["#instr"] += 1
        for k in KEYS:
            if k in body:
                stats[kern][k] += 1
dem = subprocess.run(["c++filt"] + list(stats.keys()), capture_output=True, text=True).stdout.splitlines()
lines = ["SASS summary of hashreadmapper_b200/libhrm_b200.so (%s), `cuobjdump -sass`, per kernel: instruction count and "
         "selected mnemonics" % arch, ""]
for (k, c), d in zip(stats.items(), dem):
    name = re.sub(r"\(.*", "", d)
    if "hrm::" not in name:
        continue
    sel = ", ".join("%s x%d" % (kk, c[kk]) for kk in KEYS if c[kk])
    lines.append("%-70s %6d instr  %s" % (name[:70], c["#instr"], sel))
path = os.path.join(ROOT, "profiles", "r2_sass_summary.txt")
open(path, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:12]))
print("...", len(lines), "lines ->", path)
