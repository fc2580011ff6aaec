#!/usr/bin/env python
"""Per-kernel SASS opcode evidence of the built library -> profiles/r2_sass_opcodes.txt (no GPU needed).
    python scripts/sass_opcodes.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "vsearch_b200", "lib", "libvsearch_b200.so")],
                     capture_output=True, text=True, check=True).stdout
out = ["# SASS opcode evidence per kernel of vsearch_b200/lib/libvsearch_b200.so (cuobjdump -sass, sm_100a); counts of the",
       "# Blackwell-specific mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG = TMA tensor",
       "# load, UBLKCP = 1-D bulk TMA copy, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops; HMMA would be the legacy mma.sync path.",
       "# regenerate: python scripts/sass_opcodes.py"]
keys = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "LDS", "ATOMS", "LDG", "STG",
        "SHFL", "MATCH", "REDUX", "LDL", "STL")
rows = []
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    if name.startswith("_ZN3cub"):
        continue   # CUB's scan / segmented sort (index build and export only)
    ops = collections.Counter(re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, flags=re.M))
    base = collections.Counter()
    for k, v in ops.items():
        base[k.split(".")[0]] += v
    rows.append((name, sum(ops.values()), {k: base[k] for k in keys if base[k]}))
for name, tot, i in sorted(rows):
    out.append(f"{name}\n    instructions={tot}  " + "  ".join(f"{k}={v}" for k, v in i.items()))
with open(os.path.join(ROOT, "profiles", "r2_sass_opcodes.txt"), "w") as fh:
    fh.write("\n".join(out) + "\n")
print("\n".join(out))
