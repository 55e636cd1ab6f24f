"""Per-kernel SASS opcode histogram of a built library (no GPU needed).

    python tools/sass_stats.py [lib.so] [kernel-name-substring]

Prints, per kernel: total instructions, FP64 pipe instructions (DFMA/DMUL/DADD/DSETP/...),
their share, and the top opcodes -- the static picture to look at before spending GPU time.
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "reentry_old_b200/libreentry_b200.so"
pat = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None
hist = collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
for k, h in hist.items():
    if pat not in k:
        continue
    tot = sum(h.values())
    fp64 = sum(v for o, v in h.items() if o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX", "F2F", "I2F", "F2I") and o[0] == "D")
    print(f"== {k}\n   instr {tot}  fp64 {fp64} ({100*fp64/max(tot,1):.1f}%)  DFMA {h['DFMA']} DMUL {h['DMUL']} DADD {h['DADD']} DSETP {h['DSETP']} MUFU {h['MUFU']}")
    print("   " + " ".join(f"{o}:{v}" for o, v in h.most_common(24)))
