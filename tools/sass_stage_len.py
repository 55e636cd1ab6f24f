"""Static length of the step kernel's stages: instructions between consecutive K-row store triples (STS.64 x3) of
the inlined step kernel -- a GPU-free proxy for 'did this change remove instructions from a stage'.
    python tools/sass_stage_len.py [lib.so]"""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "reentry_old_b200/libreentry_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
body, on = [], False
for line in txt.splitlines():
    if "Function : " in line:
        on = "k_propagateILi128ELi6ELb0" in line
    elif on:
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);", line)
        if m:
            body.append((int(m.group(1), 16), m.group(2), m.group(3)))
sts = [i for i, (_, op, arg) in enumerate(body) if op.startswith("STS") ]
# group consecutive STS (within 12 instructions) into triples
groups, cur = [], []
for i in sts:
    if cur and i - cur[-1] > 12:
        groups.append(cur); cur = []
    cur.append(i)
if cur:
    groups.append(cur)
trip = [g for g in groups if len(g) == 3]
print("instructions", len(body), "K-store triples at", [hex(body[g[0]][0]) for g in trip])
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
for a, b in zip(trip[:-1], trip[1:]):
    seg = body[a[-1] + 1: b[-1] + 1]
    f = sum(1 for _, op, _ in seg if op.split(".")[0] in FP64)
    print(f"  {hex(body[a[-1]][0])} -> {hex(body[b[-1]][0])}: {len(seg)} instr, FP64 {f}, other {len(seg) - f}, issue slots {len(seg) + f}")
