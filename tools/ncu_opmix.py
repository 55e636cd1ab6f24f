"""Executed-instruction mix of a kernel from an ncu report's source page (read on the CPU box).

    python tools/ncu_opmix.py prof.ncu-rep [launch-index] [traj_steps_of_the_launch]

Prints warp-level executed instructions per opcode, the FP64 share, the issue-port cost model of DESIGN.md 5.1
(an FP64 instruction holds the SMSP issue port 2 cycles, everything else 1) and the stall samples by reason.
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
steps = float(sys.argv[3]) if len(sys.argv) > 3 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(txt)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": [], "hdr": None}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
b = blocks[which]
h = b["hdr"]
iS, iE, iT = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
ops = collections.Counter()
thr = collections.Counter()
for r in b["rows"]:
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
    if not m:
        continue
    op = m.group(1)
    ops[op] += int(r[iE])
    thr[op] += int(r[iT])
tot = sum(ops.values())
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
f64 = sum(ops[o] for o in FP64)
print(b["name"], f"rows {len(b['rows'])}")
print(f"warp-instructions executed {tot}; FP64 {f64} ({100*f64/tot:.1f} %); other {tot-f64}; issue-port cycles (2*FP64 + other) {2*f64+tot-f64}")
if steps:
    tt = sum(thr.values())
    flop = 2 * thr["DFMA"] + thr["DMUL"] + thr["DADD"]
    print(f"per trajectory-step: thread-instr {tt/steps:.0f}, FP64 {sum(thr[o] for o in FP64)/steps:.0f} (DFMA {thr['DFMA']/steps:.0f} DMUL {thr['DMUL']/steps:.0f} "
          f"DADD {thr['DADD']/steps:.0f} DSETP {thr['DSETP']/steps:.0f}), other {(tt-sum(thr[o] for o in FP64))/steps:.0f}, flop {flop/steps:.0f}")
print(" ".join(f"{o}:{100*c/tot:.2f}%" for o, c in ops.most_common(40)))
st = collections.Counter()
for i, name in enumerate(h):
    if name.startswith("stall_") and "(Not Issued)" not in name:
        st[name] = sum(int(r[i] or 0) for r in b["rows"])
tots = sum(st.values())
print("stall samples: " + " ".join(f"{k[6:]}:{100*v/tots:.1f}%" for k, v in st.most_common(10)))
