"""Join an ncu source-page export (SASS, per-instruction executed counts) with nvdisasm line info of the same
build and aggregate the issue-port cost per source line / per function.

    cuobjdump -xelf all libreentry_b200.so; nvdisasm --print-line-info rb_api.sm_100a.cubin > lines.sass
    ncu -i prof.ncu-rep --page source --csv --print-source sass > src.csv
    python tools/sass_profile.py src.csv lines.sass [kernel-symbol-substring]

Cost model (tools/ubench/fp64_mix.cu): an FP64 instruction holds the SMSP issue port 2 cycles (3 with three
distinct register operands), everything else 1.
"""
import collections
import csv
import glob
import os
import re
import sys

src_csv, lines_sass = sys.argv[1], sys.argv[2]
sym = sys.argv[3] if len(sys.argv) > 3 else "k_propagateILi128ELi6ELb0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

L = open(lines_sass).read().split("\n")
start = [i for i, l in enumerate(L) if l.startswith(".text.") and sym in l][0]
off2line = {}
cur = None
for l in L[start + 1:]:
    if l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', l)
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = cur

rows = list(csv.reader(open(src_csv)))
blocks = []
for r in rows:
    if r and r[0] == "Kernel Name":
        blocks.append([r[1], None, []])
    elif r and r[0] == "Address":
        blocks[-1][1] = r
    elif blocks and r:
        blocks[-1][2].append(r)
name, h, d = blocks[0]
iS, iE = h.index("Source"), h.index("Instructions Executed")
base = int(d[0][0], 16)
cost = collections.Counter(); n_fp = collections.Counter(); n_ot = collections.Counter()
ops = collections.Counter()
for r in d:
    ln = off2line.get(int(r[0], 16) - base)
    s = r[iS].strip()
    n = int(r[iE])
    m = re.search(r"\b(DFMA|DMUL|DADD|DSETP)\b[.\w]*\s+(.*?)$", s)
    if m:
        args = [a.strip() for a in m.group(2).split(",")][1:]
        regs = {re.search(r"\bR(\d+)\b", a).group(1) for a in args if re.search(r"\bR(\d+)\b", a) and not a.startswith("c[")}
        c = 3 if (m.group(1) == "DFMA" and len(regs) == 3) else 2
        n_fp[ln] += n
    else:
        c = 1
        n_ot[ln] += n
        op = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", s).group(2)
        ops[op] += n
    cost[ln] += c * n
tot = sum(cost.values())
print(f"kernel {name[:60]}: issue-port cycles (model) {tot:.4g}; fp64 instr {sum(n_fp.values()):.4g}, other {sum(n_ot.values()):.4g}")
files = {}
def text(k):
    if not k:
        return ""
    p = glob.glob(os.path.join(ROOT, "reentry_old_b200", "csrc", k[0]))
    if not p:
        return ""
    if k[0] not in files:
        files[k[0]] = open(p[0]).read().split("\n")
    return files[k[0]][k[1] - 1].strip()[:100]
for k, v in cost.most_common(int(os.environ.get("TOP", "60"))):
    print(f"{str(k):26s} {100 * v / tot:5.2f}%  fp64 {n_fp[k]:>10d} other {n_ot[k]:>10d}  {text(k)}")
print("non-FP64 opcodes:", ", ".join(f"{o} {100 * n / tot:.1f}%" for o, n in ops.most_common(16)))
