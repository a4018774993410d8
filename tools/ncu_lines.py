"""Map ncu per-SASS-instruction counts to source lines (ncu's CSV source page has no line
correlation): join `ncu --page source --csv` with `nvdisasm -g` line info of the same cubin.

usage: python tools/ncu_lines.py <report.ncu-rep> <object.o> <kernel-substring> [top]"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
ncu_filter = sys.argv[5] if len(sys.argv) > 5 else kname     # demangled-name regex for ncu (kname matches the MANGLED name)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
func = cur = None
line_of = []
for l in txt:
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        func = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and func and kname in func:
        line_of.append((int(m.group(1), 16), cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + ncu_filter], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hdr_i]
ie, isamp = h.index("Instructions Executed"), h.index("# Samples")
data = [r for r in rows[hdr_i + 1:] if len(r) > ie and r[0].startswith("0x")]
base = int(data[0][0], 16)
per_line, samp, ops = collections.Counter(), collections.Counter(), collections.Counter()
tot = tots = 0
seen = set()
for r in data:
    off = int(r[0], 16) - base
    if off in seen:
        break
    seen.add(off)
    n, s = int(r[ie]), int(r[isamp])
    tot += n
    tots += s
    idx = off // 16
    if idx < len(line_of):
        per_line[line_of[idx][1]] += n
        samp[line_of[idx][1]] += s
    ops[r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]] += n
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
stall_tot = collections.Counter()
for r in data[:len(seen)]:
    for c in stalls:
        stall_tot[c] += int(r[h.index(c)] or 0)
print("stall samples:", ", ".join(f"{k}:{v}" for k, v in stall_tot.most_common(8)))
print(f"{len(line_of)} SASS instructions, {tot} executed warp-instructions, {tots} samples")
for k, v in per_line.most_common(top):
    print(f"{str(k):32s} {v:10d} {100 * v / tot:5.1f}%  samples {100 * samp[k] / max(tots, 1):5.1f}%")
print("top opcodes:", ", ".join(f"{k}:{100 * v / tot:.1f}%" for k, v in ops.most_common(18)))
