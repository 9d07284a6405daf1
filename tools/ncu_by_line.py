#!/usr/bin/env python
"""Join an ncu source-page CSV (per SASS instruction: executed count, stall samples) with nvdisasm -g line info and
aggregate per source line.  usage: ncu_by_line.py <ncu_source.csv> <nvdisasm_-g.sass> <mangled kernel name> [top]"""
import csv, re, sys
from collections import defaultdict
src_csv, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ins = [(int(r[ia], 16), int(r[ie] or 0), int(r[isamp] or 0), r[1]) for r in rows[2:] if r and r[0].startswith("0x")]
base = ins[0][0]
by_off = {a - base: (e, s, t) for a, e, s, t in ins}
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kern + ":"))
cur = None
agg = defaultdict(lambda: [0, 0, 0])
tot_e = tot_s = 0
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/", l)
    if m:
        off = int(m.group(1), 16)
        if off in by_off:
            e, s, _ = by_off[off]
            a = agg[cur]; a[0] += e; a[1] += s; a[2] += 1
            tot_e += e; tot_s += s
print(f"total executed warp-instructions {tot_e}, samples {tot_s}")
for k, (e, s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:<5d} exec {e:>11d} ({100.0*e/tot_e:5.1f}%)  samples {s:>6d} ({100.0*s/max(1,tot_s):5.1f}%)  sass {n}")
