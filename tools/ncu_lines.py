#!/usr/bin/env python
"""Attribute the executed warp instructions of an `ncu --page source --csv` (SASS view) capture to CUDA source
lines: the SASS rows of the capture are matched in order with `nvdisasm -g` of the same function in the cubin.
usage: ncu_lines.py X.source.csv <cubin> <mangled function name> [top N]"""
import csv, re, subprocess, sys, collections
src_csv, cubin, fun = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
iS, iE, iSmp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
sass = [(r[iS].strip(), int(r[iE]), int(r[iSmp])) for r in rows[2:] if len(r) > iE and r[iE].isdigit()]
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(fun + ":"))
end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith("//-----") or re.match(r"\s*\.section", dis[i])), len(dis))
line, ins = None, []
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((line, m.group(2).strip()))
if len(ins) != len(sass):
    print(f"warning: {len(ins)} disassembled vs {len(sass)} profiled instructions", file=sys.stderr)
per = collections.Counter(); smp = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for (ln, text), (s, n, sm) in zip(ins, sass):
    per[ln] += n; smp[ln] += sm; ops[ln][text.split()[1] if text.startswith("@") else text.split()[0]] += n
tot, ts = sum(per.values()), sum(smp.values())
px = float(sys.argv[5]) if len(sys.argv) > 5 else None
print(f"total warp instructions {tot}")
for ln, n in per.most_common(top):
    o = " ".join(f"{k}:{v * 100 // max(n, 1)}%" for k, v in ops[ln].most_common(4))
    print(f"{str(ln):28s} {n:11d} {100 * n / tot:5.1f}%  stalls {100 * smp[ln] / max(ts, 1):5.1f}%" + (f" {n * 32 / px:6.2f}/px" if px else "") + f"  {o}")
