#!/usr/bin/env python
"""Instruction mix of a kernel from `ncu --page source --csv` (SASS view): executed warp instructions per opcode,
with their share and the stall samples attributed to them.  usage: ncu_opmix.py X.source.csv [pixels]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iE, iSmp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iX = hdr.index("L1 Wavefronts Shared Excessive") if "L1 Wavefronts Shared Excessive" in hdr else None
mix, smp, exc = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= iE: continue
    toks = r[iS].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.split(".")[0] in ("LDS", "STS", "LDG", "STG", "IMAD", "I2F", "F2I", "I2FP", "F2FP") and "." in op else "")
    try:
        n = int(r[iE]); s = int(r[iSmp])
    except ValueError:
        continue
    mix[op] += n; smp[op] += s
    if iX is not None:
        try: exc[op] += int(r[iX])
        except ValueError: pass
tot = sum(mix.values()); ts = sum(smp.values())
px = float(sys.argv[2]) if len(sys.argv) > 2 else None
print(f"total warp instructions {tot}" + (f" = {tot * 32 / px:.1f} thread instructions per pixel" if px else ""))
for op, n in mix.most_common(28):
    print(f"{op:14s} {n:12d} {100 * n / tot:5.1f}%  stall samples {100 * smp[op] / max(ts, 1):5.1f}%" + (f"  smem excess wavefronts {exc[op]}" if exc[op] else "") + (f"  {n * 32 / px:6.2f}/px" if px else ""))
