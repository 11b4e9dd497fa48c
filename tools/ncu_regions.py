#!/usr/bin/env python
"""Summarise `ncu --page source --csv` of one kernel into regions of consecutive SASS lines with the same
execution count: first line, #lines, executions and warp-instructions per unit of work, stall share.
    python tools/ncu_regions.py source.csv UNITS [min_instr_per_unit]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]); thr = float(sys.argv[3]) if len(sys.argv) > 3 else 5.0
hdr = rows[1]; body = rows[2:]
iS, iE, iN, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
tot_e = sum(float(r[iE] or 0) for r in body); tot_s = sum(float(r[iN] or 0) for r in body)
print(f"# SASS lines {len(body)}  warp-instr {tot_e:.0f}  per unit {tot_e/units:.1f}  samples {tot_s:.0f}")
print("first  lines  exec/unit  instr/unit  stall%  thr/instr  opcodes")
i = 0
while i < len(body):
    j = i; e = float(body[i][iE] or 0)
    while j + 1 < len(body) and abs(float(body[j + 1][iE] or 0) - e) <= 0.002 * max(e, 1): j += 1
    n = j - i + 1; ins = sum(float(r[iE] or 0) for r in body[i:j + 1]); smp = sum(float(r[iN] or 0) for r in body[i:j + 1])
    thrd = sum(float(r[iT] or 0) for r in body[i:j + 1])
    if ins / units >= thr:
        ops = " ".join(r[iS].split()[0] if not r[iS].startswith("@") else r[iS].split()[1] for r in body[i:min(j + 1, i + 14)])
        print(f"{i:5d} {n:6d} {e/units:10.2f} {ins/units:11.1f} {100*smp/tot_s:7.1f} {thrd/max(ins,1):9.1f}  {ops}")
    i = j + 1
