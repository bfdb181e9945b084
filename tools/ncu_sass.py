#!/usr/bin/env python
"""prints the SASS of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME --launch-count 1` with the
executed warp-instruction count and stall samples per instruction.   usage: python tools/ncu_sass.py src.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > ix['Instructions Executed'] and r[ix['Instructions Executed']].isdigit()]
tot = sum(int(r[ix['Instructions Executed']]) for r in data)
samp = sum(int(r[ix['# Samples']]) for r in data)
print("total warp instr", tot, "samples", samp, "n sass", len(data))
for k, r in enumerate(data):
    ie = int(r[ix['Instructions Executed']])
    s = int(r[ix['# Samples']])
    print("%4d %9d %5.2f%% %5d  %s" % (k, ie, 100.0 * ie / tot, s, r[ix['Source']].strip()[:100]))
