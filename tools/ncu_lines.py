"""Aggregate an ncu source-page CSV (`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`)
per CUDA source line: warp instructions executed, stall samples.  Usage: ncu_lines.py file.csv [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
inst = defaultdict(float); samp = defaultdict(float); src = {}
cur = None; hdr = None; fname = ''
for r in rows:
    if len(r) >= 2 and r[0] in ('File Name', 'File Path'):
        fname = r[1].split('/')[-1]; continue
    if len(r) > 4 and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < 8:
        continue
    if r[0] != '':
        cur = (fname, int(r[0])); src[cur] = r[1].strip()[:90]; continue
    if r[2] in ('', '...') or cur is None:
        continue
    try:
        ie = float(r[hdr.index('Instructions Executed')]); sa = float(r[hdr.index('# Samples')])
    except ValueError:
        continue
    inst[cur] += ie; samp[cur] += sa
ti, ts = sum(inst.values()), sum(samp.values())
print('total warp-instr %.3e  samples %d' % (ti, ts))
for k in sorted(inst, key=lambda k: -samp[k])[:top]:
    print('%5.1f%% smp %5.1f%% inst  %s:%d  %s' % (100 * samp[k] / max(ts, 1), 100 * inst[k] / max(ti, 1), k[0], k[1], src[k]))
