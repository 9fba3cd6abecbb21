"""Compact text summary of one kernel capture (`ncu --set full --import-source on`):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top] [kernel-name-regex] > profiles/rNN_kernel.txt
Key raw metrics (duration, DRAM bytes, occupancy, issue/stall picture, divergence) + the hottest source lines."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 25
ksel = [a for a in sys.argv[2:] if not a.isdigit()]            # optional: substring of the kernel name (a report may hold several kernels)
ksub = ksel[0] if ksel else ''
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
cand = [r for r in rows[2:] if ksub in r[h.index('Kernel Name')]]
v = cand[-1]
WANT = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
print('# ncu summary of', rep.split('/')[-1])
for w in WANT:
    if w in h:
        i = h.index(w)
        print('%-85s %s %s' % (w, v[i], units[i]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
from collections import defaultdict
inst, samp, text = defaultdict(float), defaultdict(float), {}
cur = hdr = None
in_kernel = True
fname = ''
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'Function Name':
        in_kernel = ksub in r[1].replace('(int)', '').replace('(bool)', '')
        continue
    if len(r) >= 2 and r[0] in ('File Name', 'File Path'):
        fname = r[1].split('/')[-1]
        continue
    if not in_kernel:
        continue
    if len(r) > 4 and r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or len(r) < 8:
        continue
    if r[0] != '':
        cur = (fname, int(r[0]))
        text[cur] = r[1].strip()[:100]
        continue
    if r[2] in ('', '...') or cur is None:
        continue
    try:
        inst[cur] += float(r[hdr.index('Instructions Executed')])
        samp[cur] += float(r[hdr.index('# Samples')])
    except ValueError:
        pass
ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
print('\n# hottest source lines (stall samples %, warp instructions %)')
for k in sorted(inst, key=lambda k: -samp[k])[:top]:
    print('%5.1f%% %5.1f%%  %s:%d  %s' % (100 * samp[k] / ts, 100 * inst[k] / ti, k[0], k[1], text[k]))
