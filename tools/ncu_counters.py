"""Regenerate profiles/dram_traffic.json and profiles/kernel_counters.json from the round's ncu --set full captures:
    python tools/ncu_counters.py gpurun_out/<dir>/prof_sca.ncu-rep gpurun_out/<dir>/prof_lp.ncu-rep
(bench.py copies the dominant stage's entries into roofline.traffic / roofline.counters and says they come from profiles/)."""
import csv
import io
import json
import subprocess
import sys


def load(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = rows[0]
    return [dict(zip(h, r)) for r in rows[2:]]


def pick(rows, sub):
    c = [r for r in rows if sub in r['Kernel Name']]
    if not c:
        raise SystemExit('no kernel matching %r' % sub)
    return c[-1]


def f(r, k):
    return float(r[k])


STALL = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio'


def counters(r, name, source):
    stalls = {s: round(f(r, STALL % s), 2) for s in ('wait', 'long_scoreboard', 'no_instruction', 'branch_resolving', 'short_scoreboard',
                                                       'math_pipe_throttle', 'not_selected', 'lg_throttle') if (STALL % s) in r}
    top = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
    return {'kernel': name, 'duration_ms': round(f(r, 'gpu__time_duration.sum'), 4),
            'registers': int(f(r, 'launch__registers_per_thread')),
            'issue_active_pct': round(f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'), 1),
            'warps_active_pct': round(f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'), 1),
            'threads_per_warp_instruction': round(f(r, 'smsp__thread_inst_executed_per_inst_executed.ratio'), 1),
            'pipe_fma_pct': round(f(r, 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'), 1),
            'pipe_fp64_pct': round(f(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'), 1),
            'pipe_alu_pct': round(f(r, 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'), 1),
            'pipe_lsu_pct': round(f(r, 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'), 1),
            'warp_instructions': int(f(r, 'smsp__inst_executed.sum')),
            'l1_hit_pct': round(f(r, 'l1tex__t_sector_hit_rate.pct'), 1), 'l2_hit_pct': round(f(r, 'lts__t_sector_hit_rate.pct'), 1),
            'top_stalls_per_issue': top, 'source': source}


def dram(r):
    unit = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    return f(r, 'dram__bytes_read.sum') * unit[UNITS['dram__bytes_read.sum']] + f(r, 'dram__bytes_write.sum') * unit[UNITS['dram__bytes_write.sum']]


def units_of(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return dict(zip(rows[0], rows[1]))


sca_rep, lp_rep = sys.argv[1], sys.argv[2]
sca, lp = load(sca_rep), load(lp_rep)
UNITS = units_of(sca_rep)
s1, s2, integ, knn, stat = (pick(sca, k) for k in ('k_select_fast<0, 0>', 'k_select_fast<0, 1>', 'k_integrate', 'k_neighbors_direct<0>', 'k_status'))
traffic = {'_comment': 'dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, BASELINE configs[4] (10^6 agents), round-2 '
                       'final build (tools/ncu_counters.py); bench.py reports the dominant stage\'s figure as roofline.traffic (from '
                       'profiles/, not measured in the run)',
           'policy': dram(s1), 'policy_detail': {'k_select_fast<SCA,0>': dram(s1), 'k_select_fast<SCA,1>': dram(s2), 'k_integrate': dram(integ)},
           'neighbors': dram(knn), 'status': dram(stat)}
UNITS = units_of(lp_rep)
l1, l2 = pick(lp, 'k_policy_lp<5>'), pick(lp, 'k_policy_lp<4>')
traffic['policy_orca_lp'] = dram(l1) + dram(l2)
traffic['policy_orca_lp_detail'] = {'k_policy_lp<5> (register-only pass)': dram(l1), 'k_policy_lp<4> (list pass)': dram(l2)}
json.dump(traffic, open('profiles/dram_traffic.json', 'w'), indent=1)
c1, c2 = counters(l1, 'k_policy_lp<5> (register-only pass)', 'profiles/r2_ncu_k_policy_lp_pass1.txt'), counters(l2, 'k_policy_lp<4> (list pass)', 'profiles/r2_ncu_k_policy_lp_pass2.txt')
kc = {'_comment': 'ncu --set full counters of the dominant kernel of each stage (BASELINE configs[4], 10^6 agents, round-2 final build, '
                  'tools/ncu_counters.py); bench.py copies the entry of the dominant stage into roofline.counters (from profiles/, not '
                  'measured in the run)',
      'policy': counters(s1, 'k_select_fast<SCA, pass 1>', 'profiles/r2_ncu_k_select_fast.txt'),
      'neighbors': counters(knn, 'k_neighbors_direct<pass 1>', 'profiles/r2_ncu_k_neighbors_direct.txt'),
      'status': counters(stat, 'k_status', 'profiles/r2_ncu_k_status.txt'),
      'policy_orca_lp': {'kernel': c1['kernel'] + ' + ' + c2['kernel'],
                         **{k: [c1[k], c2[k]] for k in ('duration_ms', 'issue_active_pct', 'warps_active_pct', 'threads_per_warp_instruction', 'registers')},
                         'source': c1['source'] + ', ' + c2['source']}}
json.dump(kc, open('profiles/kernel_counters.json', 'w'), indent=1)
print(json.dumps(traffic, indent=1))
print(json.dumps(kc, indent=1))
