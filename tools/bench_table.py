"""Print one line per bench JSON file: value, ms/step, stage times, diagnostics.  Usage: bench_table.py file.json ..."""
import json
import sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print('%-40s ERR %s' % (f, e))
        continue
    r = d.get('roofline', {})
    print('%-40s %.1fM  ms/step %.3f  stages %s  %s e2e %s' % (
        f.split('/')[-1], d['value'] / 1e6, d['ms_per_step'], r.get('stage_ms'),
        {k: v for k, v in r.get('last_step_diag', {}).items() if k in ('n_tier2', 'n_slow')}, (d.get('e2e') or {}).get('ms_per_step')))
