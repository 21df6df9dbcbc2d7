#!/bin/bash
# developer tool: full GPU suite, then the default bench line (summary to gpurun_out/suite.log)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --no-cpu > gpurun_out/suite_bench.json 2> gpurun_out/suite_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/suite_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['stages_ms_per_step'], d['e2e']['value'], d['e2e'].get('serial_calls_value'))
c = d.get('configs', {})
for k, v in c.items():
    if k == 'single_frame':
        print(k, {kk: (vv.get('p50_ms') if isinstance(vv, dict) else None) for kk, vv in v.items()})
    else:
        print(k, v['value'], v['e2e']['value'])
PY
