#!/bin/bash
# developer tool: front-end variants side by side (device-resident arm only)
mkdir -p gpurun_out
run() {
  echo "$*"
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-configs 2>gpurun_out/fe_try.err | tail -1 | python -c "
import sys, json
t = sys.stdin.read()
try:
    d = json.loads(t)
    print(d['value'], d['ms_per_step'], d['stages_ms_per_step']['preprocess'], d['e2e']['value'])
except Exception as e:
    print('FAILED', t[:200]); print(open('gpurun_out/fe_try.err').read()[-1500:])
"
}
run CUBOID_FE_SOLO=0
run CUBOID_FE_SOLO=1
run CUBOID_FE_SOLO=0
run CUBOID_FE_SOLO=1
run CUBOID_FE_SOLO=1 CUBOID_FE_RUNS=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "one_pass_front_end or fused_frontend_equals or (knobs and FE_) or full_size or taps_off" 2>&1 | tail -5
