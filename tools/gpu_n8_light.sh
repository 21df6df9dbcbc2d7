#!/bin/bash
# developer tool (8-GPU box): the headline bench line at N = 8 only (the full set of 8-GPU runs is tools/gpu_n8.sh)
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-configs 2> gpurun_out/n8_light.err | tail -1 > gpurun_out/n8_light.json
python - <<'PY'
import json
d = json.loads(open('gpurun_out/n8_light.json').read())
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('frac_of_h2d_ceiling'), d['stages_ms_per_step'])
PY
