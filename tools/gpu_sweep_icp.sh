set -x
mkdir -p gpurun_out
for cfg in "4 8" "8 8" "8 16" "8 4"; do
  set -- $cfg
  CUBOID_ICP_NSUB=$1 CUBOID_ICP_SLICE=$2 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-configs --e2e-handles 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nsub $1 slice $2', round(d['value']), d['stages_ms_per_step']['icp'])"
done
