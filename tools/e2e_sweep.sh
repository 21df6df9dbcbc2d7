# developer tool: end-to-end throughput against sub-chunk size, pipelining and handle count (run on the GPU box)
for cfg in "256 2 0" "256 3 0" "256 4 0" "128 3 0" "128 4 0" "512 3 0"; do
  set -- $cfg
  CUBOID_SUB_BATCH=$1 CUBOID_E2E_PIPELINE=$3 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu --e2e-handles $2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sub $1 handles $2 pipeline $3', round(d['value']), round(d['e2e']['value']), round(d['e2e']['serial_calls_value']))"
done
