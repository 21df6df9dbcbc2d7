# developer tool (GPU box): tests, bench with the queued ICP search on / off, ncu of k_icp
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --e2e-handles 1 > gpurun_out/r2a_bench_q1.json 2> gpurun_out/r2a_bench_q1.err
CUBOID_ICP_QUEUED=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --e2e-handles 1 > gpurun_out/r2a_bench_q0.json 2> gpurun_out/r2a_bench_q0.err
python - <<'PY'
import json
for n in ("q1","q0"):
    try:
        d=json.loads(open("gpurun_out/r2a_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["value"]), d["stages_ms_per_step"], round(d["e2e"]["value"]))
    except Exception as e: print(n, "failed", e)
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'^k_icp$' --launch-skip 2 --launch-count 1 -o gpurun_out/prof_icp_r2_a -f python bench.py --steps 1 --warmup 1 --no-cpu --e2e-handles 1 > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_ncu.log
