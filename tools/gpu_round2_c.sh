set -x
mkdir -p gpurun_out
CUBOID_CUDA_LIB=$PWD/tools/dev/libcuboid_cuda_stats.so timeout 300 python tools/icp_stats.py 256 > gpurun_out/r2c_stats.log 2>&1
cat gpurun_out/r2c_stats.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
tail -15 gpurun_out/r2c_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-configs --e2e-handles 1 > gpurun_out/r2c_bench_t1.json 2> gpurun_out/r2c_bench_t1.err
CUBOID_ICP_TABLE=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-configs --e2e-handles 1 > gpurun_out/r2c_bench_t0.json 2> gpurun_out/r2c_bench_t0.err
python - <<'PY'
import json
for n in ("t1","t0"):
    try:
        d=json.loads(open("gpurun_out/r2c_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["value"]), d["stages_ms_per_step"], round(d["e2e"]["value"]))
    except Exception as e: print(n, "failed", e)
PY
tail -3 gpurun_out/r2c_bench_t1.err
