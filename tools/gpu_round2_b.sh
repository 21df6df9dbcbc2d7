set -x
mkdir -p gpurun_out
timeout 300 python tools/icp_stats.py 256 > gpurun_out/r2b_stats.log 2>&1
cat gpurun_out/r2b_stats.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'^k_icp$' --launch-skip 1 --launch-count 1 -o gpurun_out/prof_icp_r2_b -f python bench.py --steps 1 --warmup 1 --no-cpu --no-configs --e2e-handles 1 > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log
