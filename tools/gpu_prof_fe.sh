# developer tool (GPU box): ncu --set full of the 1024-frame k_frontend launch of the bench (second launch = the timed device-resident step)
set -x
mkdir -p gpurun_out
TAG=${1:-x}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'^k_frontend$' --launch-skip 1 --launch-count 1 -o gpurun_out/prof_fe_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu --no-configs --e2e-handles 1 > gpurun_out/ncu_fe_$TAG.log 2>&1
tail -2 gpurun_out/ncu_fe_$TAG.log
