# developer tool (GPU box): ncu --set full of the 1024-frame k_frontend launch of the bench (second launch = the timed device-resident step),
# digested on the box (metric table + per-source-line table)
set -x
mkdir -p gpurun_out
TAG=${1:-x}
k=k_frontend
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"^${k}\$" --launch-skip 1 --launch-count 1 -o /tmp/prof_${k} -f python bench.py --steps 1 --warmup 1 --no-cpu --no-configs --e2e-handles 1 > gpurun_out/${TAG}_ncu_${k}.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_${k}.log
ncu -i /tmp/prof_${k}.ncu-rep --page source --print-source cuda,sass --csv > /tmp/src_${k}.csv 2>/dev/null
python tools/ncu_lines.py /tmp/src_${k}.csv 70 > gpurun_out/${TAG}_${k}_lines.txt
python tools/ncu_summary.py /tmp/prof_${k}.ncu-rep --json gpurun_out/${TAG}_${k}_summary.json > gpurun_out/${TAG}_${k}_summary.md
