# developer tool (GPU box): ncu --set full of k_icp for ONE frame (the single-frame latency path), digested on the box
set -x
mkdir -p gpurun_out
TAG=${1:-single}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'^k_icp$' --launch-skip 3 --launch-count 1 -o /tmp/prof_single -f python tools/single_frame.py 3 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
ncu -i /tmp/prof_single.ncu-rep --page source --print-source cuda,sass --csv > /tmp/src_single.csv 2>/dev/null
python tools/ncu_lines.py /tmp/src_single.csv 60 > gpurun_out/${TAG}_k_icp_lines.txt
python tools/ncu_summary.py /tmp/prof_single.ncu-rep --json gpurun_out/${TAG}_k_icp_summary.json > gpurun_out/${TAG}_k_icp_summary.md
