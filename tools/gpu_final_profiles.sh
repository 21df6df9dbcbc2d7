# developer tool (GPU box): the artefacts profiles/ is built from: default bench line, reference arm, launch list, ncu --set full of the
# four big kernels (digested on the box into a one-row metric csv + a per-source-line table: gpurun brings back at most 64 MiB)
set -x
mkdir -p gpurun_out
TAG=${1:-r2_final}
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null
for f in 1 64; do timeout 300 python bench.py --workload guess64 --shard hypotheses --frames $f --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/${TAG}_hyp_n1_f$f.json; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-configs > gpurun_out/${TAG}_ncu_launches.log 2>&1
for k in k_icp k_frontend k_sac_plane k_cluster; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:"^${k}\$" --launch-skip 1 --launch-count 1 -o /tmp/prof_${k} -f python bench.py --steps 1 --warmup 1 --no-cpu --no-configs --e2e-handles 1 > gpurun_out/${TAG}_ncu_${k}.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_${k}.log
  ncu -i /tmp/prof_${k}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${k}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_${k}.ncu-rep --page source --print-source cuda,sass --csv > /tmp/src_${k}.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/src_${k}.csv 40 > gpurun_out/${TAG}_${k}_lines.txt
  python tools/ncu_summary.py /tmp/prof_${k}.ncu-rep --json gpurun_out/${TAG}_${k}_summary.json > gpurun_out/${TAG}_${k}_summary.md
done
du -sh gpurun_out
