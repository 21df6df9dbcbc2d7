# developer tool (8-GPU box): the bench at N = 8 (frames sharded), BASELINE config 4 (multi8) at N = 8, the hypothesis-sharded mode, NCCL test
set -x
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu 2>gpurun_out/n${N}_full.err | tail -1 > gpurun_out/n${N}_full.json
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --e2e-handles 2 2>/dev/null | tail -1 > gpurun_out/n${N}_full_h2.json
timeout 600 $TR bench.py --gpus $N --workload multi8 --steps 3 --warmup 2 --no-cpu 2>gpurun_out/n${N}_multi8.err | tail -1 > gpurun_out/n${N}_multi8.json
timeout 600 $TR bench.py --gpus $N --workload hd720 --steps 3 --warmup 2 --no-cpu --e2e-handles 2 2>gpurun_out/n${N}_hd720.err | tail -1 > gpurun_out/n${N}_hd720.json
for f in 1 64; do timeout 300 $TR bench.py --gpus $N --workload guess64 --shard hypotheses --frames $f --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/n${N}_hyp_f$f.json; done
timeout 600 python -m pytest tests/test_hypothesis_sharding.py -m gpu -x -q 2>&1 | tail -3
python - <<PY
import json
for n in ("full","full_h2","multi8","hd720","hyp_f1","hyp_f64"):
    try:
        d=json.loads(open("gpurun_out/n${N}_%s.json"%n).read())
        e=d.get("e2e",{})
        print(n, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(e.get("value",0)), "ceiling", e.get("h2d_ceiling_gbs_per_gpu"), "frac", e.get("frac_of_h2d_ceiling"), d.get("latency_ms"), d.get("results_digest"))
    except Exception as ex: print(n, "failed", ex)
PY
