"""Developer tool: writes profiles/README.md from the artefacts of one capture tag (tools/gpu_final_profiles.sh + tools/make_profiles.py).
usage: python tools/make_profiles_readme.py r2_h"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1]


def jl(name):
    return json.loads(open(os.path.join(P, name)).read().strip().splitlines()[-1])


d = jl(tag + "_bench.json")
ref = jl(tag + "_bench_reference.json")
issue = json.load(open(os.path.join(P, "icp_issue.json")))
traffic = json.load(open(os.path.join(P, "frontend_traffic.json")))
rows = [r for r in csv.reader(open(os.path.join(P, tag + "_launches.csv"))) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}


def short(n):
    return re.sub(r"\(.*", "", n).replace("cuboid::", "").replace("void ", "").replace("<unnamed>::", "")


seq = [(short(r[ki]), float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-3)) for r in rows[1:]]
setup = ("k_peak", "k_nn_table_build", "k_nn_seed_build")
work = [(n, us) for n, us in seq if not n.startswith(setup)]
first, nsel = [], 0
for n, us in work:
    first.append((n, us))
    if n.startswith("k_icp_select"):
        nsel += 1
        if nsel == 2:
            break


def table(items):
    agg = collections.OrderedDict()
    for n, us in items:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    out = "| kernel | launches | total us | us/launch | share |\n|---|---|---|---|---|\n"
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out += "| %s | %d | %.1f | %.1f | %.1f %% |\n" % (n, a[0], a[1], a[1] / a[0], 100 * a[1] / tot)
    return out


def md(k):
    return open(os.path.join(P, "%s_%s_summary.md" % (tag, k))).read()


def lines(k, n=16):
    return "\n".join(open(os.path.join(P, "%s_%s_lines.txt" % (tag, k))).read().split("\n")[:n + 3])


st = d["stages_ms_per_step"]
stot = sum(st.values())
c = d["configs"]
sf = c["single_frame"]
setup_us = {n: us for n, us in seq if n.startswith(("k_nn_table_build", "k_nn_seed_build"))}
out = f"""# profiles/ - round 2 (B200, sm_100a, CUDA 12.9, driver 580); round-1 files (`r1_*`) are kept for the history

All numbers below come from files in this directory; bench lines are plain runs (never under a profiler), the ncu launch list is
cold-cache and serialised, so only the SHARES are comparable with the CUDA-event stage times. Captures are made by
`tools/gpu_final_profiles.sh` (the `.ncu-rep` files are digested on the GPU box into a one-row metric csv, a metric table and a
per-source-line table: a `gpurun` call brings back at most 64 MiB), copied here by `tools/make_profiles.py`; this file is written by
`tools/make_profiles_readme.py {tag}`.

| file | what |
|---|---|
| `{tag}_bench.json` | end of round 2: `python bench.py` (default: 1 GPU, 1024 frames x 10 steps, every BASELINE config in `configs`) |
| `{tag}_bench_reference.json` | `python bench.py --impl reference` (CPU restatement of the PCL path, literal mode, all host cores) |
| `{tag}_launches.csv` | ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`) of `python bench.py --steps 1 --warmup 1 --no-cpu --no-configs` |
| `{tag}_k_icp_*`, `{tag}_k_frontend_*`, `{tag}_k_sac_plane_*`, `{tag}_k_cluster_*` | `ncu --set full --import-source on --clock-control none` of the 1024-frame launch of each of the four big kernels: `_summary.md` / `_summary.json` (key metrics), `_raw.csv` (every metric of the capture), `_lines.txt` (source lines by stall samples and by executed instructions, with lanes per instruction) |
| `icp_issue.json`, `frontend_traffic.json` | what `bench.py` reads for `roofline.issue` and `roofline_hbm.traffic` (derived from the two captures above) |
| `{tag}_sass_histogram.md` | `cuobjdump -sass` opcode histogram of every kernel of the built `.so` (`tools/sass_hist.py`): `UBLKCP` / `SYNCS` (TMA bulk copies + mbarriers) in `k_icp` and `k_sac_plane`, `LDGSTS` in `k_frontend` (and the cluster barriers `UCGABAR_*` in its general instance; none in the `SOLO` instance a throughput launch runs), no `UTMALDG`, no tensor-core opcodes (nothing on this path is a contraction) |
| `{tag}_hyp_n1_f1.json`, `{tag}_hyp_n1_f64.json`, `r2_c_hyp_n2_*`, `r2_g_hyp_n8_*` | `bench.py --workload guess64 --shard hypotheses` at 1 / 2 / 8 GPUs, one frame and 64 frames per step (identical `results_digest`) |
| `r2_i_bench_n8.json`, `r2_g_bench_n8_multi8.json`, `r2_f_bench_n8_hd720.json` | the bench on 8 GPUs of one box (torchrun; frames sharded, NCCL only for the timing all-reduce): headline workload, BASELINE config 4 (multi8), config 5 (hd720) |

## End-of-round bench line (`{tag}_bench.json`)

* value (device-resident input, CUDA events): **{d['value']:.0f} frames/s** ({d['ms_per_step']:.2f} ms per 1024-frame step); round 1: 55 398
* e2e (pinned host depth in, results out; three handles, one host thread each): **{d['e2e']['value']:.0f} frames/s** = {d['e2e']['frac_of_h2d_ceiling']:.3f} of the bare
  pinned H2D copy of the same bytes measured in the same run ({d['e2e']['h2d_ceiling_gbs_per_gpu']:.1f} GB/s = {d['e2e']['h2d_ceiling_gbs_per_gpu'] * 1e9 / 614400:.0f} frames/s); one handle, serial calls: {d['e2e']['serial_calls_value']:.0f}
* stage ms per step: {json.dumps({k: round(v, 3) for k, v in st.items()})} ("preprocess" = the fused front end)
* CPU restatement, literal mode, 1 thread: {d['cpu_baseline']['value']:.1f} frames/s (`gpu_matches_oracle_on_sample`: {d['cpu_baseline']['gpu_matches_oracle_on_sample']}); reference arm on {ref['cpu_baseline']['cores']} host threads: {ref['value']:.1f} frames/s
* clocks during the timed region: {json.dumps(d['clocks'])}
* `roofline` (`k_icp`, FP32 un-fused): executed {d['roofline']['achieved']:.2f} of {d['roofline']['peak']:.1f} TFLOP/s = {d['roofline']['frac']:.3f}: the kernel is not FP32 bound; its `issue` object
  (from `icp_issue.json`): {issue['issue_active_pct']:.1f} % of the issue slots, {issue['lanes_per_instruction']:.1f} of 32 lanes per instruction, {issue['warp_instructions_per_launch'] / 1e9:.2f} G warp instructions per launch
  (round 1: 73 %, 14.7 lanes, 8.9 G)
* `roofline_hbm` (`k_frontend`): algorithmic {d['roofline_hbm']['algorithmic_bytes_per_step'] / 1e9:.2f} GB per launch -> {d['roofline_hbm']['achieved']:.0f} GB/s = {d['roofline_hbm']['frac']:.3f} of the measured 6537 GB/s; DRAM
  traffic of the same launch under ncu: {traffic['dram_bytes_per_launch'] / 1e9:.2f} GB
* configs (same line): single frame `cuboid_process_cloud` p50 {sf['process_cloud']['p50_ms']:.2f} ms / p99 {sf['process_cloud']['p99_ms']:.2f} ms from pageable memory{(', %.2f ms from pinned memory' % sf['process_cloud_pinned']['p50_ms']) if 'process_cloud_pinned' in sf else ''}, `cuboid_process_batch` of one depth
  frame p50 {sf['process_batch_1']['p50_ms']:.2f} ms (round 1: ~4.9 ms of kernels); seg {c['seg']['value']:.0f}; guess64 {c['guess64']['value']:.0f} (round 1: 1 060); multi8 {c['multi8']['value']:.0f} (5 771);
  hd720 {c['hd720']['value']:.0f} (1 331) frames/s

## ncu launch list (`{tag}_launches.csv`)

The two whole-chunk passes of the device-resident arm (warm-up + timed step), i.e. the launches the CUDA events time (the one-off
launches of handle set-up are left out: {', '.join('`%s` %.1f ms' % (n, us / 1e3) for n, us in setup_us.items())} in `cuboid_set_template`, and the FP32 peak
micro-benchmarks):

{table(first)}
CUDA-event stage shares of the 1024-frame step: ICP {100 * st['icp'] / stot:.0f} %, front end {100 * st['preprocess'] / stot:.0f} %, plane {100 * st['plane'] / stot:.0f} %, cluster {100 * st['cluster'] / stot:.0f} %.

All launches of the file (with `--steps 1` the end-to-end arm is ONE handle with its internal pipeline: 4 sub-chunks of 256 frames
per call; under ncu they run one after the other):

{table(seq)}
## ncu `--set full` of `k_icp<256, 3>` (1024 problems, 148 persistent CTAs x 4 sub-workers, candidate table + seed grid + queued BVH search, taps off)

{md('k_icp')}
```
{lines('k_icp')}
```

Reading: a quarter of the round-1 instruction count, 26 of 32 lanes per instruction; half the issue slots are used and neither the
FP32 pipe (18 %) nor memory (DRAM 2 %, L2 hit rate 94 %) is near a limit: the kernel waits on dependent L2 loads (the point, then
its 64-byte table record) and on the barriers between the passes of an iteration. DESIGN.md section 9.

## ncu `--set full` of `k_frontend<0, 512, false, false, true>` (the `SOLO` instance: 1024 depth frames in one persistent launch, 296 CTAs, one-pass mode)

{md('k_frontend')}
```
{lines('k_frontend')}
```

Reading: half the issue slots, DRAM at 39 % of peak with 14 GB of traffic against 5.2 GB algorithmic (the radix ping-pong of 296
frames in flight does not stay in L2). Half the instructions are the three radix passes (a quarter of all instructions rank equal
digits with eight votes per record and pass). Two variants that move fewer bytes are in the library as opt-in knobs and are slower:
the voxel-hash path (`CUBOID_FE_HASH=1`: 8.4 GB, 7.1 ms) and run records (`CUBOID_FE_RUNS=1`: 9.2 GB, 4.74 ms); DESIGN.md section 8.

## ncu `--set full` of `k_sac_plane<256>` (1024 frames, one 256-thread CTA per frame)

{md('k_sac_plane')}
```
{lines('k_sac_plane')}
```

Reading: 37 % of the issue slots; a fifth of the stall samples wait on the nine sequential float accumulators of the refinement (PCL's
own accumulation order), the rest is the scoring loop (one warp per hypothesis over TMA-staged tiles) and the three ordered compactions.

## ncu `--set full` of `k_cluster` (1024 frames, one 1024-thread CTA per frame, fine-cell algorithm)

{md('k_cluster')}
```
{lines('k_cluster')}
```
"""
open(os.path.join(P, "README.md"), "w").write(out)
print("profiles/README.md", len(out), "bytes")
