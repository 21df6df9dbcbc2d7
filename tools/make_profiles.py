"""Developer tool: copies the artefacts of tools/gpu_final_profiles.sh from gpurun_out/ into profiles/ (tracked) and derives the two
small JSON files bench.py reads for its roofline objects:
  profiles/icp_issue.json         issue-slot use / lanes per instruction / instructions of k_icp from the committed ncu capture
  profiles/frontend_traffic.json  DRAM bytes of one k_frontend launch from the committed ncu capture
usage: python tools/make_profiles.py r2_g"""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
copied = []
for f in sorted(os.listdir(src)):
    if f.startswith(tag + "_") and f.endswith((".json", ".csv", ".md", ".txt")) and "ncu_launches" not in f:
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
        copied.append(f)
print("copied", len(copied), "files")


def unit_scale(u):
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


icp = json.load(open(os.path.join(src, tag + "_k_icp_summary.json")))
json.dump({
    "kernel": "k_icp", "workload": "full", "frames_per_launch": 1024, "capture": "profiles/%s_k_icp_summary.md" % tag,
    "source": "ncu --set full --clock-control none -k regex:^k_icp$ --launch-skip 1 -c 1, python bench.py --steps 1 --warmup 1 --no-cpu --no-configs (tools/gpu_final_profiles.sh)",
    "issue_active_pct": icp["smsp__issue_active.avg.pct_of_peak_sustained_active"],
    "lanes_per_instruction": icp["smsp__thread_inst_executed_per_inst_executed.ratio"],
    "warp_instructions_per_launch": icp["smsp__inst_executed.sum"],
    "warps_active_pct": icp["sm__warps_active.avg.pct_of_peak_sustained_active"],
    "fma_pipe_pct": icp["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"],
    "ms_under_ncu": icp["gpu__time_duration.sum"],
    "stall_long_scoreboard_per_issue": icp["smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"],
    "stall_barrier_per_issue": icp["smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"],
    "bound": "instruction issue / latency: neither the FP32 pipe nor memory bandwidth is close to its peak",
}, open(os.path.join(dst, "icp_issue.json"), "w"), indent=1)
fe = json.load(open(os.path.join(src, tag + "_k_frontend_summary.json")))
rd = fe["dram__bytes_read.sum"] * unit_scale(fe["dram__bytes_read.sum.unit"])
wr = fe["dram__bytes_write.sum"] * unit_scale(fe["dram__bytes_write.sum.unit"])
json.dump({
    "workload": "full", "frames_per_launch": 1024, "dram_bytes_per_launch": rd + wr, "dram__bytes_read.sum": rd, "dram__bytes_write.sum": wr,
    "capture": "profiles/%s_k_frontend_summary.md" % tag,
    "source": "ncu --set full --clock-control none -k regex:^k_frontend$ --launch-skip 1 -c 1, python bench.py --steps 1 --warmup 1 --no-cpu --no-configs (tools/gpu_final_profiles.sh)",
}, open(os.path.join(dst, "frontend_traffic.json"), "w"), indent=1)
print("wrote icp_issue.json, frontend_traffic.json")
