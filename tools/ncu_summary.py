"""Developer tool: key metrics of an .ncu-rep (one kernel launch, `ncu --set full`) as a markdown table + a JSON dict.
usage: python tools/ncu_summary.py gpurun_out/prof_k_icp_r2_g.ncu-rep [--json out.json]"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[-1]  # header, units, the (single) launch
    return {h: (vals[i], units[i]) for i, h in enumerate(hdr)}


def main():
    rep = sys.argv[1]
    m = load(rep)
    print("| metric | unit | value |\n|---|---|---|")
    js = {"kernel": m.get("Kernel Name", ("", ""))[0]}
    for w in WANT:
        if w in m:
            v, u = m[w]
            print("| %s | %s | %s |" % (w, u, v))
            try:
                js[w] = float(v.replace(",", ""))
                js[w + ".unit"] = u
            except ValueError:
                js[w] = v
    if "--json" in sys.argv:
        json.dump(js, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
