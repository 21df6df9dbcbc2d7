#!/usr/bin/env python
"""bench.py — VGA frames/s of the cuboid_detection hot path (plane segmentation + clustering + cuboid ICP).

Contract (driver): python bench.py --gpus N --steps K --warmup W  [--impl reference]
  * one "step" = one pass of the whole hot path over this rank's batch of synthetic VGA depth frames
    (stage 1a unprojection, PassThrough x2, VoxelGrid, RANSAC plane + ExtractIndices, Euclidean clustering,
    ICP of every cluster against the make_cuboid.py template) through libcuboid_cuda's C ABI;
  * `value`  : whole-job frames/s with the depth frames already resident in HBM, device time from CUDA events
    recorded on the library's own stream (max over ranks);
  * `e2e`    : the same metric through cuboid_process_batch with PINNED HOST depth buffers: host->device copy of
    every frame and device->host copy of every frame's result inside the timed region; steps are dealt round-robin over
    --e2e-handles library handles (default 3), one host thread each, so that one batch's copies run under another's ICP.
    `h2d_ceiling_gbs_per_gpu` is the bare pinned copy of the same bytes on every rank at once (nothing else running):
    what the box's PCIe / host memory allows, and `frac_of_h2d_ceiling` is the end-to-end figure against it;
  * `roofline`: the dominant kernel, k_icp. Its nominal roofline is the FP32 pipe with un-fused FMUL / FADD (bit-exactness
    forbids FFMA): `achieved` = the flops the kernel EXECUTED (8 per source-template pair it evaluated, counted by the kernel)
    over its CUDA-event time, `frac` = that over the un-fused peak measured in the same run. The brute-force-equivalent rate
     is kept under `algorithmic` as context only; `issue` (profiles/icp_issue.json, from the committed ncu
    capture) says what actually bounds the kernel: issue slots and latency. `roofline_hbm` is the HBM-bound fused front end
    (k_frontend: unproject + passthrough + voxel grid in one kernel, algorithmic bytes 2*P + 32*N + 16*V per frame)
    against MEASURED_PEAKS.json; its `traffic` is the dram__bytes figure of the committed ncu capture (profiles/);
  * `configs`: every other BASELINE.json config in the same line (1 GPU, default run): single-frame latency through
    cuboid_process_cloud with host buffers (p50 / p99 of 200 calls), ground-plane segmentation only, 64 hypotheses per
    cluster, 8 objects per frame, 1280x720 - each with value / e2e / stage ms and its own CPU baseline;
  * `cpu_baseline`: the CPU oracle (restatement of the PCL path) in LITERAL mode, 1 thread like the reference node, 3 warm-up
    frames + the median of 10 per-frame times; a sample of the same batch is also run through the
    oracle in canonical mode and re-run on the GPU with the parity taps on (the timed runs carry none): it must match the
    oracle's hashes, and the timed run's results in every other byte;
  * --impl reference: that CPU restatement (literal mode) on all host cores; each step = a bounded sample (2 x cores frames) of
    the same workload. The real PCL/ROS reference cannot be built offline; oracle/pcl_probe.py looks for a PCL at
    run time and, if it finds one, builds and times the literal PCL harness beside it. Rank 0 only;
  * --shard hypotheses (with --workload guess64): the low-latency multi-GPU mode - every rank runs the same frames with its
    slice of the initial-pose hypotheses, one all_gather of 80-byte records, exact arg-min (strong scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "VGA frames/s (plane seg + cuboid ICP)"
W, H = 640, 480
T_TEMPLATE = (0.2, 0.1, 0.03, 0.002)   # template_cuboid_L200_W100_H30_3faces.pcd, the launch default


def _measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None
        self.nvml, self.samples, self.stop_flag, self.max_mhz = None, [], False, None

    def start(self):
        """NVML polled every 5 ms from a thread (a 0.2 s timed region still gets dozens of samples); nvidia-smi -lms as fallback."""
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.gpu
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thr = threading.Thread(target=self._poll, daemon=True)
            self.thr.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM)),
                                     int(get_reasons(self.dev))))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Summary of the samples that arrived in [t0, t1] (the timed region); the sampler is started before the warm-up so
        that nvidia-smi is already streaming when the region begins."""
        if self.nvml:
            self.stop_flag = True
            self.thr.join(timeout=1)
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            sel = [(mhz, rs) for (ts, mhz, rs) in self.samples if (t0 is None or ts >= t0) and (t1 is None or ts <= t1)]
            reasons = sorted(nm for nm, bit in bits.items() if any(rs & bit for _, rs in sel))
            return {"sm_mhz": float(np.median([m for m, _ in sel])) if sel else None, "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(sel), "source": "nvml, 5 ms poll"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (ts, r) in self.rows if (t0 is None or ts >= t0) and (t1 is None or ts <= t1 + 0.25)]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _without_taps(r):
    """Bytes of a frame result with the parity-tap hashes that CUBOID_OPT_TAPS = 0 leaves at zero blanked."""
    c = type(r).from_buffer_copy(bytes(r))
    c.points_hash = c.voxel_key_hash = c.voxel_hash = 0
    for k in range(len(c.cluster)):
        c.cluster[k].corr_hash = 0
    return bytes(c)


def _bind_to_gpu_cpus(gpu_index):
    """Pin this process (and the host threads it starts later) to the CPUs NVML names as local to its GPU, before the pinned
    depth buffer is allocated, so that the buffer's pages and the threads feeding the copies sit on the GPU's NUMA node.
    A no-op on this pool's boxes (one NUMA node: 16 CPUs with 1 GPU, 32 CPUs with 8), where the 8-GPU end-to-end figure is
    bounded by the host side regardless (~170 GB/s of depth frames out of host memory); it matters on two-socket hosts."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[gpu_index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu_index
        dev = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(dev, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if len(cpus) >= 4:
            os.sched_setaffinity(0, cpus)
            return "%d cpus local to gpu %d" % (len(cpus), phys)
        return "unchanged (NVML names %d usable cpus)" % len(cpus)
    except Exception as e:      # no NVML, a restricted cpuset, ...: run unbound
        return "unchanged (%s)" % type(e).__name__


def make_frames(n, seed0, kind="bench"):
    from perception_b200 import synth
    return synth.depth_batch(kind, range(seed0, seed0 + n))


def template():
    from perception_b200 import pcd
    return pcd.template_points(*T_TEMPLATE)


def icp_flops(results, n_tmpl):
    """Algorithmic FP32 ops of the ICP kernel: 8*S*T per nearest-neighbour pass; iterations + 1 fitness pass."""
    ops = 0
    for r in results:
        for c in range(min(r.n_clusters, 16)):
            cl = r.cluster[c]
            ops += 8.0 * cl.size * n_tmpl * (cl.iterations + 1)
    return ops


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the PCL path on all host cores (oracle port; PCL itself is unbuildable here)."""
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pcl_probe
    from oracle import pyoracle as O
    from perception_b200.params import default_params
    cores = os.cpu_count() or 1
    pcl = pcl_probe.find_pcl()                     # a real PCL, if this box has one, pins the oracle and is timed beside it
    pcl_report = {"found": pcl["found"], "detail": pcl["detail"]}
    wl = WORKLOADS[args.workload]
    dp = default_params(wl["variant"])
    dp.n_guess, dp.guess_mode = wl["n_guess"], (1 if wl["n_guess"] > 1 else 0)
    p = O.params_from(dp)
    tm = template() if wl["stages"] & 8 else None
    rots = guess_rotations() if wl["n_guess"] > 1 else None
    per_step = max(cores, 2 * cores if args.frames >= 2 * cores else cores)
    frames = make_frames(per_step, 0, wl["kind"])

    def one(i):
        return O.process_frame(p, frames[i], tm, guesses=rots, mode=O.LITERAL)   # literal mode

    def step():
        with ThreadPoolExecutor(max_workers=cores) as ex:   # ctypes releases the GIL: frames run in parallel
            return list(ex.map(one, range(per_step)))

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.frames),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "mode": "literal",
                         "sample": "each step = the first %d frames of the bench workload (NOT the GPU arm's %d: a bounded sample, the figure is a "
                                   "rate), frames parallel over %d host threads" % (per_step, args.frames, cores)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the PCL path (oracle/), not a PCL binary: PCL/ROS cannot be built offline",
        "pcl_probe": pcl_report,
    }
    if pcl["found"] and args.workload == "full":
        try:   # the literal PCL call sequence on frame 0: its own timing and the restatement-vs-PCL difference
            import tempfile
            binary = pcl_probe.build_harness(pcl)
            if binary:
                pts = O.unproject(frames[0], dp.fx, dp.fy, dp.cx, dp.cy, dp.depth_scale)
                tpl = os.path.join(ROOT, "tests", "golden", "template_cuboid_L200_W100_H30_3faces.pcd")
                dump = pcl_probe.run_harness(binary, pts, tpl, tempfile.mkdtemp(prefix="pcl_harness_"))
                if dump:
                    ms = dump["ms_segmentation"] + dump["ms_icp"]
                    pcl_report.update({"frames_per_s_one_thread": 1e3 / ms if ms > 0 else None, "ms_segmentation": dump["ms_segmentation"],
                                       "ms_icp": dump["ms_icp"], "vs_restatement": pcl_probe.compare_with_oracle(dump, O.process_frame(p, frames[0], tm, mode=O.LITERAL))})
                    line["cpu_baseline"]["kind"] = "port (+ reference PCL timed on one frame: pcl_probe)"
        except Exception as e:
            pcl_report["error"] = repr(e)
    print(json.dumps(line), flush=True)


# BASELINE.json configs: "full" is the headline workload (configs[1]'s batch with the metric's full path);
# the others are secondary lines for profiles/ (python bench.py --workload seg|guess64|multi8|hd720).
WORKLOADS = {
    "full":    dict(kind="bench",  variant="cuboid", stages=15, n_guess=1,  w=640,  h=480,
                    text="table + one 200x100x30 mm cuboid; full hot path: unproject + passthrough + voxel(5 mm) + RANSAC plane + extract + "
                         "Euclidean clustering + ICP vs template_cuboid_L200_W100_H30_3faces (7250 pts), 1 initial-pose hypothesis"),
    "seg":     dict(kind="plane_var", variant="cuboid", stages=3, n_guess=1, w=640, h=480,
                    text="BASELINE configs[1]: ground-plane segmentation only (unproject + passthrough + voxel + RANSAC plane + extract)"),
    "guess64": dict(kind="bench",  variant="cuboid", stages=15, n_guess=64, w=640,  h=480,
                    text="BASELINE configs[2]: full path with 64 initial-pose hypotheses per cluster (8 yaw x 8 flips about the cluster centroid)"),
    "multi8":  dict(kind="multi8", variant="multi8", stages=15, n_guess=1,  w=640,  h=480,
                    text="BASELINE configs[3]: 8 cuboids per frame, pass_x +-0.4, one ICP per cluster"),
    "hd720":   dict(kind="hd720",  variant="hd720",  stages=15, n_guess=1,  w=1280, h=720,
                    text="BASELINE configs[4]: 1280x720, 2 mm leaf (~870k points, ~620k voxels per frame)"),
}


def guess_rotations():
    """identity + 63 rotations: 8 yaw steps x (4 roll quarter-turns x 2 pitch half-turns), applied about the cluster centroid."""
    def rx(a): return np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    def ry(a): return np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    def rz(a): return np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    out = []
    for r in range(4):
        for pch in range(2):
            for yw in range(8):
                out.append(rz(yw * np.pi / 4) @ ry(pch * np.pi) @ rx(r * np.pi / 2))
    out = np.asarray(out, dtype=np.float32)
    out[0] = np.eye(3, dtype=np.float32)
    return out


def workload_config(args, frames_per_gpu):
    wl = WORKLOADS[args.workload]
    return {"workload": "batch of %d synthetic %dx%d D435-shaped depth frames per GPU; %s" % (frames_per_gpu, wl["w"], wl["h"], wl["text"]),
            "name": args.workload, "frames_per_gpu": frames_per_gpu, "image": [wl["w"], wl["h"]], "template_points": 7250,
            "n_guess": wl["n_guess"],
            "l2": "inputs (%.0f MB of depth per GPU per step) are larger than the 126 MB L2" % (frames_per_gpu * wl["w"] * wl["h"] * 2 / 1e6),
            "parallelism": "frames sharded per GPU, no data-path collective"}


def _median(xs):
    return float(np.median(np.asarray(xs, dtype=np.float64))) if len(xs) else None


def cpu_baseline_literal(wl_name, frames, tm, rots, warm, runs):
    """the CPU restatement of the PCL path in LITERAL mode (std::sort voxel order, sequential Eigen-style
    sums, this host's libm), one thread like the reference's ros::spin(), `warm` untimed frames then the median over `runs`
    per-frame times. Template parsing is excluded (hoisted, as in the GPU path)."""
    from oracle import pyoracle as O
    from perception_b200.params import default_params
    wl = WORKLOADS[wl_name]
    dp = default_params(wl["variant"])
    dp.n_guess, dp.guess_mode = wl["n_guess"], (1 if wl["n_guess"] > 1 else 0)
    op = O.params_from(dp)
    otm = tm if wl["stages"] & 8 else None
    n = min(len(frames), warm + runs)
    times = []
    for i in range(n):
        t0 = time.perf_counter()
        O.process_frame(op, frames[i], otm, guesses=rots, mode=O.LITERAL)
        times.append(time.perf_counter() - t0)
    timed = times[min(warm, max(n - 1, 0)):]
    med = _median(timed)
    return {"value": 1.0 / med if med else None, "unit": "frames/s", "cores": 1, "kind": "port", "mode": "literal",
            "ms_per_frame_median": 1e3 * med if med else None,
            "sample": "%d warm-up + %d timed frames of this workload, one thread, median per-frame time "
                      "(CPU restatement of the PCL path, not a PCL binary)" % (min(warm, n), len(timed))}


def h2d_ceiling(torch, dist, world, host, dev, reps=3):
    """Bare pinned host -> device copy of one step's depth frames on every rank at once (one cudaMemcpyAsync per rank, nothing
    else running): the ceiling any end-to-end figure with these inputs can reach on this box. GB/s per GPU, min over ranks."""
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dev.copy_(host, non_blocking=True)
        b.record()
        b.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    gbs = host.numel() * host.element_size() / (best * 1e-3) / 1e9
    if world > 1:
        t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        gbs = float(t.item())
    return gbs


class Measurement:
    """One workload on this rank: the device-resident pass (value, stage times, rooflines) and the end-to-end pass."""
    pass


def measure(args, wl_name, F, chunk, steps, warmup, e2e_handles, local_rank, rank, world, torch, dist, with_ceiling=False):
    from perception_b200 import api
    from perception_b200.params import FrameResult, default_params
    import ctypes as C
    wl = WORKLOADS[wl_name]
    Wd, Ht = wl["w"], wl["h"]
    p = default_params(wl["variant"])
    p.n_guess, p.guess_mode = wl["n_guess"], (1 if wl["n_guess"] > 1 else 0)
    tm = template()
    t_gen = time.perf_counter()
    frames = make_frames(F, rank * F, wl["kind"])             # uint16 [F,h,w], unique seeds per rank
    t_gen = time.perf_counter() - t_gen
    host = torch.empty((F, Ht, Wd), dtype=torch.uint16, pin_memory=True)
    host.numpy()[...] = frames
    dev = host.to("cuda", non_blocking=False)                 # resident input for the kernel-only number

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    m = Measurement()
    m.wl, m.F, m.frames, m.tm, m.p = wl, F, frames, tm, p
    m.h2d_ceiling_gbs = h2d_ceiling(torch, dist, world, host, dev) if with_ceiling else None
    cc = api.CuboidCuda(p, device=local_rank, max_points=Wd * Ht, max_batch=min(chunk, F))
    cc.set_template(0, tm)
    rots = guess_rotations() if wl["n_guess"] > 1 else None
    m.rots = rots
    if rots is not None:
        cc.set_guesses(rots, mode=1)
    stages = wl["stages"]
    cc.set_option(api.OPT_TAPS, 0)   # no parity taps in the timed runs (key / count arrays, point / voxel / correspondence hashes); see cpu_baseline
    m.peak_unfused, m.peak_ffma = cc.measure_fp32_peak()

    # ---- device-resident: value + rooflines ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(warmup):
        cc.process_batch_device(dev.data_ptr(), Wd, Ht, F, stages=stages)
    barrier()
    l0 = cc.launch_count()
    stage = {k: 0.0 for k in ("preprocess", "voxel", "plane", "cluster", "icp")}
    wall0 = time.perf_counter()
    for _ in range(steps):
        cc.process_batch_device(dev.data_ptr(), Wd, Ht, F, stages=stages)
        for k, v in cc.stage_ms().items():
            stage[k] += v
    barrier()
    m.wall = time.perf_counter() - wall0
    m.launches = cc.launch_count() - l0
    m.clocks = sampler.stop(wall0, wall0 + m.wall)
    m.stage = stage
    m.dev_ms = sum(stage.values())                             # CUDA events on the library's stream, summed over chunks
    m.res = cc.batch_results(F)
    m.work_eval, m.work_brute = cc.icp_work()                  # pairs evaluated / brute-force-equivalent pairs, last step
    m.n_chunks = (F + cc.max_batch - 1) // cc.max_batch
    m.max_batch = cc.max_batch

    # ---- end to end: pinned host depth in, host results out ----
    # (a) one handle, strictly serial calls; (b) the way a throughput user drives the library: one handle per host thread
    # (a handle is thread-compatible, one call in flight), steps dealt round-robin, so the depth copy and front end of one
    # step overlap the ICP tail of the other. Every step's H2D copy and D2H result read are inside the timed region in both.
    cc.set_option(api.OPT_STAGES, stages)   # the host-buffer entry runs the same stage set
    for _ in range(min(warmup, 1)):
        cc.process_batch(host)
    barrier()
    e0 = time.perf_counter()
    for _ in range(steps):
        res_e2e = cc.process_batch(host)
    barrier()
    m.e2e_serial_s = time.perf_counter() - e0
    m.e2e_s, m.n_handles = m.e2e_serial_s, 1
    if e2e_handles > 1 and steps > 1:
        handles = [cc]
        pipe = int(os.environ.get("CUBOID_E2E_PIPELINE", "0"))
        cc.set_option(api.OPT_PIPELINE, pipe)           # handles overlap each other: chunk-wide launches inside each
        for _ in range(e2e_handles - 1):
            hx = None
            try:
                hx = api.CuboidCuda(p, device=local_rank, max_points=Wd * Ht, max_batch=min(chunk, F))
                hx.set_template(0, tm)
                if rots is not None:
                    hx.set_guesses(rots, mode=1)
                hx.set_option(api.OPT_STAGES, stages)
                hx.set_option(api.OPT_TAPS, 0)
                hx.set_option(api.OPT_PIPELINE, pipe)
                hx.process_batch(host)                  # warm-up of the extra handle (allocates its ICP scratch)
            except api.CuboidError as e:                # a handle holds a whole resident chunk: the big workloads fit fewer of them
                print("bench.py: end-to-end arm continues with %d handle(s): %s" % (len(handles), e), file=sys.stderr)
                if hx is not None:
                    hx.close()
                break
            handles.append(hx)
        if len(handles) > 1:
            out = [None] * steps

            def drive(k):
                for sidx in range(k, steps, len(handles)):
                    out[sidx] = handles[k].process_batch(host)

            barrier()
            e0 = time.perf_counter()
            thr = [threading.Thread(target=drive, args=(k,)) for k in range(len(handles))]
            for t in thr:
                t.start()
            for t in thr:
                t.join()
            barrier()
            m.e2e_s, m.n_handles = time.perf_counter() - e0, len(handles)
            same_all = all(bytes(a) == bytes(b) for o in out for a, b in zip(o, res_e2e))
            for hx in handles[1:]:
                hx.close()
            cc.set_option(api.OPT_PIPELINE, 1)
            if not same_all:
                raise SystemExit("bench.py: concurrent handles returned different results")
        else:
            cc.set_option(api.OPT_PIPELINE, 1)          # nothing to overlap with: the serial figure stands
    m.res_e2e = res_e2e
    m.h2d_bytes, m.d2h_bytes = F * Wd * Ht * 2, F * C.sizeof(FrameResult)
    m.t_gen = t_gen
    m.cc = cc

    # max over ranks (device time), sum of frames
    t_dev, t_e2e, t_wall, e2e_serial = m.dev_ms / 1e3, m.e2e_s, m.wall, m.e2e_serial_s
    if world > 1:
        t = torch.tensor([t_dev, t_e2e, t_wall, e2e_serial], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, t_wall, e2e_serial = [float(x) for x in t.cpu()]
        cnt = torch.tensor([F], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt)
        m.total_frames = int(cnt.item())
    else:
        m.total_frames = F
    m.t_dev, m.t_e2e, m.t_wall, m.t_e2e_serial = t_dev, t_e2e, t_wall, e2e_serial
    m.steps = steps
    return m


def single_frame_latency(local_rank, calls=200, warm=20):
    """BASELINE.json configs[0] on the GPU: ONE frame per call, the way a ROS node calls the library (subscriber queue depth 1,
    gps.cpp:146): cuboid_process_cloud on a host PointCloud2-shaped blob (x, y, z float32 at offsets 0 / 4 / 8, point_step 16),
    every call's host -> device copy of the 4.9 MB cloud and the result read-back inside its time. p50 / p99 of wall-clock
    per-call times; the depth-frame entry (cuboid_process_batch with one 614 KB frame, stage 1a on the device) beside it."""
    from perception_b200 import api, synth
    from perception_b200.params import default_params
    p = default_params("cuboid")
    depth = synth.depth_frame("cuboid1", 0)
    tm = template()
    out = {}
    with api.CuboidCuda(p, device=local_rank, max_points=depth.size, max_batch=1) as cc:
        cc.set_template(0, tm)
        cc.set_option(api.OPT_TAPS, 0)
        cloud = np.ascontiguousarray(cc.unproject(depth))      # what realsense2_camera would publish: all 307 200 points, xyz + pad
        try:                                                   # the same message in page-locked memory (a node that cares pins its buffer)
            import torch
            pinned = torch.from_numpy(cloud).pin_memory().numpy()
        except Exception:
            pinned = None
        cases = [("process_cloud", lambda: cc.process_cloud(cloud)), ("process_batch_1", lambda: cc.process_batch(depth[None]))]
        if pinned is not None:
            cases.append(("process_cloud_pinned", lambda: cc.process_cloud(pinned)))
        for name, fn in cases:
            for _ in range(warm):
                r = fn()
            ts = []
            for _ in range(calls):
                t0 = time.perf_counter()
                r = fn()
                ts.append(time.perf_counter() - t0)
            ts = np.asarray(ts) * 1e3
            r0 = r[0] if name == "process_batch_1" else r
            out[name] = {"p50_ms": float(np.percentile(ts, 50)), "p99_ms": float(np.percentile(ts, 99)), "mean_ms": float(ts.mean()),
                         "calls": calls, "frames_per_s_at_p50": 1e3 / float(np.percentile(ts, 50)),
                         "h2d_bytes_per_call": int(depth.nbytes if name == "process_batch_1" else cloud.nbytes),
                         "host_buffer": "pinned" if name == "process_cloud_pinned" else "pageable (numpy)",
                         "icp_iterations": int(r0.cluster[0].iterations), "accepted": int(r0.cluster[0].accepted)}
    out["workload"] = ("BASELINE configs[0]: single synthetic 640x480 frame (cuboid1, seed 0), voxel + RANSAC plane + cluster + ICP vs the "
                       "7250-point template, one call per frame with host buffers")
    return out


def summarize(m, world):
    """Line fragments shared by the headline workload and the secondary configs."""
    F, steps, res = m.F, m.steps, m.res
    n_pts = sum(r.n_points for r in res)
    n_vox = sum(r.n_voxels for r in res)
    return {
        "value": m.total_frames * steps / m.t_dev, "ms_per_step": 1e3 * m.t_dev / steps,
        "e2e_value": m.total_frames * steps / m.t_e2e, "e2e_serial_calls_value": m.total_frames * steps / m.t_e2e_serial,
        "stages_ms_per_step": {k: v / steps for k, v in m.stage.items()},
        "frame_stats": {"mean_points": n_pts / F, "mean_voxels": n_vox / F, "mean_remain": sum(r.n_remain for r in res) / F,
                        "mean_clusters": sum(r.n_clusters for r in res) / F,
                        "mean_icp_iterations": float(np.mean([r.cluster[c].iterations for r in res for c in range(min(r.n_clusters, 16))] or [0])),
                        "accepted": int(sum(r.cluster[c].accepted for r in res for c in range(min(r.n_clusters, 16)))),
                        "e2e_equals_device": bool(all(bytes(a) == bytes(b) for a, b in zip(res, m.res_e2e)))},
    }


def run_hypothesis_sharded(args, local_rank, rank, world, torch, dist):
    """SURVEY.md 8e, second axis: the hypotheses of every cluster are split over the GPUs. A step = cuboid_process_batch on the
    (same) host frames with this rank's slice of the hypotheses + one all_gather of cuboid_guess_record + the exact arg-min.
    Timed with a barrier on both sides, max over ranks; with --frames 1 the step time is the single-frame latency."""
    from perception_b200 import api
    from perception_b200 import dist as pd
    from perception_b200.params import default_params
    wl = WORKLOADS[args.workload]
    if wl["n_guess"] < 2:
        raise SystemExit("--shard hypotheses needs a workload with several hypotheses (--workload guess64)")
    Wd, Ht, F = wl["w"], wl["h"], args.frames
    p = default_params(wl["variant"])
    rots = guess_rotations()
    p.n_guess, p.guess_mode = len(rots), 1
    tm = template()
    frames = make_frames(F, 0, wl["kind"])                       # the same frames on every rank
    host = torch.empty((F, Ht, Wd), dtype=torch.uint16, pin_memory=True)
    host.numpy()[...] = frames
    g0, cnt = pd.shard_range(len(rots), rank, world)
    cc = api.CuboidCuda(p, device=local_rank, max_points=Wd * Ht, max_batch=min(args.chunk, F))
    cc.set_template(0, tm)
    cc.set_guesses(rots[g0:g0 + cnt], mode=1)
    cc.set_guess_offset(g0)
    cc.set_option(api.OPT_TAPS, 0)
    cc.set_option(api.OPT_STAGES, wl["stages"])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        res = pd.reduce_hypotheses(cc.process_batch(host), p.icp_fitness_gate)
    barrier()
    l0 = cc.launch_count()
    times = []
    w0 = time.perf_counter()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        res = pd.reduce_hypotheses(cc.process_batch(host), p.icp_fitness_gate)
        times.append(time.perf_counter() - t0)
    barrier()
    wall = time.perf_counter() - w0
    launches = cc.launch_count() - l0
    clocks = sampler.stop(w0, w0 + wall)
    t = wall
    if world > 1:
        tt = torch.tensor([wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt.item())
    import hashlib
    digest = hashlib.sha256(b"".join(bytes(r) for r in res)).hexdigest()[:16]
    if rank == 0:
        import ctypes as C
        from perception_b200.params import FrameResult
        line = {"metric": METRIC, "value": F * args.steps / t, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": dict(workload_config(args, F), parallelism="every GPU runs the same %d frame(s) with %d of the %d "
                "initial-pose hypotheses per cluster; one all_gather of 80-byte records per step, exact (fitness, guess id) arg-min" % (F, cnt, len(rots))),
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": F * args.steps / t, "unit": "frames/s", "h2d_bytes_per_step": F * Wd * Ht * 2, "d2h_bytes_per_step": F * C.sizeof(FrameResult),
                        "note": "the timed step IS the end-to-end call: pinned host depth in, results to host, records all-gathered over %s" % ("NCCL" if world > 1 else "nothing (1 GPU)")},
                "latency_ms": {"p50": 1e3 * float(np.percentile(times, 50)), "max": 1e3 * float(np.max(times)), "frames_per_step": F},
                "results_digest": digest,
                "frame_stats": {"best_guess_frame0": int(res[0].cluster[0].best_guess) if res[0].n_clusters else None,
                                "fitness_frame0": float(res[0].cluster[0].fitness) if res[0].n_clusters else None}}
        print(json.dumps(line), flush=True)
    cc.close()


def _profile_json(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None


# frames per step / steps of the secondary BASELINE configs inside the default run (the headline config runs at --frames / --steps).
# The per-frame kernels (one CTA per frame in the front end, plane and cluster stages) are latency bound per frame, so throughput
# needs enough frames in flight to fill 148 SMs with several CTAs each: 256 frames of 720p, 1024 of VGA.
CONFIG_RUNS = {"seg": (1024, 5), "guess64": (256, 3), "multi8": (1024, 3), "hd720": (256, 3)}
CONFIG_CPU = {"full": (3, 10), "seg": (3, 10), "guess64": (1, 3), "multi8": (3, 10), "hd720": (1, 3)}   # (warm-ups, timed frames)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (default: 1024, or the config's own size for --workload)")
    ap.add_argument("--chunk", type=int, default=1024, help="frames resident per chunk (max_batch)")
    ap.add_argument("--cpu-sample", type=int, default=64, help="frames of the oracle parity sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary BASELINE configs and the single-frame latency")
    ap.add_argument("--e2e-handles", type=int, default=3,
                    help="library handles (one host thread each) the end-to-end loop spreads its steps over; 1 = strictly serial calls")
    ap.add_argument("--workload", default="full", choices=sorted(WORKLOADS))
    ap.add_argument("--shard", default="frames", choices=["frames", "hypotheses"],
                    help="frames: every rank takes its own block of frames (weak scaling, the default); hypotheses: every rank runs the "
                         "SAME frames with its slice of the initial-pose hypotheses and one all-gather of 80-byte records picks the winner "
                         "(strong scaling / single-frame latency; use with --workload guess64)")
    args = ap.parse_args()
    global W, H
    wl = WORKLOADS[args.workload]
    W, H = wl["w"], wl["h"]
    if args.frames <= 0:
        args.frames = 1024 if args.workload in ("full", "seg") else CONFIG_RUNS[args.workload][0]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from perception_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libcuboid_cuda has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = _bind_to_gpu_cpus(local_rank) if os.environ.get("CUBOID_BENCH_BIND", "1") != "0" else "off"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    F = args.frames
    if args.shard == "hypotheses":
        run_hypothesis_sharded(args, local_rank, rank, world, torch, dist)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    m = measure(args, args.workload, F, args.chunk, args.steps, args.warmup, args.e2e_handles, local_rank, rank, world, torch, dist,
                with_ceiling=True)
    cc, res, tm, p, frames, rots, stage = m.cc, m.res, m.tm, m.p, m.frames, m.rots, m.stage
    stages = wl["stages"]

    if rank == 0:
        peaks, peak_src = _measured_peaks()
        sm = summarize(m, world)
        n_chunks = m.n_chunks
        ops = icp_flops(res, len(tm))                      # brute-force equivalent: 8*S*T per nearest-neighbour pass
        if wl["n_guess"] > 1:
            ops = 8.0 * m.work_brute                       # every hypothesis counts, not only the winner reported per cluster
        ops_exec = 8.0 * m.work_eval                       # what the culled kernel actually executed
        icp_s = stage["icp"] / 1e3 / args.steps
        executed = ops_exec / icp_s / 1e12 if icp_s > 0 else 0.0
        brute_equiv = ops / icp_s / 1e12 if icp_s > 0 else 0.0
        n_pts = sum(r.n_points for r in res)
        pre_bytes = 2.0 * F * W * H + 16.0 * n_pts
        pre_s = stage["preprocess"] / 1e3 / args.steps
        n_vox = sum(r.n_voxels for r in res)
        vox_bytes = 16.0 * n_pts + 16.0 * n_vox
        vox_s = stage["voxel"] / 1e3 / args.steps
        fused = vox_s * 50 < pre_s          # the fused front end reports its whole time under "preprocess"
        fe_bytes = pre_bytes + (vox_bytes if fused else 0.0)
        fe_gbs = fe_bytes / pre_s / 1e9 if pre_s > 0 else 0.0
        traffic = None
        tj = _profile_json("frontend_traffic.json")   # per-launch DRAM bytes of the same launch under `ncu --set full` (profiles/README.md)
        if tj and tj.get("frames_per_launch") == min(m.max_batch, F) and tj.get("workload") == args.workload:
            traffic = tj["dram_bytes_per_launch"]
        issue = _profile_json("icp_issue.json")       # issue-slot use and lanes per instruction of k_icp from the committed ncu capture
        e2e_gbs = m.h2d_bytes * args.steps / m.t_e2e / 1e9      # per GPU
        line = {
            "metric": METRIC, "value": sm["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sm["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, F),
            "clocks": m.clocks,
            "e2e": {"value": sm["e2e_value"], "unit": "frames/s", "h2d_bytes_per_step": m.h2d_bytes,
                    "d2h_bytes_per_step": m.d2h_bytes, "handles_per_gpu": m.n_handles, "cpu_binding": numa,
                    "serial_calls_value": sm["e2e_serial_calls_value"],
                    "h2d_gbs_per_gpu": e2e_gbs, "h2d_ceiling_gbs_per_gpu": m.h2d_ceiling_gbs,
                    "frac_of_h2d_ceiling": (e2e_gbs / m.h2d_ceiling_gbs) if m.h2d_ceiling_gbs else None,
                    "note": "cuboid_process_batch on pinned host depth, results to host; value = steps dealt round-robin over "
                            "%d handle(s), one host thread each (CUBOID_OPT_PIPELINE=0: chunk-wide launches inside a handle, the "
                            "handles overlap each other's copies); serial_calls_value = one handle with its internal sub-chunk "
                            "pipeline, one call after the other. h2d_ceiling = the same depth bytes copied by a bare cudaMemcpyAsync on "
                            "every rank at once with nothing else running (min over ranks). `value` times ONE handle with its stages "
                            "back to back on one stream, so the overlapped multi-handle figure can come out slightly above it (ICP "
                            "tails of one batch are filled by the next batch's front end)" % m.n_handles},
            "gpu_launches": int(m.launches),
            # k_icp is the dominant kernel. Its roofline is the FP32 pipe with un-fused FMUL / FADD (bit-exactness forbids FFMA).
            # achieved = the flops the kernel EXECUTED (8 per source-template pair it evaluated, counted by the kernel itself) over its
            # CUDA-event time; frac = achieved / the un-fused peak measured in this run. The kernel returns brute force's exact answer
            # but proves most pairs irrelevant with an exact box bound, so the brute-force-equivalent rate (SURVEY.md 8d's 8*S*T per
            # pass) is far above the peak: it is kept under `algorithmic` as context and is NOT a roofline fraction. What bounds
            # the kernel is instruction issue and lane use, not the FP32 pipe: see `issue`.
            "roofline": {"kernel": "k_icp", "bound": "fp32", "achieved": executed, "peak": m.peak_unfused, "unit": "TFLOP/s",
                         "frac": executed / m.peak_unfused if m.peak_unfused else None, "traffic": None,
                         "peak_source": "un-fused FMUL+FADD micro-benchmark in this run (bit-exactness forbids FFMA); FFMA peak %.1f" % m.peak_ffma,
                         "launches_per_step": n_chunks, "ms_per_launch": 1e3 * icp_s / n_chunks,
                         "executed_flops_per_step": ops_exec,
                         "algorithmic": {"flops_per_step": ops, "tflops_equivalent": brute_equiv,
                                         "culled_fraction": 1.0 - m.work_eval / max(m.work_brute, 1),
                                         "note": "brute-force equivalent (8*S*T per nearest-neighbour pass); exact culling skips pairs that "
                                                 "provably cannot win, results are bit-identical to the brute-force scan "
                                                 "(CUBOID_OPT_ICP_CULL=0), see tests"},
                         "issue": issue},
            "roofline_hbm": {"kernel": "k_frontend" if fused else "k_preprocess", "bound": "hbm", "achieved": fe_gbs, "peak": peaks.get("hbm_gbs"),
                             "unit": "GB/s", "frac": fe_gbs / peaks.get("hbm_gbs") if peaks.get("hbm_gbs") else None, "traffic": traffic,
                             "traffic_over_algorithmic": (traffic / (fe_bytes / n_chunks)) if traffic else None,
                             "peak_source": peak_src, "launches_per_step": n_chunks, "ms_per_launch": 1e3 * pre_s / n_chunks,
                             "algorithmic_bytes_per_step": fe_bytes,
                             "algorithmic_bytes_per_frame": "2*P + 16*N (unproject + passthrough) + 16*N + 16*V (voxel grid)" if fused
                             else "2*P + 16*N (unproject + passthrough)"},
            "stages_ms_per_step": sm["stages_ms_per_step"],
            "wall_ms_per_step": 1e3 * m.t_wall / args.steps,
            "frame_stats": sm["frame_stats"],
            "input_generation_s": m.t_gen,
        }
        if world == 1 and not args.no_cpu:
            from oracle import pyoracle as O
            op = O.params_from(p)
            ns = min(args.cpu_sample, F)
            otm = tm if stages & 8 else None
            if wl["n_guess"] > 1 or wl["w"] > 640:
                ns = min(ns, 2)                          # a 64-hypothesis or 720p oracle frame takes seconds
            c0 = time.perf_counter()
            cpu_res = [O.process_frame(op, frames[i], otm, guesses=rots) for i in range(ns)]
            cdt = time.perf_counter() - c0
            # The timed runs carry no parity taps (CUBOID_OPT_TAPS = 0 leaves the point / voxel / correspondence hashes at 0). The
            # sample is run once more with the taps on: its hashes must equal the oracle's (every point, voxel, inlier and every
            # ICP iteration's correspondences) and everything else must equal the timed run's results byte for byte.
            cc.set_option(api.OPT_TAPS, 1)
            chk = cc.process_batch(np.ascontiguousarray(frames[:ns]))
            cc.set_option(api.OPT_TAPS, 0)
            same = all(cpu_res[i].cluster[0].corr_hash == chk[i].cluster[0].corr_hash and cpu_res[i].inlier_hash == chk[i].inlier_hash
                       and cpu_res[i].points_hash == chk[i].points_hash and cpu_res[i].voxel_hash == chk[i].voxel_hash
                       and _without_taps(chk[i]) == _without_taps(res[i]) for i in range(ns))
            cb = cpu_baseline_literal(args.workload, frames, tm, rots, *CONFIG_CPU[args.workload])
            cb["gpu_matches_oracle_on_sample"] = bool(same)
            cb["parity_sample"] = "first %d frames, oracle in canonical mode (the mode the GPU is bit-exact against): %.1f frames/s on one thread" % (ns, ns / cdt)
            line["cpu_baseline"] = cb
    cc.close()
    del m

    # ---- the other BASELINE.json configs (rank 0 of a 1-GPU run): few steps each, same measurement, same JSON line ----
    if world == 1 and rank == 0 and not args.no_configs and args.workload == "full":
        cfgs = {}
        try:
            cfgs["single_frame"] = single_frame_latency(local_rank)
            if not args.no_cpu:
                from perception_b200 import synth
                cfgs["single_frame"]["cpu_baseline"] = cpu_baseline_literal("full", synth.depth_batch("cuboid1", [0] * 13), tm, None, 3, 10)
        except Exception as e:   # a secondary line must never take the headline down
            cfgs["single_frame"] = {"error": repr(e)}
        for name, (cf, cs) in CONFIG_RUNS.items():
            try:
                ca = argparse.Namespace(**vars(args))
                ca.workload = name
                W, H = WORKLOADS[name]["w"], WORKLOADS[name]["h"]
                mm = measure(ca, name, cf, args.chunk, cs, 3, min(args.e2e_handles, 2), local_rank, rank, world, torch, dist)
                sm2 = summarize(mm, world)
                entry = {"workload": WORKLOADS[name]["text"], "frames_per_step": cf, "steps": cs, "warmup": 3,
                         "value": sm2["value"], "unit": "frames/s", "ms_per_step": sm2["ms_per_step"],
                         "e2e": {"value": sm2["e2e_value"], "serial_calls_value": sm2["e2e_serial_calls_value"], "handles_per_gpu": mm.n_handles,
                                 "h2d_bytes_per_step": mm.h2d_bytes, "d2h_bytes_per_step": mm.d2h_bytes},
                         "stages_ms_per_step": sm2["stages_ms_per_step"], "frame_stats": sm2["frame_stats"],
                         "gpu_launches": int(mm.launches)}
                if not args.no_cpu:
                    entry["cpu_baseline"] = cpu_baseline_literal(name, mm.frames, mm.tm, mm.rots, *CONFIG_CPU[name])
                mm.cc.close()
                del mm
                cfgs[name] = entry
            except Exception as e:
                cfgs[name] = {"error": repr(e)}
        line["configs"] = cfgs
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
