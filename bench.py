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
    --e2e-handles library handles (default 3), one host thread each, so that one batch's copies run under another's ICP;
  * `roofline`: the dominant kernel (k_icp, FP32-pipe bound: un-fused FMUL/FADD) — algorithmic
    flops from the per-frame results (8*S*T per nearest-neighbour pass) over its CUDA-event time, against the
    un-fused FP32 peak measured in the same run; `roofline_hbm` is the same for the HBM-bound fused front end
    (k_frontend: unproject + passthrough + voxel grid in one kernel, algorithmic bytes 2*P + 32*N + 16*V per frame)
    against MEASURED_PEAKS.json; its `traffic` is the dram__bytes figure of the committed ncu capture (profiles/);
  * `cpu_baseline`: the CPU oracle (restatement of the PCL path, 1 thread like the reference node) on a bounded
    sample of the same frames; the same sample is re-run on the GPU with the parity taps on (the timed runs carry
    none) and must match the oracle's hashes, and the timed run's results in every other byte;
  * --impl reference: that CPU restatement on all host cores (the real PCL/ROS reference cannot be built
    offline: DESIGN.md), same metric/config, rank 0 only.
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
    from oracle import pyoracle as O
    from perception_b200.params import default_params
    cores = os.cpu_count() or 1
    wl = WORKLOADS[args.workload]
    dp = default_params(wl["variant"])
    dp.n_guess, dp.guess_mode = wl["n_guess"], (1 if wl["n_guess"] > 1 else 0)
    p = O.params_from(dp)
    tm = template() if wl["stages"] & 8 else None
    rots = guess_rotations() if wl["n_guess"] > 1 else None
    per_step = max(cores, 2 * cores if args.frames >= 2 * cores else cores)
    frames = make_frames(per_step, 0, wl["kind"])

    def one(i):
        return O.process_frame(p, frames[i], tm, guesses=rots)

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
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "%d frames per step of the bench workload, frames parallel over %d host threads" % (per_step, cores)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the PCL path (oracle/), not a PCL binary: PCL/ROS cannot be built offline",
    }
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--frames", type=int, default=1024, help="frames per GPU per step")
    ap.add_argument("--chunk", type=int, default=1024, help="frames resident per chunk (max_batch)")
    ap.add_argument("--cpu-sample", type=int, default=64, help="frames of the CPU baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-handles", type=int, default=3,
                    help="library handles (one host thread each) the end-to-end loop spreads its steps over; 1 = strictly serial calls")
    ap.add_argument("--workload", default="full", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    global W, H
    wl = WORKLOADS[args.workload]
    W, H = wl["w"], wl["h"]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from perception_b200 import api
    from perception_b200.params import FrameResult, default_params
    import ctypes as C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libcuboid_cuda has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = _bind_to_gpu_cpus(local_rank) if os.environ.get("CUBOID_BENCH_BIND", "1") != "0" else "off"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    F = args.frames
    p = default_params(wl["variant"])
    p.n_guess, p.guess_mode = wl["n_guess"], (1 if wl["n_guess"] > 1 else 0)
    tm = template()
    t_gen = time.perf_counter()
    frames = make_frames(F, rank * F, wl["kind"])             # uint16 [F,h,w], unique seeds per rank
    t_gen = time.perf_counter() - t_gen
    host = torch.empty((F, H, W), dtype=torch.uint16, pin_memory=True)
    host.numpy()[...] = frames
    dev = host.to("cuda", non_blocking=False)                 # resident input for the kernel-only number

    cc = api.CuboidCuda(p, device=local_rank, max_points=W * H, max_batch=min(args.chunk, F))
    cc.set_template(0, tm)
    rots = guess_rotations() if wl["n_guess"] > 1 else None
    if rots is not None:
        cc.set_guesses(rots, mode=1)
    stages = wl["stages"]
    cc.set_option(api.OPT_TAPS, 0)   # no parity taps in the timed runs (key / count arrays, point / voxel / correspondence hashes); see cpu_baseline
    peak_unfused, peak_ffma = cc.measure_fp32_peak()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident: value + rooflines ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        cc.process_batch_device(dev.data_ptr(), W, H, F, stages=stages)
    barrier()
    l0 = cc.launch_count()
    stage = {k: 0.0 for k in ("preprocess", "voxel", "plane", "cluster", "icp")}
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        cc.process_batch_device(dev.data_ptr(), W, H, F, stages=stages)
        for k, v in cc.stage_ms().items():
            stage[k] += v
    barrier()
    wall = time.perf_counter() - wall0
    launches = cc.launch_count() - l0
    clocks = sampler.stop(wall0, wall0 + wall)
    dev_ms = sum(stage.values())                               # CUDA events on the library's stream, summed over chunks
    res = cc.batch_results(F)
    work_eval, work_brute = cc.icp_work()                      # pairs evaluated / brute-force-equivalent pairs, last step

    # ---- end to end: pinned host depth in, host results out ----
    # (a) one handle, strictly serial calls; (b) the way a throughput user drives the library: one handle per host thread
    # (a handle is thread-compatible, one call in flight), steps dealt round-robin, so the depth copy and front end of one
    # step overlap the ICP tail of the other. Every step's H2D copy and D2H result read are inside the timed region in both.
    cc.set_option(api.OPT_STAGES, stages)   # the host-buffer entry runs the same stage set
    for _ in range(min(args.warmup, 1)):
        cc.process_batch(host)
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        res_e2e = cc.process_batch(host)
    barrier()
    e2e_serial_s = time.perf_counter() - e0
    e2e_s, n_handles = e2e_serial_s, 1
    if args.e2e_handles > 1 and args.steps > 1:
        handles = [cc]
        pipe = int(os.environ.get("CUBOID_E2E_PIPELINE", "0"))
        cc.set_option(api.OPT_PIPELINE, pipe)           # handles overlap each other: chunk-wide launches inside each
        for _ in range(args.e2e_handles - 1):
            hx = None
            try:
                hx = api.CuboidCuda(p, device=local_rank, max_points=W * H, max_batch=min(args.chunk, F))
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
            out = [None] * args.steps

            def drive(k):
                for sidx in range(k, args.steps, len(handles)):
                    out[sidx] = handles[k].process_batch(host)

            barrier()
            e0 = time.perf_counter()
            thr = [threading.Thread(target=drive, args=(k,)) for k in range(len(handles))]
            for t in thr:
                t.start()
            for t in thr:
                t.join()
            barrier()
            e2e_s, n_handles = time.perf_counter() - e0, len(handles)
            same_all = all(bytes(a) == bytes(b) for o in out for a, b in zip(o, res_e2e))
            for hx in handles[1:]:
                hx.close()
            cc.set_option(api.OPT_PIPELINE, 1)
            if not same_all:
                raise SystemExit("bench.py: concurrent handles returned different results")
        else:
            cc.set_option(api.OPT_PIPELINE, 1)          # nothing to overlap with: the serial figure stands

    # max over ranks (device time), sum of frames
    t_dev, t_e2e, t_wall = dev_ms / 1e3, e2e_s, wall
    if world > 1:
        t = torch.tensor([t_dev, t_e2e, t_wall, e2e_serial_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, t_wall, e2e_serial_s = [float(x) for x in t.cpu()]
        cnt = torch.tensor([F], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt)
        total_frames = int(cnt.item())
    else:
        total_frames = F

    if rank == 0:
        peaks, peak_src = _measured_peaks()
        n_chunks = (F + cc.max_batch - 1) // cc.max_batch
        ops = icp_flops(res, len(tm))                      # brute-force equivalent: 8*S*T per nearest-neighbour pass
        if wl["n_guess"] > 1:
            ops = 8.0 * work_brute                         # every hypothesis counts, not only the winner reported per cluster
        ops_exec = 8.0 * work_eval                         # what the culled kernel actually executed
        icp_s = stage["icp"] / 1e3 / args.steps
        achieved = ops_exec / icp_s / 1e12 if icp_s > 0 else 0.0
        effective = ops / icp_s / 1e12 if icp_s > 0 else 0.0
        n_pts = sum(r.n_points for r in res)
        pre_bytes = 2.0 * F * W * H + 16.0 * n_pts
        pre_s = stage["preprocess"] / 1e3 / args.steps
        pre_gbs = pre_bytes / pre_s / 1e9 if pre_s > 0 else 0.0
        n_vox = sum(r.n_voxels for r in res)
        vox_bytes = 16.0 * n_pts + 16.0 * n_vox
        vox_s = stage["voxel"] / 1e3 / args.steps
        fused = vox_s * 50 < pre_s          # the fused front end reports its whole time under "preprocess"
        fe_bytes = pre_bytes + (vox_bytes if fused else 0.0)
        fe_gbs = fe_bytes / pre_s / 1e9 if pre_s > 0 else 0.0
        traffic = None
        try:   # per-launch DRAM bytes of the same launch under `ncu --set full` (profiles/README.md says which capture)
            tj = json.load(open(os.path.join(ROOT, "profiles", "frontend_traffic.json")))
            if tj.get("frames_per_launch") == min(cc.max_batch, F) and tj.get("workload") == args.workload:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": total_frames * args.steps / t_dev, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, F),
            "clocks": clocks,
            "e2e": {"value": total_frames * args.steps / t_e2e, "unit": "frames/s", "h2d_bytes_per_step": F * W * H * 2,
                    "d2h_bytes_per_step": F * C.sizeof(FrameResult), "handles_per_gpu": n_handles, "cpu_binding": numa,
                    "serial_calls_value": total_frames * args.steps / e2e_serial_s,
                    "note": "cuboid_process_batch on pinned host depth, results to host; value = steps dealt round-robin over "
                            "%d handle(s), one host thread each (CUBOID_OPT_PIPELINE=0: chunk-wide launches inside a handle, the "
                            "handles overlap each other's copies); serial_calls_value = one handle with its internal sub-chunk "
                            "pipeline, one call after the other. `value` times ONE handle with its stages back to back on one "
                            "stream, so the overlapped multi-handle figure can come out slightly above it (ICP tails of one "
                            "batch are filled by the next batch's front end)" % n_handles},
            "gpu_launches": int(launches),
            # achieved = ALGORITHMIC flops (8*S*T per nearest-neighbour pass, the brute-force figure) / CUDA-event time.
            # The kernel returns brute force's exact answer but proves most pairs irrelevant with an exact AABB bound, so this
            # exceeds the FP32 peak; `executed_*` is what the FP32 pipe actually did.
            "roofline": {"kernel": "k_icp", "bound": "fp32", "achieved": effective, "peak": peak_unfused, "unit": "TFLOP/s",
                         "frac": effective / peak_unfused if peak_unfused else None, "traffic": None,
                         "peak_source": "un-fused FMUL+FADD micro-benchmark in this run (bit-exactness forbids FFMA); FFMA peak %.1f" % peak_ffma,
                         "launches_per_step": n_chunks, "ms_per_launch": 1e3 * icp_s / n_chunks,
                         "algorithmic_flops_per_step": ops, "executed_flops_per_step": ops_exec,
                         "executed_tflops": achieved, "executed_frac": achieved / peak_unfused if peak_unfused else None,
                         "culled_fraction": 1.0 - work_eval / max(work_brute, 1),
                         "note": "frac > 1 because exact culling (BVH + bit-exact lower bound, DESIGN.md §4) skips pairs that provably cannot "
                                 "win; results are bit-identical to the brute-force scan (CUBOID_OPT_ICP_CULL=0), see tests"},
            "roofline_hbm": {"kernel": "k_frontend" if fused else "k_preprocess", "bound": "hbm", "achieved": fe_gbs, "peak": peaks.get("hbm_gbs"),
                             "unit": "GB/s", "frac": fe_gbs / peaks.get("hbm_gbs") if peaks.get("hbm_gbs") else None, "traffic": traffic,
                             "peak_source": peak_src, "launches_per_step": n_chunks, "ms_per_launch": 1e3 * pre_s / n_chunks,
                             "algorithmic_bytes_per_step": fe_bytes,
                             "algorithmic_bytes_per_frame": "2*P + 16*N (unproject + passthrough) + 16*N + 16*V (voxel grid)" if fused
                             else "2*P + 16*N (unproject + passthrough)"},
            "stages_ms_per_step": {k: v / args.steps for k, v in stage.items()},
            "wall_ms_per_step": 1e3 * t_wall / args.steps,
            "frame_stats": {"mean_points": n_pts / F, "mean_voxels": n_vox / F, "mean_remain": sum(r.n_remain for r in res) / F,
                            "mean_icp_iterations": float(np.mean([r.cluster[0].iterations for r in res if r.n_clusters > 0] or [0])),
                            "accepted": int(sum(r.cluster[0].accepted for r in res if r.n_clusters > 0)),
                            "e2e_equals_device": bool(all(bytes(a) == bytes(b) for a, b in zip(res, res_e2e)))},
            "input_generation_s": t_gen,
        }
        if world == 1 and not args.no_cpu:
            from oracle import pyoracle as O
            op = O.params_from(p)
            ns = min(args.cpu_sample, F)
            otm = tm if stages & 8 else None
            O.process_frame(op, frames[0], otm, guesses=rots)
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
            line["cpu_baseline"] = {"value": ns / cdt, "unit": "frames/s", "cores": 1, "kind": "port",
                                    "sample": "first %d frames of the same batch, single thread (the reference node is one ros::spin thread)" % ns,
                                    "gpu_matches_oracle_on_sample": bool(same)}
        print(json.dumps(line), flush=True)
    cc.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
