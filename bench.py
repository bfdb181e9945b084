#!/usr/bin/env python
"""bench.py -- SLAM frames/sec on synthetic RGB-D of Replica shape (1200x680), BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W [--mode train|recon] [--impl reference]

A "step" is one local-optimisation cycle of the reference's SLAMTrainCams loop (reference slam/slam_pipeline.cpp:52-173
with configs/release/replica/office0.yaml): `local_opt_interval` = 10 frames of TSDF fusion (allocate + integrate +
expected depth + raycast + ICP maps per frame) followed, in train mode, by the window/keyframe free-view raycasts,
the Gaussian spawn, `local_opt_iters` = 20 optimiser iterations (GES forward, L1, backward, Adam) and the prune.
value = frames / device time of the K timed steps (max over ranks), inputs resident in HBM.
e2e   = the same through the host-buffer C-ABI calls (pinned host frames, H2D per frame, pose/loss D2H per step).

--impl reference times the reference's own InfiniTAM CPU engine (work_mode=recon, oracle/_ref/libitm_ref_fast.so,
all host threads) on a bounded sample of the same frames.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FRAMES_PER_STEP = 10  # local_opt_interval (office0.yaml:52)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=None, choices=["train", "recon"])
    ap.add_argument("--track", type=int, default=0, help="0: ground-truth poses (use_gt_pose=true, office0 config); 1: extended ICP tracker; 2: icp")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-timing", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--timing-reps", type=int, default=20)
    ap.add_argument("--ref-frames", type=int, default=0, help="frames per step for --impl reference (0 = auto)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.lines = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append(line.strip())

    def stop(self, t0=None, t1=None):
        """t0, t1: time.time() bounds of the timed region; the sampler itself is started before the warm-up steps (nvidia-smi needs
        ~0.2 s to deliver its first line, longer than a timed region of a few steps) and only the samples inside [t0, t1] are kept"""
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        import datetime
        rows = []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), f[5:9]))
            except ValueError:
                continue
        inside = [r for r in rows if t0 is None or (t0 - 0.02 <= r[0] <= t1 + 0.02)]
        window = "timed region"
        if not inside:
            inside, window = rows, "warm-up + timed region (no sample fell inside the timed region)"
        sm, mx, reasons = [], [], set()
        for _, a, b, fl in inside:
            sm.append(a)
            mx.append(b)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), fl):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of each kernel from the newest committed
    `ncu --set full` capture (profiles/*_avg.json, written by tools/ncu_summary.py); {} when none is committed"""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_*_avg.json")))
    if not files:
        return {}, None
    with open(files[-1]) as f:
        k = json.load(f)["kernels"]
    return {name.split("::")[-1].split("<")[0]: v["dram_bytes_per_launch"] for name, v in k.items()}, os.path.basename(files[-1])


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------
def make_frames(n_frames, device):
    """synthetic Replica-shaped sequence, generated on the GPU: poses [n,4,4], rgba u8 [n,H,W,4], depth i16 [n,H,W]"""
    import torch
    from gps_slam_b200 import synthetic as syn
    intr = syn.intrinsics("replica")
    poses = syn.trajectory(n_frames)
    rgba = torch.empty((n_frames, intr["height"], intr["width"], 4), dtype=torch.uint8, device=device)
    depth = torch.empty((n_frames, intr["height"], intr["width"]), dtype=torch.int16, device=device)
    for i in range(n_frames):
        r, d = syn.render_frame(poses[i], intr, device=device)
        rgba[i], depth[i] = r, d
    return intr, poses, rgba, depth


def run_reference(args):
    """reference arm: the reference's InfiniTAM CPU engine, work_mode=recon, all host threads (rank 0 only)"""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import numpy as np
    import torch
    from gps_slam_b200 import synthetic as syn
    from oracle import itm_ref
    kind = "fast" if itm_ref.available("fast") else "exact"
    if not itm_ref.available(kind):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libitm_ref_*.so not built (needs /root/reference at build time)"}))
        return
    cores = os.cpu_count() or 1
    intr = syn.intrinsics("replica")
    fps_step = args.ref_frames or 2           # bounded sample: a CPU frame costs ~0.1-1 s
    n = (args.warmup + args.steps) * fps_step
    poses = syn.trajectory(n)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    frames = [tuple(t.cpu().numpy() for t in syn.render_frame(poses[i], intr, device=dev)) for i in range(n)]
    ref = itm_ref.ItmRef(intr, tracker=0, threads=cores, kind=kind)
    k = 0
    for _ in range(args.warmup * fps_step):
        ref.process_frame(frames[k][0], frames[k][1], syn.c2w_to_colmajor(poses[k]))
        k += 1
    t0 = time.perf_counter()
    for _ in range(args.steps * fps_step):
        ref.process_frame(frames[k][0], frames[k][1], syn.c2w_to_colmajor(poses[k]))
        k += 1
    dt = time.perf_counter() - t0
    ref.close()
    fps = args.steps * fps_step / dt
    sample = "%d frames/step x %d steps of the same synthetic Replica-shaped sequence, ITMBasicEngine CPU (%s build), use_gt_pose" % (
        fps_step, args.steps, kind)
    print(json.dumps({
        "impl": "reference", "metric": "slam_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "InfiniTAM CPU engine work_mode=recon (TSDF only), Replica 1200x680 synthetic", "frames_per_step": fps_step},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def eval_psnr(pipe, intr, poses, rgba, n_frames, dev, every=10):
    """renderEvalImgs on every `every`-th training camera: PSNR = 20 log10(1 / sqrt(mse)) of the composited render against the
    frame (scripts/utils/image_utils.py psnr), and of the TSDF colour raycast alone for context.  Not timed."""
    import torch
    H, W = intr["height"], intr["width"]
    rgb = torch.empty((H, W, 3), device=dev)
    depth = torch.empty((H, W), device=dev)
    alpha = torch.empty((H, W), device=dev)
    ps, ps_tsdf = [], []
    with torch.cuda.stream(pipe.stream):
        for i in range(0, n_frames, every):
            base = pipe.render_eval(poses[i], rgb, depth, alpha)
            gt = rgba[i][..., :3].float() / 255.0
            mse = float(((rgb.clamp(0, 1) - gt) ** 2).mean())
            mse_t = float(((base - gt) ** 2).mean())
            ps.append(20.0 * math.log10(1.0 / math.sqrt(mse)))
            ps_tsdf.append(20.0 * math.log10(1.0 / math.sqrt(mse_t)))
    return {"psnr_db": sum(ps) / len(ps), "psnr_tsdf_only_db": sum(ps_tsdf) / len(ps_tsdf), "cameras": len(ps)}


def cpu_baseline(intr, poses, rgba, depth, n_frames=6):
    import numpy as np
    from gps_slam_b200 import synthetic as syn
    from oracle import itm_ref
    kind = "fast" if itm_ref.available("fast") else "exact"
    if not itm_ref.available(kind):
        return None
    cores = os.cpu_count() or 1
    ref = itm_ref.ItmRef(intr, tracker=0, threads=cores, kind=kind)
    fr = [(rgba[i].cpu().numpy(), depth[i].cpu().numpy()) for i in range(n_frames + 1)]
    ref.process_frame(fr[0][0], fr[0][1], syn.c2w_to_colmajor(poses[0]))
    t0 = time.perf_counter()
    for i in range(1, n_frames + 1):
        ref.process_frame(fr[i][0], fr[i][1], syn.c2w_to_colmajor(poses[i]))
    dt = time.perf_counter() - t0
    ref.close()
    return {"value": n_frames / dt, "unit": "frames/s", "cores": cores, "kind": "reference",
            "sample": "%d frames (after 1 warm-up) of the bench sequence through the reference InfiniTAM CPU engine "
                      "(work_mode=recon: fusion + raycast per frame, use_gt_pose), %s build, OMP threads = cores" % (n_frames, kind)}


# ---------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    from gps_slam_b200 import engine as E
    from gps_slam_b200 import synthetic as syn

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gps_slam_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    E.load_library()
    from gps_slam_b200 import slam
    mode = args.mode or slam.DEFAULT_MODE

    n_frames = (args.warmup + args.steps) * FRAMES_PER_STEP
    intr, poses, rgba, depth = make_frames(n_frames, dev)
    rgba_h = torch.empty(rgba.shape, dtype=rgba.dtype, pin_memory=True).copy_(rgba)
    depth_h = torch.empty(depth.shape, dtype=depth.dtype, pin_memory=True).copy_(depth)
    stream = torch.cuda.Stream(device=dev, priority=int(os.environ.get("GSB_MAIN_STREAM_PRIORITY", "-1")))
    pipe = slam.SlamPipeline(intr, mode=mode, device=local, stream=stream, rank=rank, world=world, use_gt_pose=args.track == 0,
                             tracker=args.track or 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_leg(resident):
        """W warm-up + K timed steps from a fresh map; returns (ms, launches, stats)"""
        pipe.reset()
        f = 0
        sampler = ClockSampler(local)
        sampler.start()
        with torch.cuda.stream(stream):
            for _ in range(args.warmup):
                for _ in range(FRAMES_PER_STEP):
                    pipe.process_frame(f, rgba if resident else rgba_h, depth if resident else depth_h, poses, resident)
                    f += 1
                pipe.end_of_step(resident)
            barrier()
            t_begin = time.time()
            l0 = E.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            marks = [e0]
            for _ in range(args.steps):
                for _ in range(FRAMES_PER_STEP):
                    pipe.process_frame(f, rgba if resident else rgba_h, depth if resident else depth_h, poses, resident)
                    f += 1
                pipe.end_of_step(resident)
                marks.append(torch.cuda.Event(enable_timing=True))
                marks[-1].record(stream)
            e1 = marks[-1]
            barrier()
            t_end = time.time()
            ms = e0.elapsed_time(e1)
            run_leg.per_step_ms = [round(marks[i].elapsed_time(marks[i + 1]), 3) for i in range(len(marks) - 1)]
            launches = E.launch_count() - l0
            clocks = sampler.stop(t_begin, t_end)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, clocks

    ms, launches, clocks = run_leg(True)
    per_step_ms = list(run_leg.per_step_ms)
    stats = pipe.stats()
    psnr = eval_psnr(pipe, intr, poses, rgba, n_frames, dev) if mode == "train" else None
    ms_e2e = run_leg(False)[0] if not args.no_e2e else float("nan")
    frames = args.steps * FRAMES_PER_STEP
    fps = frames / (ms * 1e-3)
    fps_e2e = frames / (ms_e2e * 1e-3)
    h2d, d2h = pipe.io_bytes_per_step(FRAMES_PER_STEP)

    roofline = None
    if not args.no_kernel_timing and rank == 0:
        peak, peak_src = load_peaks()
        torch.cuda.nvtx.range_push("kernel_timing")   # ncu --nvtx --nvtx-include "kernel_timing/" captures steady-state launches
        roofline = pipe.time_dominant_kernel(stream, peak, reps=args.timing_reps)
        torch.cuda.nvtx.range_pop()
        roofline["peak_source"] = peak_src
        traffic, src = ncu_traffic()
        for r in (roofline, roofline.get("tsdf_integrate") or {}):
            if r.get("kernel") in traffic:
                r["traffic"] = traffic[r["kernel"]]
                r["traffic_source"] = "profiles/" + src
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(intr, poses, rgba, depth)

    if rank == 0:
        out = {
            "metric": "slam_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": pipe.scaling(), "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict({"workload": slam.workload_name(mode) if args.track == 0 else slam.workload_name(mode).replace(
                                "use_gt_pose=true", "use_gt_pose=false: online ICP tracking, %s tracker" % ("extended" if args.track == 1 else "icp")),
                            "parallelism": "single GPU" if world == 1 else "Gaussians sharded by spatial block over %d GPUs, one [H,W,5] "
                                           "all-reduce per optimiser iteration; TSDF replicated" % world, "frames_per_step": FRAMES_PER_STEP, "width": intr["width"], "height": intr["height"],
                            "l2": "no explicit flush: every frame is new input (4.9 MB) and each step streams the visible voxel "
                                  "blocks 10x (V x 8 KB per frame), working set > 126 MB L2", "quality": psnr, "ms_per_step_each": per_step_ms}, **stats),
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(out))
    pipe.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
