#!/usr/bin/env python3
"""bench.py -- detectAndCompute throughput on synthetic uniform-noise frames (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (sm_100a kernels through the C ABI)
  python bench.py --impl reference --gpus N ...            the reference's CPU path (oracle) on the host cores

A step = one detectAndCompute pass over one batch of `--batch` synthetic 4K frames per GPU (HashSIFT-512,
nfeatures 40000 by default = BASELINE.json configs[3]/[4] workload at the 4K resolution the metric is quoted on).
Frames are sharded over ranks (one process per GPU, no data-path collective): scaling = weak.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "cuda-efficient-features_b200"))

DESC = {"BAD_256": (0, 32), "BAD_512": (1, 64), "HASH_SIFT_256": (2, 32), "HASH_SIFT_512": (3, 64)}


def level_pixels(w, h, scale=1.2, nlevels=8):
    import numpy as np
    s = np.float32(1.0); tot = w * h
    for _ in range(1, nlevels):
        s = np.float32(s * np.float32(scale)); inv = np.float32(1.0) / s
        tot += int(np.rint(inv * np.float32(h))) * int(np.rint(inv * np.float32(w)))
    return tot


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag.is_set():
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, nme in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def cpu_reference_run(args, steps, warmup, frames_per_step=1, threads=None):
    """Times the CPU path (oracle = the reference's CPU descriptors + CPU restatement of its CUDA-only detector)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import efo
    o = efo.Oracle()
    threads = threads or o.max_threads()
    o.set_threads(threads)
    frames = [o.synth_frame(0xEFB20004, f, args.width, args.height) for f in range(frames_per_step)]
    params = o.make_params(nfeatures=args.nfeatures, desc_type=DESC[args.desc][0])
    nk = 0
    for _ in range(warmup):
        o.detect_and_compute(frames[0], params)
    t0 = time.perf_counter()
    for _ in range(steps):
        for f in frames:
            k, _, _ = o.detect_and_compute(f, params)
            nk += len(k)
    dt = time.perf_counter() - t0
    mpix = steps * frames_per_step * args.width * args.height / dt / 1e6
    return {"value": mpix, "seconds": dt, "threads": threads, "keypoints_per_s": nk / dt,
            "sample": f"{steps} step(s) x {frames_per_step} frame(s) of {args.width}x{args.height}, {args.desc}, nfeatures {args.nfeatures}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="frames per GPU per step")
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--nfeatures", type=int, default=40000)
    ap.add_argument("--desc", default="HASH_SIFT_512", choices=list(DESC))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--tiled", action="store_true",
                    help="instead of sharding frames, cut EVERY frame into horizontal bands over the ranks (ef_band_*): image broadcast from "
                         "rank 0, all-gather of the band candidates and MAX all-reduce of the descriptors inside the timed region; strong scaling")
    args = ap.parse_args()

    # NCCL's own log lines (e.g. "NCCL version ..." when the box sets NCCL_DEBUG) go to stderr: stdout carries the ONE JSON line
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "detectAndCompute_throughput_4K_40k"
    config = {"workload": f"detectAndCompute {args.desc} on {args.width}x{args.height} synthetic uniform-noise frames, nfeatures {args.nfeatures}, "
                          f"8 levels x1.2, FAST th 20, NMS r 15 (BASELINE.json configs[3]/[4] at the 4K metric resolution)",
              "width": args.width, "height": args.height, "nfeatures": args.nfeatures, "descriptor": args.desc,
              "frames_per_gpu_per_step": args.batch,
              "sharding": (f"every frame cut into {world} horizontal band(s) with NMS halo; NCCL broadcast of the image, all-gather of band candidates, "
                           f"MAX all-reduce of descriptors" if args.tiled else f"frames over {world} rank(s), no collective on the data path"),
              "l2_policy": f"batch of {args.batch} frames = {args.batch * args.width * args.height / 1e6:.0f} MB input + >1 GB intermediates per step, larger than the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(args, max(1, args.steps), max(0, min(args.warmup, 1)))
        line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic", "config": config, "keypoints_per_s": r["keypoints_per_s"],
                "cpu_baseline": {"value": r["value"], "unit": "Mpix/s", "cores": r["threads"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import efb200

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B, H, W = args.batch, args.height, args.width
    dtype_id, dbytes = DESC[args.desc]

    gen = torch.Generator(device="cpu").manual_seed(0xEFB2 + rank)
    frames_host = torch.randint(0, 256, (B, H, W), dtype=torch.uint8, generator=gen).pin_memory()
    frames = frames_host.to(dev)
    ef = efb200.EfficientFeatures.create(nfeatures=args.nfeatures, dtype=dtype_id, max_width=W, max_height=H, max_batch=B, device=local_rank)
    kp = torch.empty((B, 5, args.nfeatures), dtype=torch.float32, device=dev)
    desc = torch.empty((B, args.nfeatures, dbytes), dtype=torch.uint8, device=dev)
    counts = torch.zeros(B, dtype=torch.int32, device=dev)
    out = (kp, desc, counts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.tiled:
        from efb200 import tiling
        def step():
            tiling.detect_and_compute_tiled(ef, frames, src=0 if world > 1 else None, out=out)
    else:
        def step():
            ef.detectAndComputeBatchRaw(frames, out=out)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ef.stageTimingEnable(True)
    launches0 = ef.kernelLaunchCount()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ef.kernelLaunchCount() - launches0
    stage_ms, ncalls = ef.stageTimes()
    ef.stageTimingEnable(False)
    nk = int(counts.sum().item())
    t = torch.tensor([ms, float(nk)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms = float(tmax[0].item()); nk_total = float(tsum[1].item())
    else:
        nk_total = float(nk)
    if args.tiled:   # the same B frames on every rank: total work is fixed
        value = B * args.steps * W * H / (ms * 1e-3) / 1e6
        kps = float(nk) * args.steps / (ms * 1e-3)
    else:
        value = world * B * args.steps * W * H / (ms * 1e-3) / 1e6
        kps = nk_total * args.steps / (ms * 1e-3)

    # ---- e2e: same metric through the host-buffer C-ABI call (H2D of the frames + D2H of keypoints/descriptors inside)
    e2e = None
    if not args.no_e2e and not args.tiled:
        kp_h = torch.empty((B, 5, args.nfeatures), dtype=torch.float32).pin_memory()
        desc_h = torch.empty((B, args.nfeatures, dbytes), dtype=torch.uint8).pin_memory()
        for _ in range(min(args.warmup, 3)):
            cnt = ef.hostBatchInto(frames_host, kp_h, desc_h)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cnt = ef.hostBatchInto(frames_host, kp_h, desc_h)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0].item())
        e2e = {"value": world * B * args.steps * W * H / dt / 1e6, "unit": "Mpix/s", "h2d_bytes_per_step": B * H * W,
               "d2h_bytes_per_step": int(sum(cnt)) * (20 + dbytes) + 4 * B, "ms_per_step": 1e3 * dt / args.steps}

    clocks = sampler.finish() if sampler else None   # sampled through both timed regions (device-resident and host-buffer)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (algorithmic bytes per DESIGN.md section 5)
    P = level_pixels(W, H)
    n_frame = nk / B
    alg = {"pyramid": P, "score": P, "nms": 12 * n_frame, "compact": 12 * n_frame, "select": 12 * n_frame,
           "angle_pack": n_frame * (709 + 20), "blur": 2 * P, "describe": n_frame * (1024 + 16 + (dbytes if dtype_id < 2 else 128)),
           "project": n_frame * (128 + dbytes)}
    peak, peak_kind = measured_peak()
    per_call = {k: v / max(args.steps, 1) for k, v in stage_ms.items()}   # per step (a tiled step is two C-ABI calls)
    dom = max(per_call, key=per_call.get)
    def roof(stage):
        ach = alg[stage] * B / (per_call[stage] * 1e-3) / 1e9 if per_call[stage] > 0 else 0.0
        return {"kernel": stage, "bound": "hbm", "achieved": ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak,
                "traffic": None, "ms_per_launch_group": per_call[stage], "algorithmic_bytes_per_step": alg[stage] * B}
    # traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch group from the committed ncu capture (profiles/traffic.json,
    # written by tools/ncu_summary.py for a batch of `frames` frames), rescaled to this run's batch
    traffic_file = ROOT / "profiles" / "traffic.json"
    tr = {}
    if traffic_file.exists():
        try:
            tr = json.loads(traffic_file.read_text())
        except Exception:
            tr = {}
    def with_traffic(r):
        v = tr.get(("describe_bad" if (r["kernel"] == "describe" and dtype_id < 2) else r["kernel"]))
        if isinstance(v, (int, float)) and tr.get("_frames"):
            r["traffic"] = v * B / float(tr["_frames"])
        return r
    roofline = with_traffic(roof(dom))
    roofline_pyr = with_traffic(roof("pyramid"))
    rooflines = {st: with_traffic(roof(st)) for st in per_call if per_call[st] > 0}
    # What actually bounds these kernels: issue slots and the shared-memory pipe.  Executed warp instructions
    # and shared-memory wavefronts per launch group come from the committed ncu capture (profiles/traffic.json), the time from
    # this run; peak = 4 warp instructions resp. 1 wavefront per SM and clock at the SM clock sampled under load.
    stage_issue, stage_smem = {}, {}
    if clocks and clocks.get("sm_mhz") and tr.get("_frames"):
        n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
        clk_per_ms = n_sm * float(clocks["sm_mhz"]) * 1e3
        for st, t_ms in per_call.items():
            key = "describe_bad" if (st == "describe" and dtype_id < 2) else st
            scale = B / float(tr["_frames"])
            if t_ms > 0 and key in tr.get("_warp_inst", {}):
                stage_issue[st] = round(tr["_warp_inst"][key] * scale / (4.0 * clk_per_ms * t_ms), 4)
            if t_ms > 0 and key in tr.get("_smem_wavefronts", {}):
                stage_smem[st] = round(tr["_smem_wavefronts"][key] * scale / (clk_per_ms * t_ms), 4)

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args, steps=2, warmup=0)
        r1 = cpu_reference_run(args, steps=1, warmup=0, threads=1)
        cpu_baseline = {"value": r["value"], "unit": "Mpix/s", "cores": r["threads"], "kind": "port", "sample": r["sample"],
                        "single_thread_value": r1["value"], "keypoints_per_s": r["keypoints_per_s"]}

    line = {"metric": metric, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.tiled else "weak", "vs_baseline": None, "dtype": "u8/f32",
            "data": "synthetic", "config": config, "frames_per_s": value * 1e6 / (W * H), "keypoints_per_s": kps,
            "keypoints_per_frame": n_frame, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roofline, "roofline_pyramid": roofline_pyr, "stage_ms_per_step": per_call,
            "stage_roofline_frac": {k: round(v["frac"], 5) for k, v in rooflines.items()},
            "stage_issue_frac": stage_issue, "stage_smem_pipe_frac": stage_smem, "cpu_baseline": cpu_baseline}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
