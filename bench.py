#!/usr/bin/env python3
"""bench.py -- detectAndCompute throughput on synthetic uniform-noise frames (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (sm_100a kernels through the C ABI)
  python bench.py --impl reference --gpus N ...            the reference's CPU path (oracle) on the host cores

A step = one detectAndCompute pass over one batch of `--batch` synthetic 4K frames per GPU (HashSIFT-512,
nfeatures 40000 by default = BASELINE.json configs[3]/[4] workload at the 4K resolution the metric is quoted on).
Frames are sharded over ranks (one process per GPU, no data-path collective): scaling = weak.
Both arms read the SAME pixels: frame f of the workload is pix(f, y, x) = lowbias32(seed ^ ((f * H + y) * W + x)) >> 24 with
seed 0xEFB20004, generated on the device for the product arm (ef_synth_frames_async) and by the oracle for the CPU arm.
The line also carries one driver-run number for every other BASELINE.json config under "configs" (detect-only 4K, compute-only on
40 000 keypoints, 8K detectAndCompute, BAD-512 on 64 frames split over the ranks = strong scaling, and -- at N > 1 -- two 8K
frames cut into bands over the ranks).  Prints ONE JSON line (rank 0).
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
SEED = 0xEFB20004            # seed = 0xEFB20000 + config id; one seed for every bench workload, frames are numbered globally

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


def kernel_source_sha():
    """sha256 over the kernel sources: profiles/traffic.json records it, so a capture taken from other kernels is never reported"""
    import hashlib
    h = hashlib.sha256()
    for f in sorted((ROOT / "cuda-efficient-features_b200" / "csrc").glob("*.cu*")):
        h.update(f.name.encode()); h.update(f.read_bytes())
    return h.hexdigest()[:16]


def pin_host_to_gpu(local_rank, local_world):
    """Pin this rank's host threads (and, by first touch + preferred-node policy, its pinned staging buffers) to the CPUs of the
    GPU's NUMA node, sharing the node's CPUs evenly between the ranks that sit on it.  Returns what was done (for the JSON line)."""
    info = {"numa_node": None, "cpus": None, "pinned": False}
    try:
        import torch
        def node_of(i):
            pr = torch.cuda.get_device_properties(i)
            path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/numa_node"
            try:
                return int(open(path).read().strip())
            except Exception:
                return -1
        def parse(cl):
            out = []
            for part in cl.strip().split(","):
                if "-" in part:
                    a, b = part.split("-"); out += list(range(int(a), int(b) + 1))
                elif part:
                    out.append(int(part))
            return out
        nodes = [node_of(i) for i in range(local_world)]
        node = nodes[local_rank]
        allowed = sorted(os.sched_getaffinity(0))
        cpus = allowed
        if node >= 0:
            try:
                on_node = set(parse(open(f"/sys/devices/system/node/node{node}/cpulist").read()))
                cpus = [c for c in allowed if c in on_node] or allowed
            except Exception:
                pass
        peers = [r for r in range(local_world) if nodes[r] == node]
        k, m = peers.index(local_rank), len(peers)
        per = max(1, len(cpus) // m)
        mine = cpus[k * per:(k + 1) * per] if (k + 1) * per <= len(cpus) else cpus[-per:]
        os.sched_setaffinity(0, set(mine))
        if node >= 0:
            try:    # set_mempolicy(MPOL_PREFERRED, node): host staging buffers land on the GPU's node
                import ctypes
                libc = ctypes.CDLL(None, use_errno=True)
                mask = ctypes.c_ulong(1 << node)
                libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))
            except Exception:
                pass
        info = {"numa_node": node, "cpus": f"{mine[0]}-{mine[-1]}" if mine else None, "ncpus": len(mine), "pinned": True}
    except Exception as e:  # never fatal: affinity is an optimisation
        info["error"] = str(e)[:80]
    return info


def cpu_reference_run(args, steps, warmup, frames_per_step=1, threads=None):
    """Times the CPU path (oracle = the reference's CPU descriptors + CPU restatement of its CUDA-only detector)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import efo
    o = efo.Oracle()
    threads = threads or o.max_threads()
    o.set_threads(threads)
    frames = [o.synth_frame(SEED, f, args.width, args.height) for f in range(frames_per_step)]   # frame 0.. of the product arm's rank 0
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
    ap.add_argument("--batch", type=int, default=32, help="frames per GPU per step")
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--nfeatures", type=int, default=40000)
    ap.add_argument("--desc", default="HASH_SIFT_512", choices=list(DESC))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-config numbers under \"configs\"")
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

    host = pin_host_to_gpu(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", str(world))))
    # the pinned counter-based generator, on the device; rank r owns frames [r B, (r + 1) B) of the global sequence (tiled: all ranks the same)
    frames = efb200.synth_frames(B, H, W, SEED, first_frame=0 if args.tiled else rank * B, device=dev)
    frames_host = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
    frames_host.copy_(frames)
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

    # ---- one driver-run number for every other BASELINE.json config (all ranks take part: the barriers and max-over-ranks apply)
    def timed(fn, steps, warmup=2):
        for _ in range(warmup):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        t_ms = a.elapsed_time(b)
        if world > 1:
            tt = torch.tensor([t_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_ms = float(tt[0].item())
        return t_ms / steps

    extras = {}
    if not args.no_extras and not args.tiled and (W, H) == (3840, 2160):
        from efb200.sharding import shard_range
        # configs[1]: detect-only (FAST pyramid + radius-15 NMS + top-K + angles), same frames
        t1 = timed(lambda: ef.detectAndComputeBatchRaw(frames, want_descriptors=False, out=(kp, None, counts)), 5)
        extras["configs[1] detect-only 4K nfeatures 40000 r 15"] = {
            "ms_per_step": t1, "frames_per_step": world * B, "Mpix_per_s": world * B * W * H / (t1 * 1e-3) / 1e6, "frames_per_s": world * B / (t1 * 1e-3), "scaling": "weak"}
        # configs[2]: compute-only on 40 000 precomputed keypoints (detector with r = 7 fills every quota at 4K), through the 5 x N
        # GpuMat path = benchmark-type 2 of samples/sample_benchmark.cpp:129-139; each rank describes its own frame
        ef2 = efb200.EfficientFeatures.create(nfeatures=40000, nonmaxRadius=7, dtype=efb200.BAD_512, max_width=W, max_height=H, max_keypoints=40000, device=local_rank)
        k5 = ef2.detectAsync(frames[0]).contiguous()
        c2 = {"keypoints": int(k5.shape[1])}
        for name, dt_id in (("BAD_512", efb200.BAD_512), ("HASH_SIFT_512", efb200.HASH_SIFT_512)):
            ef2.setDescriptorType(dt_id)
            t2 = timed(lambda: ef2.computeAsync(frames[0], k5), 20)
            c2[name] = {"ms_per_call": t2, "Mkeypoints_per_s": world * int(k5.shape[1]) / (t2 * 1e-3) / 1e6}
        extras["configs[2] compute-only 40000 keypoints 4K"] = c2
        del ef2
        # configs[3]: 7680x4320 HashSIFT-512 (every quota binds: 40 000 delivered keypoints per frame)
        B8 = 4
        ef8 = efb200.EfficientFeatures.create(nfeatures=40000, dtype=efb200.HASH_SIFT_512, max_width=7680, max_height=4320, max_batch=B8, device=local_rank)
        frames8 = efb200.synth_frames(B8, 4320, 7680, SEED, first_frame=100000 + rank * B8, device=dev)
        out8 = (kp[:B8], desc[:B8], counts[:B8])
        t3 = timed(lambda: ef8.detectAndComputeBatchRaw(frames8, out=out8), 5)
        extras["configs[3] detectAndCompute HASH_SIFT_512 8K"] = {
            "ms_per_step": t3, "frames_per_step": world * B8, "Mpix_per_s": world * B8 * 7680 * 4320 / (t3 * 1e-3) / 1e6,
            "frames_per_s": world * B8 / (t3 * 1e-3), "keypoints_per_frame": float(counts[:B8].float().mean().item()), "scaling": "weak"}
        # N > 1: the same two 8K frames cut into horizontal bands over the ranks (ef_band_*): strong scaling of ONE oversized frame
        if world > 1:
            from efb200 import tiling
            fr2 = efb200.synth_frames(2, 4320, 7680, SEED, first_frame=100000, device=dev)
            out2 = (kp[:2], desc[:2], counts[:2])
            ef8.stageTimingEnable(True)
            t5 = timed(lambda: tiling.detect_and_compute_tiled(ef8, fr2, src=0, out=out2), 5)
            st8, _ = ef8.stageTimes()            # kernels of this rank (warm-up steps included: 7 steps); the rest of a step is collectives + gaps
            ef8.stageTimingEnable(False)
            extras["oversized frame: 2 x 8K HASH_SIFT_512 cut into bands"] = {
                "ms_per_step": t5, "frames_per_step": 2, "Mpix_per_s": 2 * 7680 * 4320 / (t5 * 1e-3) / 1e6, "bands": world, "scaling": "strong",
                "single_gpu_ms_per_step": t3 * 2 / B8, "rank0_kernel_ms_per_step": {k: round(v / 7, 4) for k, v in st8.items() if v > 0},
                "collectives": tiling.COLLECTIVES}
            del fr2
        del ef8, frames8
        # configs[4]: BAD-512, ONE batch of 64 4K frames split over the ranks (strong scaling), sub-batches of <= B frames per call
        a64, b64 = shard_range(64, rank, world)
        n64 = b64 - a64
        f64 = efb200.synth_frames(n64, H, W, SEED, first_frame=200000 + a64, device=dev) if n64 else None
        ef.setDescriptorType(efb200.BAD_512)
        def run64():
            for o in range(0, n64, B):
                n = min(B, n64 - o)
                ef.detectAndComputeBatchRaw(f64[o:o + n], out=(kp[:n], desc[:n], counts[:n]))
        t4 = timed(run64, 3)
        ef.setDescriptorType(dtype_id)
        extras["configs[4] BAD_512 batch of 64 4K frames split over the ranks"] = {
            "ms_per_batch": t4, "frames": 64, "frames_per_rank": n64, "Mpix_per_s": 64 * W * H / (t4 * 1e-3) / 1e6, "frames_per_s": 64 / (t4 * 1e-3), "scaling": "strong"}
        del f64

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
    traffic_note = "no capture"
    if traffic_file.exists():
        try:
            tr = json.loads(traffic_file.read_text())
        except Exception:
            tr = {}
        # a capture is only reported for the kernels it was taken from (tools/ncu_summary.py records the hash of csrc/)
        if tr.get("_src_sha") != kernel_source_sha():
            traffic_note = f"capture {tr.get('_src_sha')} is stale for kernels {kernel_source_sha()}: traffic not reported"
            tr = {}
        else:
            traffic_note = tr.get("_note", "")
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
            "stage_issue_frac": stage_issue, "stage_smem_pipe_frac": stage_smem, "traffic_source": traffic_note,
            "host": host, "configs": extras, "cpu_baseline": cpu_baseline}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
