#!/usr/bin/env python3
"""The reference's own CUDA detector kernels on this GPU next to ours, on the same pyramid of one synthetic 4K frame (and one 'natural-like'
frame with few corners).  Reference side: oracle/_ref/libef_ref_cuda.so (cuda_fast.cu + cuda_efficient_features.cu compiled unmodified),
replaying the per-level sequence of detectAndComputeAsync (cuda_efficient_features.cpp:244-272,310) with its 0.1*area candidate cap and its
two host synchronisations per level; cv::cuda::resize / Gaussian / descriptors are NOT included on either side.  Ours: the detector stages
(score + nms + compact + select + angle_pack) of ef_detect_and_compute_async, one frame per call.  Prints one JSON line."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "cuda-efficient-features_b200", ROOT / "oracle"):
    sys.path.insert(0, str(p))

import numpy as np
import torch

import efb200
import efo


def main():
    ref = efo.ReferenceCuda()
    o = efo.Oracle()
    w, h, nf = 3840, 2160, 40000
    out = []
    rng = np.random.default_rng(1)
    base = rng.integers(0, 256, (h // 8 + 1, w // 8 + 1), dtype=np.uint8)
    frames = {"noise": o.synth_frame(0xEFB20004, 0, w, h),
              "blocks8": np.kron(base, np.ones((8, 8), np.uint8))[:h, :w].copy()}       # ~natural corner density, no cap overflow
    for name, img in frames.items():
        ef = efb200.EfficientFeatures.create(nf, dtype=efb200.BAD_256, max_width=w, max_height=h)
        d = torch.from_numpy(img).cuda()
        for _ in range(3):
            ef.detectAsync(d)
        ef.stageTimingEnable(True)
        iters = 20
        for _ in range(iters):
            ef.detectAndComputeRaw(d, want_descriptors=False)
        torch.cuda.synchronize()
        st, ncalls = ef.stageTimes()
        ef.stageTimingEnable(False)
        ours = {k: v / ncalls for k, v in st.items() if v > 0}
        ours_detect = sum(v for k, v in ours.items() if k != "pyramid")
        kp = ef.detect(d)
        levels = [ef.debugLevelArrays(l, want=("image",))["image"] for l in range(8)]
        _, _, scales = o.level_geometry(w, h)
        quotas = o.level_quotas(nf)
        ms, n = ref.time_detect_levels(levels, scales, quotas, 20, 15.0, iters)
        out.append({"frame": name, "reference_kernels_ms": ms, "reference_keypoints": sum(n), "ours_detector_stages_ms": ours_detect,
                    "ours_keypoints": len(kp), "ours_stage_ms": ours, "speedup": ms / ours_detect})
    # ---- descriptors: the reference's GPU HashSIFT-512 vs ours (compute-only API) on the same 40 000 keypoints of the noise frame
    img = frames["noise"]
    det = efb200.EfficientFeatures.create(nf, dtype=efb200.BAD_256, max_width=w, max_height=h)
    kd = det.detect(torch.from_numpy(img).cuda())
    k = np.stack([kd["x"], kd["y"], np.full(len(kd), 31.0, np.float32), kd["angle"]], axis=1).astype(np.float32)
    k = np.concatenate([k, k[: 40000 - len(k)]])[:40000]
    ms_ref, desc_ref = ref.time_hashsift(img, k, 512, 1.0, 20)
    hs = efb200.HashSIFT.create(1.0, 100, max_width=w, max_height=h, max_keypoints=40000)
    dimg = torch.from_numpy(img).cuda(); dk = torch.from_numpy(k).cuda()
    desc = torch.empty((len(k), 64), dtype=torch.uint8, device="cuda")
    L, hnd = hs._ef._h.L, hs._ef._h.h
    # (a) std::vector<KeyPoint> path (arbitrary sizes / sub-pixel positions: generic kernel); (b) 5 x N GpuMat path, the one the reference's
    # sample_benchmark --benchmark-type=2 times (integer positions, size 31: window-staging kernel)
    rows = np.zeros((5, len(k)), np.float32)
    rows[0] = np.stack([k[:, 0].astype(np.int16), k[:, 1].astype(np.int16)], axis=1).copy().view(np.float32)[:, 0]
    rows[2] = k[:, 3]
    drows = torch.from_numpy(rows).cuda()
    def ours_vec():
        L.ef_compute_async(hnd, dimg.data_ptr(), dimg.stride(0), w, h, dk.data_ptr(), len(k), desc.data_ptr(), 64, efb200._stream_ptr(None))
    def ours_rows():
        L.ef_compute_rows_async(hnd, dimg.data_ptr(), dimg.stride(0), w, h, drows.data_ptr(), drows.stride(0) * 4, len(k), desc.data_ptr(), 64, efb200._stream_ptr(None))
    def timeit(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20
    ms_vec = timeit(ours_vec)
    ms_ours = timeit(ours_rows)
    cpu = o.hashsift(img, k, 1.0, 512)
    d = desc.cpu().numpy()
    bits = lambda a, b: int(np.unpackbits(a ^ b, axis=1).sum())
    hs_cmp = {"keypoints": len(k), "reference_gpu_ms": ms_ref, "ours_ms": ms_ours, "ours_vector_keypoint_path_ms": ms_vec, "speedup": ms_ref / ms_ours,
              "bits_differing_from_cpu_reference": {"reference_gpu": bits(desc_ref, cpu), "ours": bits(d, cpu), "of": int(cpu.size * 8)}}
    # ---- BAD-512: the reference's GPU kernel (integral image excluded on its side: cudev is third-party) vs ours (5 x N path, window integral included)
    ms_bref, bdesc_ref = ref.time_bad(img, k, 512, 1.0, 20)
    bad = efb200.BAD.create(1.0, 100, max_width=w, max_height=h, max_keypoints=40000)
    Lb, hb = bad._ef._h.L, bad._ef._h.h
    def ours_bad_rows():
        Lb.ef_compute_rows_async(hb, dimg.data_ptr(), dimg.stride(0), w, h, drows.data_ptr(), drows.stride(0) * 4, len(k), desc.data_ptr(), 64, efb200._stream_ptr(None))
    def ours_bad_vec():
        Lb.ef_compute_async(hb, dimg.data_ptr(), dimg.stride(0), w, h, dk.data_ptr(), len(k), desc.data_ptr(), 64, efb200._stream_ptr(None))
    ms_bvec = timeit(ours_bad_vec)
    ms_bours = timeit(ours_bad_rows)
    bcpu = o.bad(img, k, 1.0, 512)
    bd = desc.cpu().numpy()
    bad_cmp = {"keypoints": len(k), "reference_gpu_kernel_ms_without_integral": ms_bref, "ours_ms": ms_bours, "ours_vector_keypoint_path_ms_with_integral": ms_bvec,
               "speedup": ms_bref / ms_bours,
               "bits_differing_from_cpu_reference": {"reference_gpu": bits(bdesc_ref, bcpu), "ours": bits(bd, bcpu), "of": int(bcpu.size * 8)}}
    print(json.dumps({"detector_kernels_4k_one_frame": out, "hashsift512_compute_40k": hs_cmp, "bad512_compute_40k": bad_cmp}))


if __name__ == "__main__":
    main()
