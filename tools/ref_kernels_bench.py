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
    print(json.dumps({"detector_kernels_4k_one_frame": out}))


if __name__ == "__main__":
    main()
