#!/usr/bin/env python3
"""Natural-image throughput next to the numbers the reference publishes (README.md:47-71): the reference's OWN benchmark CLI
(samples/sample_benchmark.cpp compiled unmodified against cpp/opencv_adapter.cpp + libef_b200.so: oracle/_ref/ref_sample_benchmark)
on the 11 SceauxCastle photographs of its test suite, at the native 2832 x 2128 and as 3840 x 2160 frames, in the three modes of
samples/sample_benchmark.cpp:113-142 (detect with the sample's default 10 000 keypoints like README.md:54; compute and
detectAndCompute with 40 000 requested keypoints and the four descriptor types like README.md:62,70).  Wall clock around
*Async + Stream::waitForCompletion, 1 discarded iteration, mean of --num-iterations -- the reference's own protocol.

Beside it, on the same photographs: the reference's own CUDA kernels compiled unmodified (oracle/_ref/libef_ref_cuda.so) --
the per-level detector sequence (calcKeypoints .. scalePoints; cv::cuda::resize / Gaussian excluded, they are third-party and
absent here) and its HashSIFT-512 / BAD-512 kernels on the detector's keypoints.

    python tools/natural_bench.py [--iterations 50] > profiles/r02/natural_images.json      (GPU box, from the repository root)
"""
import argparse
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "cuda-efficient-features_b200", ROOT / "oracle", ROOT / "tests"):
    sys.path.insert(0, str(p))

import numpy as np


def write_pgm(path, img):
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img, np.uint8).tobytes())


def upscale_4k(img):
    ys = (np.arange(2160, dtype=np.int64) * img.shape[0]) // 2160
    xs = (np.arange(3840, dtype=np.int64) * img.shape[1]) // 3840
    return np.ascontiguousarray(img[ys][:, xs])


def run_cli(exe, path, nkp, dtype, bits, mode, iters):
    out = subprocess.run([exe, path, f"--max-keypoints={nkp}", f"--descriptor-type={dtype}", f"--descriptor-bits={bits}",
                          f"--benchmark-type={mode}", f"--num-iterations={iters}"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    if out.returncode != 0:
        raise RuntimeError(out.stdout + out.stderr)
    ms = float(out.stdout.split("processing time:")[1].split("[")[0])
    nk = int(out.stdout.split("keypoints found.")[0].split()[-1])
    return ms, nk


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=50)
    ap.add_argument("--images", type=int, default=11)
    a = ap.parse_args()
    import cv2
    import torch
    import efb200, efo
    exe = str(ROOT / "oracle" / "_ref" / "ref_sample_benchmark")
    tmp = Path("/tmp/ef_natural")   # scratch on the box (gpurun_out/ is size-limited)
    tmp.mkdir(parents=True, exist_ok=True)
    ref = efo.ReferenceCuda() if efo.ReferenceCuda.available() else None
    o = efo.Oracle()
    names = [f"100_71{i:02d}.JPG" for i in range(a.images)]
    res = {"protocol": "samples/sample_benchmark.cpp unmodified over cpp/opencv_adapter.cpp; mean over the images; ms per call",
           "iterations": a.iterations, "gpu": torch.cuda.get_device_name(0), "sizes": {}}
    for size in ("2832x2128", "3840x2160"):
        rows = {"detect 10000": [], "keypoints (detect 10000)": []}
        refrows = {"reference detector kernels (40000)": [], "ours detector stages (40000)": [], "reference HashSIFT-512 kernels": [],
                   "reference BAD-512 kernel (integral excluded)": [], "keypoints": []}
        for name in names:
            img = cv2.imread(str(ROOT / "tests" / "golden" / "images" / name), cv2.IMREAD_GRAYSCALE)
            if size == "3840x2160":
                img = upscale_4k(img)
            pgm = str(tmp / (name + "." + size + ".pgm"))
            write_pgm(pgm, img)
            ms, nk = run_cli(exe, pgm, 10000, 0, 256, 1, a.iterations)          # README.md:54 "detect (default params)"
            rows["detect 10000"].append(ms); rows["keypoints (detect 10000)"].append(nk)
            for dt, dname in ((0, "BAD"), (1, "HashSIFT")):
                for bits in (256, 512):
                    for mode, mname in ((2, "compute"), (0, "detectAndCompute")):
                        ms, nk = run_cli(exe, pgm, 40000, dt, bits, mode, a.iterations)
                        rows.setdefault(f"{mname} 40000 {dname}{bits}", []).append(ms)
                        rows.setdefault("keypoints (40000 requested)", []).append(nk)
            if ref is not None:
                h, w = img.shape
                ef = efb200.EfficientFeatures.create(40000, dtype=efb200.BAD_256, max_width=w, max_height=h)
                d = torch.from_numpy(img).cuda()
                for _ in range(3):
                    ef.detectAsync(d)
                ef.stageTimingEnable(True)
                for _ in range(20):
                    ef.detectAndComputeRaw(d, want_descriptors=False)
                torch.cuda.synchronize()
                st, nc = ef.stageTimes()
                ef.stageTimingEnable(False)
                ours_detect = sum(v for k, v in st.items() if k != "pyramid") / nc
                kd = ef.detect(d)
                levels = [ef.debugLevelArrays(l, want=("image",))["image"] for l in range(8)]
                _, _, scales = o.level_geometry(w, h)
                ms_ref, n_ref = ref.time_detect_levels(levels, scales, o.level_quotas(40000), 20, 15.0, 20)
                k = np.stack([kd["x"], kd["y"], np.full(len(kd), 31.0, np.float32), kd["angle"]], axis=1).astype(np.float32)
                ms_hs, _ = ref.time_hashsift(img, k, 512, 1.0, 20)
                ms_bad, _ = ref.time_bad(img, k, 512, 1.0, 20)
                refrows["reference detector kernels (40000)"].append(ms_ref); refrows["ours detector stages (40000)"].append(ours_detect)
                refrows["reference HashSIFT-512 kernels"].append(ms_hs); refrows["reference BAD-512 kernel (integral excluded)"].append(ms_bad)
                refrows["keypoints"].append(len(kd))
                del ef
        res["sizes"][size] = {"ours_through_reference_cli_ms": {k: float(np.mean(v)) for k, v in rows.items()},
                              "reference_kernels_same_gpu_ms": {k: float(np.mean(v)) for k, v in refrows.items() if v}}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
