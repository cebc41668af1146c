#!/usr/bin/env python3
"""sample_benchmark with the reference's command line (samples/sample_benchmark.cpp:27-37,96-142): same options, same three
benchmark types, same timing protocol (one untimed iteration, mean wall time of the next N with a stream synchronisation after
every call) and the same two output lines -- so numbers compare one to one with the reference's README table.

  python tools/sample_benchmark.py image.png --max-keypoints=10000 --descriptor-type=1 --descriptor-bits=256
  python tools/sample_benchmark.py synthetic:3840x2160 --benchmark-type=1

The input is an image file (decoded with cv2 when it is importable; BGR -> gray on the device through ef_bgr_to_gray_async) or
`synthetic:WxH` (uniform noise; there is no image file in this repository)."""
import argparse
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "cuda-efficient-features_b200"))


def perf(niterations, fn):
    total = 0.0
    for it in range(niterations + 1):
        t0 = time.perf_counter()
        fn()
        t1 = time.perf_counter()
        if it > 0:
            total += t1 - t0
    return 1e3 * total / niterations


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("input_image")
    ap.add_argument("--max-keypoints", type=int, default=10000)
    ap.add_argument("--fast-threshold", type=int, default=20)
    ap.add_argument("--num-levels", type=int, default=8)
    ap.add_argument("--nonmax-radius", type=int, default=15)
    ap.add_argument("--descriptor-type", type=int, default=0, help="0:BAD 1:HashSIFT")
    ap.add_argument("--descriptor-bits", type=int, default=256, help="256 or 512")
    ap.add_argument("--benchmark-type", type=int, default=0, help="0:detect-and-compute 1:detect-only 2:compute-only")
    ap.add_argument("--num-iterations", type=int, default=100)
    a = ap.parse_args()

    import torch
    import efb200

    if a.input_image.startswith("synthetic:"):
        w, h = (int(v) for v in a.input_image.split(":")[1].lower().split("x"))
        g = torch.Generator(device="cpu").manual_seed(0xEFB2)
        d_gray = torch.randint(0, 256, (h, w), dtype=torch.uint8, generator=g).cuda()
    else:
        import cv2
        img = cv2.imread(a.input_image, cv2.IMREAD_UNCHANGED)
        if img is None:
            print("imread failed.")
            return 1
        d_gray = efb200.cvtColorToGray(torch.from_numpy(img).cuda())          # convertToGray, sample_common.cpp:35-45
    h, w = d_gray.shape
    # getDescriptorType, sample_common.cpp:25-33
    dtype = {(0, 256): efb200.BAD_256, (0, 512): efb200.BAD_512, (1, 256): efb200.HASH_SIFT_256, (1, 512): efb200.HASH_SIFT_512}.get(
        (a.descriptor_type, a.descriptor_bits), efb200.HASH_SIFT_256)
    feature = efb200.EfficientFeatures.create(a.max_keypoints, max_width=w, max_height=h, max_keypoints=a.max_keypoints)
    feature.setNLevels(a.num_levels)
    feature.setFastThreshold(a.fast_threshold)
    feature.setNonmaxRadius(a.nonmax_radius)
    feature.setDescriptorType(dtype)

    out = {}
    if a.benchmark_type == 0:
        def fn():
            out["kp"], out["desc"] = feature.detectAndComputeAsync(d_gray)    # sizes the outputs: one sync, like waitForCompletion
            torch.cuda.synchronize()
    elif a.benchmark_type == 1:
        def fn():
            out["kp"] = feature.detectAsync(d_gray)
            torch.cuda.synchronize()
    else:
        out["kp"] = feature.detectAsync(d_gray)
        torch.cuda.synchronize()
        def fn():
            out["desc"] = feature.computeAsync(d_gray, out["kp"])
            torch.cuda.synchronize()
    t = perf(a.num_iterations, fn)
    print("%5d keypoints found." % out["kp"].shape[1])
    print("processing time: %.1f[milli sec]" % t)
    return 0


if __name__ == "__main__":
    sys.exit(main())
