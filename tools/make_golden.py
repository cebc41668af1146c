#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ (run in the build container, needs
/root/reference and oracle/_ref):

  desc_golden.npz     a 320x240 crop of the reference's own test image tests/data/images/100_7100.JPG
                      (decoded once with cv2.IMREAD_GRAYSCALE, stored as pixels so JPEG decoding is not
                      part of the test), 400 stress keypoints, and the outputs of the REFERENCE's unmodified
                      CPU descriptors (oracle/_ref = bad.cpp + hash_sift.cpp): BAD-256/512 at scale 1 and 5,
                      HashSIFT 129-vectors, HashSIFT-256/512 bits.
  detect_golden.npz   oracle detector output on a synthetic 640x480 frame (pins the oracle against
                      regressions; the reference has no CPU detector and no golden vectors for it).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import efo  # noqa: E402


def main():
    import cv2
    efo.build()
    o, r = efo.Oracle(), efo.Reference()
    full = cv2.imread("/root/reference/tests/data/images/100_7100.JPG", cv2.IMREAD_GRAYSCALE)
    crop = np.ascontiguousarray(full[900:1140, 1200:1520])
    k = efo.stress_keypoints(crop.shape[1], crop.shape[0], 400, seed=42)
    out = {"image": crop, "keypoints": k}
    for nbits in (256, 512):
        for scale in (1.0, 5.0):
            out[f"bad{nbits}_s{int(scale)}"] = r.bad(crop, k, scale, nbits)
        out[f"hashsift{nbits}"] = r.hashsift(crop, k, 1.0, nbits)
    out["hashsift_features"] = r.hashsift_features(crop, k, 1.0)
    gold = ROOT / "tests" / "golden"
    gold.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(gold / "desc_golden.npz", **out)

    img = o.synth_frame(0xEFB20000 + 77, 0, 640, 480)
    kp, desc, counts = o.detect_and_compute(img, o.make_params(nfeatures=1500, desc_type=efo.BAD_256))
    np.savez_compressed(gold / "detect_golden.npz", seed=np.uint32(0xEFB20000 + 77), width=640, height=480, nfeatures=1500,
                        keypoints=kp, descriptors=desc, counts=counts, image_checksum=np.uint64(int(img.astype(np.uint64).sum())))
    print("wrote", [p.name for p in gold.iterdir()], "ref descriptors from oracle/_ref; n kp", len(kp))


if __name__ == "__main__":
    main()
