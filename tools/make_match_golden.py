#!/usr/bin/env python3
"""Golden vectors for the matcher / colour-conversion oracle, generated with the real OpenCV (cv2) in the build container:
   python tools/make_match_golden.py  ->  tests/golden/match_golden.npz
cv::BFMatcher and cv::cvtColor are third-party code the reference's samples call (sample_image_sequence.cpp:115-116,
sample_feature_matching.cpp:99-101, sample_common.cpp:39-42); these fixtures pin oracle/match_oracle.py to them."""
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def dm(ms):
    return np.array([[m.queryIdx, m.trainIdx, int(m.distance)] for m in ms], np.int32).reshape(-1, 3)


def main():
    rng = np.random.default_rng(0xEFB2)
    out = {"cv2_version": np.array(cv2.__version__)}
    cases = {"ties32": (rng.integers(0, 4, (90, 32), dtype=np.uint8), rng.integers(0, 4, (130, 32), dtype=np.uint8)),
             "rand64": (rng.integers(0, 256, (120, 64), dtype=np.uint8), rng.integers(0, 256, (75, 64), dtype=np.uint8)),
             "dups64": None, "one_train": (rng.integers(0, 256, (9, 32), dtype=np.uint8), rng.integers(0, 256, (1, 32), dtype=np.uint8))}
    base = rng.integers(0, 256, (40, 64), dtype=np.uint8)
    cases["dups64"] = (np.concatenate([base, base[:10]]), np.concatenate([base[5:30], base[5:30], base[:3]]))
    for name, (q, t) in cases.items():
        knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
        idx = np.full((len(q), 2), -1, np.int32); dist = np.full((len(q), 2), 2**31 - 1, np.int32)
        for i, ms in enumerate(knn):
            for j, m in enumerate(ms):
                idx[i, j] = m.trainIdx; dist[i, j] = int(m.distance)
        out[f"{name}_q"] = q; out[f"{name}_t"] = t
        out[f"{name}_knn_idx"] = idx; out[f"{name}_knn_dist"] = dist
        out[f"{name}_cross"] = dm(cv2.BFMatcher(cv2.NORM_HAMMING, True).match(q, t))
    bgr = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    bgra = rng.integers(0, 256, (21, 19, 4), dtype=np.uint8)
    out["bgr"] = bgr; out["bgr_gray"] = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    out["bgra"] = bgra; out["bgra_gray"] = cv2.cvtColor(bgra, cv2.COLOR_BGRA2GRAY)
    np.savez_compressed(ROOT / "tests" / "golden" / "match_golden.npz", **out)
    print("wrote tests/golden/match_golden.npz with", len(out), "arrays; cv2", cv2.__version__)


if __name__ == "__main__":
    main()
