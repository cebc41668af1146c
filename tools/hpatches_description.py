#!/usr/bin/env python3
"""HPatches descriptor dump over the B200 compute path -- the reference's samples/hpatches_description.cpp
 with the same options and the same output files:

    python tools/hpatches_description.py <hpatches-release> [--result-dir ./result] [--descriptor-type 0|1]
                                         [--descriptor-bits 256|512] [--compute-angle]

For every sequence directory the patch strips (*.png, 65 x 65 patches stacked vertically) are concatenated
horizontally (hpatches_description.cpp:217-222), one keypoint is placed at every patch centre with size 64 and
angle -1 (:233-241) or the intensity-centroid angle of the 65-pixel disc through cv::fastAtan2 (:107-162),
descriptors come from EfficientFeatures::compute with the vector<KeyPoint> arguments (:246-248; here
ef_compute_async behind efb200.EfficientFeatures.compute) and are written as one CSV of 0/1 per strip, most
significant bit first (:76-105), under <result-dir>/<BAD|HashSIFT>_<bits>/<sequence>/<strip>.csv.

Differences from the reference, on purpose: strips are taken in sorted order -- the reference concatenates them in the order
std::filesystem::directory_iterator yields (:55-63, unsorted), and with --compute-angle a rotated 64-pixel patch reaches into the
neighbouring strips, so its output depends on that file-system order; `--directory-order` reproduces it (os.listdir walks the same
readdir sequence; tests/test_hpatches_tool.py compares this tool with the reference's own unmodified sample that way).  `--synthetic S`
writes S small synthetic sequences into <hpatches-release> first (the dataset is not redistributable and absent here).

Image files are decoded with cv2 (the reference uses cv::imread); everything numeric on the descriptor side runs
in libef_b200.so on the GPU -- there is no CPU fallback."""
from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "cuda-efficient-features_b200"))

PATCH_SIZE = 65                      # hpatches_description.cpp:196
HALF_PATCH_SIZE = PATCH_SIZE // 2
DESC_STR = ("BAD", "HashSIFT")


def calc_umax(patch_size: int = PATCH_SIZE) -> np.ndarray:
    """End of every row of the circular patch (hpatches_description.cpp:107-126, the ORB construction)."""
    half = patch_size // 2
    umax = np.zeros(half + 2, np.int32)
    vmax = int(np.floor(np.float32(half) * np.sqrt(np.float32(2.0)) / 2 + 1))
    vmin = int(np.ceil(np.float32(half) * np.sqrt(np.float32(2.0)) / 2))
    for v in range(vmax + 1):
        umax[v] = int(np.rint(np.sqrt(float(half) * half - v * v)))     # cvRound = round half to even
    v0 = 0
    for v in range(half, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax


def ic_angles(img: np.ndarray, pts: np.ndarray, umax: np.ndarray, half_k: int = HALF_PATCH_SIZE) -> np.ndarray:
    """Intensity-centroid angle in degrees of every point (hpatches_description.cpp:128-162): exact integer moments
    over the disc, then cv::fastAtan2 (OpenCV's polynomial approximation -- cv2's own, not atan2)."""
    import cv2
    cx = np.floor(pts[:, 0]).astype(np.int64)
    cy = np.floor(pts[:, 1]).astype(np.int64)
    im = img.astype(np.int64)
    m01 = np.zeros(len(pts), np.int64)
    m10 = np.zeros(len(pts), np.int64)
    for u in range(-half_k, half_k + 1):
        m10 += u * im[cy, cx + u]
    for v in range(1, half_k + 1):
        d = int(umax[v])
        v_sum = np.zeros(len(pts), np.int64)
        for u in range(-d, d + 1):
            plus, minus = im[cy + v, cx + u], im[cy - v, cx + u]
            v_sum += plus - minus
            m10 += u * (plus + minus)
        m01 += v * v_sum
    return np.array([cv2.fastAtan2(float(np.float32(a)), float(np.float32(b))) for a, b in zip(m01, m10)], np.float32)


def patch_keypoints(nimages: int, npatches: int) -> np.ndarray:
    """n x 4 (x, y, size, angle): strip-major, patch centres, size 64, angle -1 (hpatches_description.cpp:233-241)."""
    x = np.float32(PATCH_SIZE) * (np.arange(nimages, dtype=np.float32) + np.float32(0.5))
    y = np.float32(PATCH_SIZE) * (np.arange(npatches, dtype=np.float32) + np.float32(0.5))
    k = np.empty((nimages, npatches, 4), np.float32)
    k[..., 0] = x[:, None]
    k[..., 1] = y[None, :]
    k[..., 2] = 64.0
    k[..., 3] = -1.0
    return k.reshape(-1, 4)


def descriptor_type(desc_type: int, bits: int) -> int:
    """sample_common.cpp:22-33."""
    import efb200
    if desc_type == 0:
        return efb200.BAD_256 if bits == 256 else efb200.BAD_512
    if desc_type == 1:
        return efb200.HASH_SIFT_256 if bits == 256 else efb200.HASH_SIFT_512
    return efb200.HASH_SIFT_256


def descriptor_csv(desc: np.ndarray) -> str:
    """One line per descriptor, bits most significant first, comma separated (hpatches_description.cpp:76-105)."""
    bits = np.unpackbits(desc, axis=1)       # big-endian within a byte = k = 7..0
    return "".join(",".join(map(str, row)) + "\n" for row in bits.tolist())


def load_sequence(seq_dir: Path, directory_order: bool = False):
    import cv2
    if directory_order:
        files = [seq_dir / n for n in os.listdir(seq_dir) if n.endswith(".png")]
    else:
        files = sorted(p for p in seq_dir.iterdir() if p.suffix == ".png")
    images = [cv2.imread(str(p), cv2.IMREAD_GRAYSCALE) for p in files]
    if any(i is None for i in images):
        raise RuntimeError(f"could not read every strip of {seq_dir}")
    return files, images


def describe_sequence(images, desc_type: int, bits: int, compute_angle: bool, feature=None):
    """-> (stacked image, keypoints n x 4, descriptors n x bits/8) for the strips of one sequence."""
    import efb200
    stacked = np.ascontiguousarray(np.hstack(images))
    npatches, nimages = stacked.shape[0] // PATCH_SIZE, len(images)
    kpts = patch_keypoints(nimages, npatches)
    if compute_angle:
        kpts[:, 3] = ic_angles(stacked, kpts, calc_umax())
    if feature is None:
        feature = efb200.EfficientFeatures.create(dtype=descriptor_type(desc_type, bits), max_width=stacked.shape[1],
                                                  max_height=stacked.shape[0], nfeatures=max(len(kpts), 1))
    desc = feature.compute(stacked, kpts)
    return stacked, kpts, np.asarray(desc)


def write_synthetic(root: Path, nseq: int, nimages: int = 4, npatches: int = 12, seed: int = 0xEFB2) -> None:
    """Synthetic stand-in for hpatches-release: smooth random blobs + noise, nimages strips of npatches patches each."""
    import cv2
    rng = np.random.default_rng(seed)
    for s in range(nseq):
        d = root / f"{'iv'[s % 2]}_synthetic{s:02d}"
        d.mkdir(parents=True, exist_ok=True)
        names = ["ref"] + [f"{c}{i}" for c in "eht" for i in range(1, 6)]
        for name in names[:nimages]:
            strip = rng.integers(0, 256, (PATCH_SIZE * npatches, PATCH_SIZE), dtype=np.uint8)
            strip = cv2.GaussianBlur(strip, (0, 0), 2.5)
            strip = cv2.normalize(strip, None, 0, 255, cv2.NORM_MINMAX)
            cv2.imwrite(str(d / f"{name}.png"), strip)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="HPatches descriptor dump (reference: samples/hpatches_description.cpp)")
    ap.add_argument("hpatches_dir", help="path to hpatches-release.")
    ap.add_argument("--result-dir", default="./result", help="path to result.")
    ap.add_argument("--descriptor-type", type=int, default=0, help="descriptor type(0:BAD 1:HashSIFT).")
    ap.add_argument("--descriptor-bits", type=int, default=256, help="descriptor bits(256 or 512).")
    ap.add_argument("--compute-angle", action="store_true", help="compute angles of keypoints.")
    ap.add_argument("--directory-order", action="store_true", help="concatenate the strips in directory (readdir) order like the reference instead of sorted")
    ap.add_argument("--synthetic", type=int, default=0, metavar="S", help="first write S synthetic sequences into hpatches_dir")
    a = ap.parse_args(argv)

    root = Path(a.hpatches_dir)
    if a.synthetic > 0:
        write_synthetic(root, a.synthetic)
    print("=== configulations ===")
    print(f"HPatchs directory : {root}")
    print(f"result directory  : {a.result_dir}")
    print(f"descriptor type   : {DESC_STR[a.descriptor_type]}")
    print(f"descriptor bits   : {a.descriptor_bits}")
    print(f"compute angle     : {'Yes' if a.compute_angle else 'No'}\n")
    if not root.exists():
        print(f"No such directory: {root}", file=sys.stderr)
        return 1
    seqs = sorted(p for p in root.iterdir() if p.is_dir())
    print(f"number of patch directories: {len(seqs)}")
    desc_dir = Path(a.result_dir) / f"{DESC_STR[a.descriptor_type]}_{a.descriptor_bits}"
    for count, seq in enumerate(seqs, 1):
        print(f"sequence: {count:3d}/{len(seqs):3d} [{seq.name}]")
        files, images = load_sequence(seq, a.directory_order)
        _, kpts, desc = describe_sequence(images, a.descriptor_type, a.descriptor_bits, a.compute_angle)
        npatches = len(kpts) // len(images)
        print(f"patch num: [{npatches} x {len(images)}]\n")
        out = desc_dir / seq.name
        os.makedirs(out, exist_ok=True)
        for x, f in enumerate(files):
            (out / (f.stem + ".csv")).write_text(descriptor_csv(desc[x * npatches:(x + 1) * npatches]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
