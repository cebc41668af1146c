"""The drop-in boundary, executed: the reference's own UNMODIFIED test suite (tests/descriptor_test.cpp: 11 photographs x {256, 512}
bits x {BAD, HashSIFT} = 44 googletest cases) and its four UNMODIFIED samples (samples/sample_benchmark.cpp, sample_feature_extraction.cpp,
sample_feature_matching.cpp, sample_image_sequence.cpp -- the last two call cv::BFMatcher, which the stand-in routes to ef_match_*), compiled by
`make -C oracle adapter` against the reference's unmodified public headers with cuda-efficient-features_b200/cpp/opencv_adapter.cpp +
libef_b200.so standing where modules/cuda_efficient_features stood (OpenCV itself: the stand-in under oracle/shim).  Plus
oracle/adapter_check.cpp: argument kinds, asserts, setters, describer classes through the same headers; its dumps are compared with
the Python host mirror byte for byte.

The binaries are built in the build container (they need /root/reference) and travel to the GPU box under oracle/_ref/."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref"
IMG_DIR = ROOT / "tests" / "golden" / "images"
BINARIES = ["ref_descriptor_test", "ref_sample_benchmark", "ref_sample_feature_extraction", "ref_sample_feature_matching", "ref_sample_image_sequence", "adapter_check"]


def need(name):
    p = REF / name
    if not p.exists():
        pytest.skip(f"{p} not built (needs /root/reference at build time: make -C oracle adapter)")
    return str(p)


def write_pgm(path, img):
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img, np.uint8).tobytes())


def sidecars():
    """cv::imread stand-in (oracle/shim/opencv2/highgui.hpp) reads "<file>.pgm": the pixels cv2 decodes from the JPEG"""
    import cv2
    for i in range(11):
        src = IMG_DIR / f"100_71{i:02d}.JPG"
        dst = Path(str(src) + ".pgm")
        if not dst.exists():
            write_pgm(dst, cv2.imread(str(src), cv2.IMREAD_GRAYSCALE))


def read_dump(path):
    raw = np.fromfile(path, np.uint8)
    rows, cols, elem = raw[:12].view(np.int32)
    return raw[12:].reshape(rows, cols * elem)


# ---- CPU: the binaries exist, link against libef_b200.so and hold the reference's 44 cases (no GPU needed to list them)
def test_adapter_binaries_built_and_list_the_reference_tests():
    exe = need("ref_descriptor_test")
    out = subprocess.run([exe, "--gtest_list_tests"], capture_output=True, text=True, cwd=ROOT, timeout=60)
    assert out.returncode == 0, out.stderr
    cases = [l for l in out.stdout.splitlines() if l.startswith("  ")]
    assert len(cases) == 44 and sum(c.strip().startswith("BAD/") for c in cases) == 22 and sum(c.strip().startswith("HashSIFT/") for c in cases) == 22
    ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libef_b200.so" in ldd and "not found" not in ldd.split("libef_b200.so")[1].splitlines()[0], ldd
    helptext = subprocess.run([need("ref_sample_benchmark"), "--help"], capture_output=True, text=True, timeout=60).stdout
    for opt in ("max-keypoints", "fast-threshold", "num-levels", "nonmax-radius", "descriptor-type", "descriptor-bits", "benchmark-type", "num-iterations"):
        assert opt in helptext
    nm = subprocess.run(["nm", "-C", "--defined-only", str(REF / "opencv_adapter.o")], capture_output=True, text=True).stdout
    for sym in ("cv::cuda::EfficientFeatures::create(", "cv::cuda::EfficientFeatures::~EfficientFeatures()", "cv::cuda::BAD::create(",
                "cv::cuda::HashSIFT::create(", "cv::cuda::EfficientDescriptorsAsync::~EfficientDescriptorsAsync()"):
        assert sym in nm, f"the adapter does not define {sym}"


# ---- GPU
@pytest.mark.gpu
def test_reference_descriptor_test_suite_passes_unmodified():
    """all 44 cases of the reference's own googletest binary"""
    exe = need("ref_descriptor_test")
    sidecars()
    out = subprocess.run([exe, "--gtest_brief=1"], capture_output=True, text=True, cwd=ROOT, timeout=1500)
    tail = "\n".join(out.stdout.splitlines()[-15:])
    print(tail)
    assert out.returncode == 0, tail + out.stderr[-2000:]
    assert "[  PASSED  ] 44 tests." in out.stdout, tail


@pytest.mark.gpu
@pytest.mark.parametrize("bench_type", [0, 1, 2])
def test_reference_sample_benchmark_runs_unmodified(bench_type):
    """samples/sample_benchmark.cpp:104-141: the three modes through *Async + GpuMat + Stream::waitForCompletion"""
    exe = need("ref_sample_benchmark")
    sidecars()
    img = str(IMG_DIR / "100_7100.JPG")
    out = subprocess.run([exe, img, "--max-keypoints=40000", "--descriptor-type=1", "--descriptor-bits=512", f"--benchmark-type={bench_type}",
                          "--num-iterations=20"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "image size      : [2832 x 2128]" in out.stdout and "keypoints found." in out.stdout and "processing time:" in out.stdout
    nk = int(out.stdout.split("keypoints found.")[0].split()[-1])
    assert 5000 < nk <= 40000


@pytest.mark.gpu
def test_adapter_check_and_python_mirror_agree(tmp_path, oracle):
    """oracle/adapter_check.cpp through the reference's headers; its dumped keypoints / descriptors == efb200 (ctypes) == oracle"""
    import torch
    import efb200, efo
    import util
    exe = need("adapter_check")
    sidecars()
    pgm = str(IMG_DIR / "100_7103.JPG.pgm")
    prefix = str(tmp_path / "dump")
    out = subprocess.run([exe, pgm, prefix, "20000"], capture_output=True, text=True, cwd=ROOT, timeout=900)
    print(out.stdout)
    assert out.returncode == 0 and ", 0 failed" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
    import cv2
    img = cv2.imread(str(IMG_DIR / "100_7103.JPG"), cv2.IMREAD_GRAYSCALE)
    d_img = torch.from_numpy(img).cuda()
    for name, dtype in (("bad256", "BAD_256"), ("bad512", "BAD_512"), ("hashsift256", "HASH_SIFT_256"), ("hashsift512", "HASH_SIFT_512")):
        k = read_dump(prefix + f".{name}.kpts").view(np.float32)
        d = read_dump(prefix + f".{name}.desc")
        ef = efb200.EfficientFeatures.create(nfeatures=20000, dtype=getattr(efb200, dtype), max_width=img.shape[1], max_height=img.shape[0])
        kp, desc = ef.detectAndComputeAsync(d_img)
        assert np.array_equal(kp.cpu().numpy().view(np.uint32), k.view(np.uint32)), f"{name}: keypoints differ between the C++ adapter and the Python mirror"
        assert np.array_equal(desc.cpu().numpy(), d), f"{name}: descriptors differ between the C++ adapter and the Python mirror"
        ok, od, _ = oracle.detect_and_compute(img, oracle.make_params(nfeatures=20000, desc_type=getattr(efo, dtype)))
        g, o = efb200.EfficientFeatures.convert(k), util.oracle_to_struct(ok)
        util.assert_keypoints_equal(g, o)
        _, go = util.canon_keypoints(g)
        _, oo = util.canon_keypoints(o)
        assert np.array_equal(d[go], od[oo]), f"{name}: adapter descriptors differ from the oracle"
        del ef


# ---- the reference's other three samples, unmodified, against the Python host mirror on the same photographs
def _gray(i):
    import cv2
    return cv2.imread(str(IMG_DIR / f"100_71{i:02d}.JPG"), cv2.IMREAD_GRAYSCALE)


def _features(ef, img):
    import torch
    kp, desc = ef.detectAndComputeAsync(torch.from_numpy(img).cuda())
    return kp, desc


@pytest.mark.gpu
@pytest.mark.parametrize("async_flag", [False, True])
def test_reference_sample_feature_extraction_runs_unmodified(async_flag):
    """samples/sample_feature_extraction.cpp: Feature2D::detectAndCompute on a cv::Mat and the *Async + convert + download form"""
    import efb200
    exe = need("ref_sample_feature_extraction")
    sidecars()
    cmd = [exe, str(IMG_DIR / "100_7104.JPG"), "--max-keypoints=7000", "--descriptor-type=1", "--descriptor-bits=256"] + (["--compute-async"] if async_flag else [])
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=600)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    img = _gray(4)
    ef = efb200.EfficientFeatures.create(nfeatures=7000, dtype=efb200.HASH_SIFT_256, max_width=img.shape[1], max_height=img.shape[0])
    kp, _ = _features(ef, img)
    assert f"{kp.shape[1]} keypoints found." in out.stdout
    assert "[imshow] keypoints" in out.stdout and ("compute async   : " + ("Yes" if async_flag else "No")) in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("desc", [(0, 256, "BAD_256"), (1, 512, "HASH_SIFT_512")])
def test_reference_sample_feature_matching_runs_unmodified(desc):
    """samples/sample_feature_matching.cpp: two photographs, BFMatcher::create(NORM_HAMMING, true)->match == ef_match_cross_check_async"""
    import efb200
    exe = need("ref_sample_feature_matching")
    sidecars()
    out = subprocess.run([exe, str(IMG_DIR / "100_7101.JPG"), str(IMG_DIR / "100_7102.JPG"), "--max-keypoints=8000",
                          f"--descriptor-type={desc[0]}", f"--descriptor-bits={desc[1]}"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    a, b = _gray(1), _gray(2)
    ef = efb200.EfficientFeatures.create(nfeatures=8000, dtype=getattr(efb200, desc[2]), max_width=a.shape[1], max_height=a.shape[0])
    (k1, d1), (k2, d2) = _features(ef, a), _features(ef, b)
    ci, _ = efb200.BFMatcher.create(efb200.NORM_HAMMING, True).matchAsync(d1.contiguous(), d2.contiguous())
    assert f"number of keypoins: {k1.shape[1]} {k2.shape[1]}" in out.stdout
    assert f"number of matches: {int((ci >= 0).sum())}" in out.stdout
    assert int((ci >= 0).sum()) > 500        # overlapping views of the same castle


@pytest.mark.gpu
def test_reference_sample_image_sequence_runs_unmodified():
    """samples/sample_image_sequence.cpp over photographs 1..10: *Async + convert + download per frame, knnMatch k = 2 both ways and the
    sample's own ratio / cross-check loop; its per-frame match count (printed by the putText stand-in) == ef_match_ratio_cross_async"""
    import efb200
    exe = need("ref_sample_image_sequence")
    sidecars()
    out = subprocess.run([exe, str(IMG_DIR / "100_71%02d.JPG"), "--max-keypoints=6000", "--descriptor-type=1", "--descriptor-bits=256"],
                         capture_output=True, text=True, cwd=ROOT, timeout=900)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    got = [int(l.split(":")[1]) for l in out.stdout.splitlines() if l.startswith("[putText] number of matches")]
    assert len(got) == 9 and "imread failed." in out.stdout          # frames 2..10 are matched against their predecessor; frame 11 does not exist
    img = _gray(1)
    ef = efb200.EfficientFeatures.create(nfeatures=6000, dtype=efb200.HASH_SIFT_256, max_width=img.shape[1], max_height=img.shape[0])
    bf = efb200.BFMatcher.create()
    want, prev = [], None
    for i in range(1, 11):
        _, d = _features(ef, _gray(i))
        d = d.contiguous().clone()
        if prev is not None:
            i12, d12 = bf.knnMatchAsync(prev, d, 2); i21, d21 = bf.knnMatchAsync(d, prev, 2)
            want.append(int((efb200.ratio_cross_filter(i12, d12, i21, d21, 0.9) >= 0).sum()))
        prev = d
    assert got == want, (got, want)
