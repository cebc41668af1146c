"""The drop-in boundary, executed: the reference's own UNMODIFIED test suite (tests/descriptor_test.cpp: 11 photographs x {256, 512}
bits x {BAD, HashSIFT} = 44 googletest cases) and its UNMODIFIED benchmark sample (samples/sample_benchmark.cpp), compiled by
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
BINARIES = ["ref_descriptor_test", "ref_sample_benchmark", "adapter_check"]


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
