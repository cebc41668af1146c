"""The C++ host mirror (cuda-efficient-features_b200/cpp/ef_features.hpp) through the reference's sample_benchmark command line
(cpp/sample_benchmark.cpp): builds on the CPU box, fails loudly without a GPU, and on the GPU produces exactly what the Python
mirror produces for the same frame (both are thin layers over the same C ABI)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "cuda-efficient-features_b200"
EXE = PKG / "sample_benchmark"


def build():
    """compile the sample against the library that is already there (never relink libef_b200.so from inside a test process)"""
    src = PKG / "cpp" / "sample_benchmark.cpp"
    if EXE.exists() and EXE.stat().st_mtime >= max(src.stat().st_mtime, (PKG / "cpp" / "ef_features.hpp").stat().st_mtime):
        return
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", str(src), "-I/usr/local/cuda/include", f"-L{PKG}", "-lef_b200",
                           "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,$ORIGIN", "-o", str(EXE)])
    assert EXE.exists()


def synth(w, h, seed=0xEFB20000):
    """lowbias32 counter hash of SURVEY 8d (frame 0), as in sample_benchmark.cpp"""
    idx = (np.arange(h, dtype=np.uint32)[:, None] * np.uint32(w) + np.arange(w, dtype=np.uint32)[None, :]).astype(np.uint32)
    x = idx ^ np.uint32(seed)
    x ^= x >> 16; x = (x * np.uint32(0x7feb352d)).astype(np.uint32); x ^= x >> 15; x = (x * np.uint32(0x846ca68b)).astype(np.uint32); x ^= x >> 16
    return (x >> 24).astype(np.uint8)


def test_cpp_sample_builds_and_refuses_to_run_without_a_gpu():
    import torch
    build()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([str(EXE), "synthetic:640x480", "--num-iterations=1"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr
    assert subprocess.run([str(EXE), "--help"], capture_output=True).returncode == 1


@pytest.mark.gpu
@pytest.mark.parametrize("dtype_args,dtype_name,bench_type", [(["--descriptor-type=0", "--descriptor-bits=256"], "BAD_256", 0),
                                                               (["--descriptor-type=1", "--descriptor-bits=512"], "HASH_SIFT_512", 0),
                                                               (["--descriptor-type=0", "--descriptor-bits=512"], "BAD_512", 2)])
def test_cpp_sample_equals_python_mirror(tmp_path, oracle, dtype_args, dtype_name, bench_type):
    import torch
    import efb200
    build()
    w, h, nf = 1111, 777, 4000
    img = synth(w, h)
    assert np.array_equal(img, oracle.synth_frame(0xEFB20000, 0, w, h))        # the sample's generator is the survey's
    dump = tmp_path / "dump.txt"
    r = subprocess.run([str(EXE), f"synthetic:{w}x{h}", f"--max-keypoints={nf}", "--num-iterations=2", f"--benchmark-type={bench_type}", f"--dump={dump}"] + dtype_args,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = dump.read_text().splitlines()
    n, db = (int(v) for v in lines[0].split())
    assert f"{n:5d} keypoints found." in r.stdout and "processing time:" in r.stdout
    ef = efb200.EfficientFeatures.create(nf, dtype=getattr(efb200, dtype_name), max_width=w, max_height=h)
    d_img = torch.from_numpy(img).cuda()
    if bench_type == 0:
        kp, desc = ef.detectAndComputeAsync(d_img)
    else:
        kp = ef.detectAsync(d_img)
        desc = ef.computeAsync(d_img, kp)                                       # 5 x N GpuMat path: size forced to 31
    k = ef.convert(kp)
    d = desc.cpu().numpy()
    assert n == len(k) and db == d.shape[1]
    for i, line in enumerate(lines[1:]):
        f = line.split()
        assert (int(f[0]), int(f[1]), int(f[2])) == (int(k["x"][i]), int(k["y"][i]), int(k["octave"][i]))
        assert [int(v, 16) for v in f[3:6]] == [int(k[c][i:i + 1].view(np.uint32)[0]) for c in ("response", "angle", "size")]
        assert bytes.fromhex(f[6]) == d[i].tobytes()
