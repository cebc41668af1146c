"""tools/hpatches_description.py (reference: samples/hpatches_description.cpp): host logic on the CPU, descriptors
against the oracle on the GPU."""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _tool():
    spec = importlib.util.spec_from_file_location("hpatches_description", ROOT / "tools" / "hpatches_description.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["hpatches_description"] = mod
    spec.loader.exec_module(mod)
    return mod


def test_umax_is_the_orb_disc():
    hp = _tool()
    u = hp.calc_umax(31)     # the ORB table for PATCH_SIZE 31 (also cuda_efficient_features.cu:143)
    assert u[:16].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    u65 = hp.calc_umax(65)
    assert u65[0] == 32 and u65[32] <= 8 and all(u65[v] >= u65[v + 1] for v in range(32))
    # symmetric disc: row v ends at u65[v]  <=>  column u ends at u65[u]
    inside = np.array([[abs(x) <= u65[abs(y)] for x in range(-32, 33)] for y in range(-32, 33)])
    assert np.array_equal(inside, inside.T)


def test_ic_angles_match_a_direct_moment_sum():
    cv2 = pytest.importorskip("cv2")
    hp = _tool()
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (65 * 3, 65 * 2), dtype=np.uint8)
    k = hp.patch_keypoints(2, 3)
    assert k.shape == (6, 4) and k[0].tolist() == [32.5, 32.5, 64.0, -1.0] and k[4].tolist() == [97.5, 97.5, 64.0, -1.0]
    umax = hp.calc_umax()
    got = hp.ic_angles(img, k, umax)
    for i, (x, y) in enumerate(k[:, :2]):
        cx, cy = int(np.floor(x)), int(np.floor(y))
        m10 = m01 = 0
        for v in range(-32, 33):
            d = int(umax[abs(v)])
            row = img[cy + v, cx - d:cx + d + 1].astype(np.int64)
            m10 += int((np.arange(-d, d + 1) * row).sum())
            m01 += v * int(row.sum())
        assert got[i] == np.float32(cv2.fastAtan2(float(m01), float(m10)))


def test_csv_is_msb_first():
    hp = _tool()
    txt = hp.descriptor_csv(np.array([[0x80, 0x01], [0xFF, 0x00]], np.uint8))
    assert txt == "1,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1\n1,1,1,1,1,1,1,1,0,0,0,0,0,0,0,0\n"


@pytest.mark.gpu
@pytest.mark.parametrize("desc_type,bits,angle", [(0, 256, False), (0, 512, True), (1, 256, True), (1, 512, False)])
def test_descriptors_equal_oracle(tmp_path, oracle, desc_type, bits, angle):
    pytest.importorskip("cv2")
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    hp = _tool()
    data, res = tmp_path / "hpatches", tmp_path / "result"
    args = [str(data), "--result-dir", str(res), "--descriptor-type", str(desc_type), "--descriptor-bits", str(bits), "--synthetic", "2"]
    assert hp.main(args + (["--compute-angle"] if angle else [])) == 0
    seqs = sorted(p for p in data.iterdir() if p.is_dir())
    assert len(seqs) == 2
    for seq in seqs:
        files, images = hp.load_sequence(seq)
        stacked = np.ascontiguousarray(np.hstack(images))
        k = hp.patch_keypoints(len(images), stacked.shape[0] // hp.PATCH_SIZE)
        if angle:
            k[:, 3] = hp.ic_angles(stacked, k, hp.calc_umax())
        want = oracle.bad(stacked, k, 1.0, bits) if desc_type == 0 else oracle.hashsift(stacked, k, 1.0, bits)
        npatches = len(k) // len(images)
        for x, f in enumerate(files):
            rows = (res / f"{hp.DESC_STR[desc_type]}_{bits}" / seq.name / (f.stem + ".csv")).read_text().strip().split("\n")
            got = np.packbits(np.array([[int(b) for b in r.split(",")] for r in rows], np.uint8), axis=1)
            assert np.array_equal(got, want[x * npatches:(x + 1) * npatches]), f"{seq.name}/{f.name}"


# ---- the reference's own UNMODIFIED samples/hpatches_description.cpp over the adapter (oracle/_ref/ref_hpatches_description)
REFBIN = ROOT / "oracle" / "_ref"


def test_standin_fastatan2_equals_cv2():
    """cv::fastAtan2 of the OpenCV stand-in (used by the reference sample's ICAngles) == cv2.fastAtan2, bit for bit"""
    cv2 = pytest.importorskip("cv2")
    import subprocess
    exe = REFBIN / "fastatan2_check"
    if not exe.exists():
        pytest.skip("oracle/_ref/fastatan2_check not built (make -C oracle adapter)")
    rng = np.random.default_rng(5)
    y = np.concatenate([rng.integers(-70000, 70000, 4000).astype(np.float32), np.array([0, 0, 1, -1, 5, -5, 0, 3e-9, 1e9], np.float32)])
    x = np.concatenate([rng.integers(-70000, 70000, 4000).astype(np.float32), np.array([0, 1, 0, 0, 5, 5, -2, 1e-9, -1e9], np.float32)])
    out = subprocess.run([str(exe)], input="".join(f"{float(a)!r} {float(b)!r}\n" for a, b in zip(y, x)), capture_output=True, text=True, timeout=60)
    got = np.array([int(l) for l in out.stdout.split()], np.uint32).view(np.float32)
    want = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    assert len(got) == len(want) and np.array_equal(got.view(np.uint32), want.view(np.uint32)), np.nonzero(got != want)[0][:5]


@pytest.mark.gpu
@pytest.mark.parametrize("desc_type,bits,angle", [(0, 256, True), (1, 512, False), (1, 256, True)])
def test_reference_hpatches_sample_equals_the_tool(tmp_path, desc_type, bits, angle):
    """the reference's sample (cv::imread stand-in reads "<strip>.png.pgm") and tools/hpatches_description.py --directory-order write the
    same CSV files"""
    cv2 = pytest.importorskip("cv2")
    import subprocess
    exe = REFBIN / "ref_hpatches_description"
    if not exe.exists():
        pytest.skip("oracle/_ref/ref_hpatches_description not built (needs /root/reference at build time)")
    hp = _tool()
    data, res_tool, res_ref = tmp_path / "hpatches", tmp_path / "result_tool", tmp_path / "result_ref"
    flags = ["--descriptor-type", str(desc_type), "--descriptor-bits", str(bits)] + (["--compute-angle"] if angle else [])
    hp.write_synthetic(data, 2)
    for png in list(data.rglob("*.png")):
        img = cv2.imread(str(png), cv2.IMREAD_GRAYSCALE)
        with open(str(png) + ".pgm", "wb") as f:
            f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0])); f.write(img.tobytes())
    out = subprocess.run([str(exe), str(data), f"--result-dir={res_ref}", f"--descriptor-type={desc_type}", f"--descriptor-bits={bits}"] +
                         (["--compute-angle"] if angle else []), capture_output=True, text=True, cwd=ROOT, timeout=600)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    # the reference concatenates the strips in readdir order, and a rotated 64-pixel patch reaches into the neighbouring strips
    assert hp.main([str(data), "--result-dir", str(res_tool), "--directory-order"] + flags) == 0
    csvs = sorted(p.relative_to(res_tool) for p in res_tool.rglob("*.csv"))
    assert len(csvs) == 8 and csvs == sorted(p.relative_to(res_ref) for p in res_ref.rglob("*.csv"))
    for c in csvs:
        assert (res_tool / c).read_text() == (res_ref / c).read_text(), str(c)
