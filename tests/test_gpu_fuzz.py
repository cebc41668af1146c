"""Randomised differential test: detectAndCompute through the C ABI vs the CPU oracle on seeded random configurations -- image size,
image statistics, pyramid shape, FAST threshold, NMS radius, keypoint budget, descriptor type, capacity larger than the frame, a batch
whose frames differ.  Bit-exact (keypoint sets, responses, angles, descriptor bytes).  `EF_FUZZ_CASES=N` widens the sweep
(tools/gpu_fuzz.sh); the default set stays within a couple of minutes of oracle time."""
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

N_CASES = int(os.environ.get("EF_FUZZ_CASES", "40"))


def make_image(rng, oracle, kind, w, h, seed):
    if kind == "noise":
        return oracle.synth_frame(seed, 0, w, h)
    if kind == "smooth":      # box-filtered noise: photograph-like corner density, long runs of near-ties
        a = rng.integers(0, 256, (h + 8, w + 8)).astype(np.float32)
        c = np.cumsum(np.cumsum(a, 0), 1)
        k = int(rng.integers(2, 6))
        s = c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]
        s = s[:h, :w] / (k * k)
        s = (s - s.min()) / max(float(s.max() - s.min()), 1.0) * 255.0
        return np.ascontiguousarray(s.astype(np.uint8))
    if kind == "blocks":      # random axis-aligned rectangles: exact corners, plateaus, many equal responses
        img = np.full((h, w), int(rng.integers(0, 256)), np.uint8)
        for _ in range(int(rng.integers(20, 200))):
            x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
            x1, y1 = min(w, x0 + int(rng.integers(2, 60))), min(h, y0 + int(rng.integers(2, 60)))
            img[y0:y1, x0:x1] = int(rng.integers(0, 256))
        return img
    # "sparse": a dark frame with a few bright dots and crosses (few corners: most levels select nothing)
    img = np.full((h, w), 16, np.uint8)
    for _ in range(int(rng.integers(1, 30))):
        x, y = int(rng.integers(2, w - 2)), int(rng.integers(2, h - 2))
        img[y - 1:y + 2, x] = 240
        img[y, x - 1:x + 2] = 240
    return img


def random_case(i):
    rng = np.random.default_rng(0xEFB2F000 + i)
    w, h = int(rng.integers(36, 720)), int(rng.integers(36, 560))
    if i % 7 == 0:
        w = (w + 15) & ~15          # 16-byte aligned rows: TMA loaders on level 0
    sf = float(rng.choice([1.1, 1.2, 1.25, 1.41, 1.5, 2.0]))
    nlevels = int(rng.integers(1, 11))
    while nlevels > 1 and min(w, h) / sf ** (nlevels - 1) < 2.0:   # every level must keep at least one pixel
        nlevels -= 1
    return dict(rng=rng, w=w, h=h, kind=str(rng.choice(["noise", "smooth", "blocks", "sparse"])),
                p=dict(nfeatures=int(rng.choice([1, 7, 100, 500, 1500, 4000])), scale_factor=sf, nlevels=nlevels, first_level=0,
                       fast_threshold=int(rng.choice([1, 5, 10, 20, 35, 60, 120])), nonmax_radius=int(rng.choice([0, 1, 3, 5, 8, 15, 16, 23, 40]))),
                dtype=str(rng.choice(["BAD_256", "BAD_512", "HASH_SIFT_256", "HASH_SIFT_512"])),
                slack_w=int(rng.choice([0, 0, 13, 64])), slack_h=int(rng.choice([0, 0, 9])), batch=int(rng.choice([1, 1, 2, 3])))


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.mark.parametrize("case", range(N_CASES))
def test_random_configuration_matches_oracle(torch_cuda, oracle, case):
    import efb200, efo
    torch = torch_cuda
    c = random_case(case)
    rng, w, h, p = c["rng"], c["w"], c["h"], c["p"]
    imgs = [make_image(rng, oracle, c["kind"] if b == 0 else str(rng.choice(["noise", "smooth", "blocks", "sparse"])), w, h,
                       util.SEED + 1000 + 7 * case + b) for b in range(c["batch"])]
    ef = efb200.EfficientFeatures.create(nfeatures=p["nfeatures"], scaleFactor=p["scale_factor"], nlevels=p["nlevels"], firstLevel=p["first_level"],
                                         fastThreshold=p["fast_threshold"], nonmaxRadius=p["nonmax_radius"], dtype=getattr(efb200, c["dtype"]),
                                         max_width=w + c["slack_w"], max_height=h + c["slack_h"], max_batch=c["batch"])
    frames = torch.from_numpy(np.stack(imgs)).cuda()
    kp, desc, cnt = ef.detectAndComputeBatchRaw(frames)
    torch.cuda.synchronize()
    op = oracle.make_params(desc_type=getattr(efo, c["dtype"]), **p)
    for b, img in enumerate(imgs):
        ok, od, _ = oracle.detect_and_compute(img, op)
        n = int(cnt[b])
        g = ef.convert(kp[b][:, :n].contiguous())
        o = util.oracle_to_struct(ok)
        util.assert_keypoints_equal(g, o)
        if n:
            _, go = util.canon_keypoints(g)
            _, oo = util.canon_keypoints(o)
            assert np.array_equal(desc[b, :n].cpu().numpy()[go], od[oo]), f"case {case} frame {b}: descriptor bytes differ ({c['dtype']}, {w}x{h}, {p})"


N_COMPUTE = int(os.environ.get("EF_FUZZ_COMPUTE_CASES", "24"))


@pytest.mark.parametrize("case", range(N_COMPUTE))
def test_random_compute_only_matches_oracle(torch_cuda, oracle, case):
    """compute-only API (vector<KeyPoint> path) on random frames, describer scales and keypoint sets: positions up to the last pixel,
    sizes 2..300 (boxes and patches far outside the frame), the special angles of efo.stress_keypoints."""
    import efb200, efo
    rng = np.random.default_rng(0xEFB2C000 + case)
    w, h = int(rng.integers(40, 900)), int(rng.integers(40, 700))
    img = make_image(rng, oracle, str(rng.choice(["noise", "smooth", "blocks"])), w, h, util.SEED + 5000 + case)
    n = int(rng.choice([1, 33, 1000, 3000]))
    k = efo.stress_keypoints(w, h, n, seed=100 + case)
    k[:, 2] = np.where(rng.random(n) < 0.3, k[:, 2], rng.uniform(2, 300, n)).astype(np.float32)
    scale = float(rng.choice([0.5, 1.0, 1.0, 2.0, 5.0, 6.75]))
    nbits = int(rng.choice([256, 512]))
    if case % 2 == 0:
        g = efb200.BAD.create(scale, 100 if nbits == 512 else 101, max_width=w, max_height=h).compute(img, k)
        o = oracle.bad(img, k, scale, nbits)
    else:
        g = efb200.HashSIFT.create(scale, 100 if nbits == 512 else 101, max_width=w, max_height=h).compute(img, k)
        o = oracle.hashsift(img, k, scale, nbits)
    bad = (g != o).any(axis=1)
    assert not bad.any(), f"case {case}: {bad.sum()} of {n} descriptors differ ({w}x{h}, scale {scale}, {nbits} bits), first keypoint {k[np.nonzero(bad)[0][0]]}"


N_MATCH = int(os.environ.get("EF_FUZZ_MATCH_CASES", "16"))


@pytest.mark.parametrize("case", range(N_MATCH))
def test_random_matcher_and_ingest_match_oracle(torch_cuda, case):
    """Hamming matcher (knn k = 2, best match, cross-check, ratio + cross-check filter) and BGR(A) -> gray on random shapes: query / train
    counts from 1 to a few thousand (tile remainders of the tcgen05 kernel, fewer train rows than k), 256- and 512-bit rows, a small byte
    alphabet for masses of ties, duplicated rows."""
    import efb200
    import match_oracle as mo
    torch = torch_cuda
    rng = np.random.default_rng(0xEFB2A000 + case)
    nq, nt = int(rng.choice([1, 2, 31, 129, 500, 1025, 3000])), int(rng.choice([1, 2, 33, 127, 128, 777, 2049, 4000]))
    nbytes, lo = int(rng.choice([32, 64])), int(rng.choice([2, 4, 256]))
    q = rng.integers(0, lo, (nq, nbytes), dtype=np.uint8)
    t = rng.integers(0, lo, (nt, nbytes), dtype=np.uint8)
    if nt > 4 and case % 3 == 0:
        t[rng.integers(0, nt, nt // 4)] = t[rng.integers(0, nt, nt // 4)]          # duplicated train rows: equal distances, lowest index wins
    if nq > 4 and nt > 4 and case % 4 == 1:
        q[: min(nq, nt) // 2] = t[: min(nq, nt) // 2]                               # exact matches (distance 0)
    dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    bf = efb200.BFMatcher.create()
    i12, d12 = bf.knnMatchAsync(dq, dt, 2); i21, d21 = bf.knnMatchAsync(dt, dq, 2)
    oi12, od12 = mo.knn_match(q, t, 2); oi21, od21 = mo.knn_match(t, q, 2)
    assert np.array_equal(i12.cpu().numpy(), oi12), f"case {case}: knn indices differ ({nq} x {nt} x {nbytes})"
    v12, v21 = oi12 >= 0, oi21 >= 0
    assert np.array_equal(d12.cpu().numpy()[v12], od12[v12])
    assert np.array_equal(i21.cpu().numpy(), oi21) and np.array_equal(d21.cpu().numpy()[v21], od21[v21])
    ci, cd = efb200.BFMatcher.create(efb200.NORM_HAMMING, True).matchAsync(dq, dt)
    oci, ocd = mo.cross_check_match(q, t)
    assert np.array_equal(ci.cpu().numpy(), oci) and np.array_equal(cd.cpu().numpy()[oci >= 0], ocd[oci >= 0])
    if nq >= 2 and nt >= 2:
        f = efb200.ratio_cross_filter(i12, d12, i21, d21, 0.9)
        assert np.array_equal(f.cpu().numpy(), mo.ratio_cross_filter(oi12, od12, oi21, od21, 0.9))
    h, w, cn = int(rng.integers(1, 300)), int(rng.integers(1, 500)), int(rng.choice([3, 4]))
    img = rng.integers(0, 256, (h, w, cn), dtype=np.uint8)
    assert np.array_equal(efb200.cvtColorToGray(torch.from_numpy(img).cuda()).cpu().numpy(), mo.bgr_to_gray(img))
