"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs (bit-exact for every integer/byte stage; the only floating-point outputs -- Harris response,
angle, size -- are compared bit-for-bit too, tolerance 0 ULP)."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def make_ef(**kw):
    import efb200
    return efb200.EfficientFeatures.create(**kw)


# ---------------------------------------------------------------------------------------------------
# stage-by-stage: pyramid, blur, response map, per-level counts
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("w,h,seed", [(640, 480, 1), (1111, 625, 2), (1543, 868, 3), (333, 257, 4)])
def test_stages_match_oracle(torch_cuda, oracle, w, h, seed):
    import efb200, efo
    torch = torch_cuda
    img = oracle.synth_frame(util.SEED + seed, 0, w, h)
    ef = make_ef(nfeatures=2000, dtype=efb200.BAD_256, max_width=w, max_height=h)
    d_img = torch.from_numpy(img).cuda()
    kp, desc, count = ef.detectAndComputeRaw(d_img)
    torch.cuda.synchronize()
    nlevels = ef.getNLevels()
    pyr = oracle.pyramid(img, 1.2, nlevels, blurred=False)
    bpyr = oracle.pyramid(img, 1.2, nlevels, blurred=True)
    counts = ef.debugLevelCounts()
    params = oracle.make_params(nfeatures=2000, desc_type=efo.BAD_256)
    _, ocounts = oracle.detect(img, params)
    for l in range(nlevels):
        a = ef.debugLevelArrays(l)
        assert (a["width"], a["height"]) == (pyr[l].shape[1], pyr[l].shape[0])
        assert np.array_equal(a["image"], pyr[l]), f"pyramid level {l} differs in {(a['image'] != pyr[l]).sum()} px"
        assert np.array_equal(a["blurred"], bpyr[l]), f"blur level {l} differs in {(a['blurred'] != bpyr[l]).sum()} px"
        resp, ncorner = oracle.score_map(pyr[l], 20)
        assert np.array_equal(np.isfinite(a["response"]), np.isfinite(resp)), f"FAST corner set differs at level {l}"
        m = np.isfinite(resp)
        assert np.array_equal(a["response"][m].view(np.uint32), resp[m].view(np.uint32)), f"Harris response differs at level {l}"
        assert counts[l, 0] == ncorner
    assert np.array_equal(counts, ocounts), f"per-level counts differ:\n{counts}\n{ocounts}"


# ---------------------------------------------------------------------------------------------------
# whole detector + descriptors vs oracle
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype_name", ["BAD_256", "BAD_512", "HASH_SIFT_256", "HASH_SIFT_512"])
@pytest.mark.parametrize("w,h,nfeat,seed", [(1920, 1080, 5000, 1), (800, 600, 1500, 7)])
def test_detect_and_compute_matches_oracle(torch_cuda, oracle, dtype_name, w, h, nfeat, seed):
    import efb200, efo
    torch = torch_cuda
    dtype = getattr(efb200, dtype_name)
    img = oracle.synth_frame(util.SEED + seed, 0, w, h)
    ef = make_ef(nfeatures=nfeat, dtype=dtype, max_width=w, max_height=h)
    kp, desc = ef.detectAndComputeAsync(torch.from_numpy(img).cuda())
    g = ef.convert(kp)
    gd = desc.cpu().numpy()
    ok, od, _ = oracle.detect_and_compute(img, oracle.make_params(nfeatures=nfeat, desc_type=getattr(efo, dtype_name)))
    o = util.oracle_to_struct(ok)
    util.assert_keypoints_equal(g, o)
    _, go = util.canon_keypoints(g)
    _, oo = util.canon_keypoints(o)
    diff = gd[go] != od[oo]
    assert diff.sum() == 0, f"{dtype_name}: {diff.any(axis=1).sum()} of {len(g)} descriptors differ ({diff.sum()} bytes)"


def test_detect_only_and_params(torch_cuda, oracle):
    import efb200, efo
    torch = torch_cuda
    w, h = 1024, 768
    img = oracle.synth_frame(util.SEED + 11, 0, w, h)
    d_img = torch.from_numpy(img).cuda()
    for kw in (dict(nonmax_radius=5, fast_threshold=30, nfeatures=3000), dict(nonmax_radius=0, fast_threshold=60, nfeatures=4000),
               dict(first_level=2, nfeatures=1000), dict(nlevels=4, scale_factor=1.5, nfeatures=2000), dict(nonmax_radius=20, nfeatures=500)):
        p = dict(nfeatures=5000, scale_factor=1.2, nlevels=8, first_level=0, fast_threshold=20, nonmax_radius=15)
        p.update(kw)
        ef = make_ef(nfeatures=p["nfeatures"], scaleFactor=p["scale_factor"], nlevels=p["nlevels"], firstLevel=p["first_level"],
                     fastThreshold=p["fast_threshold"], nonmaxRadius=p["nonmax_radius"], dtype=efb200.BAD_256, max_width=w, max_height=h)
        g = ef.detect(d_img)
        ok, _ = oracle.detect(img, oracle.make_params(desc_type=efo.BAD_256, **p))
        util.assert_keypoints_equal(g, util.oracle_to_struct(ok))


# ---------------------------------------------------------------------------------------------------
# compute-only API (vector<KeyPoint> path) on a stress keypoint set: BAD bit-exact, HashSIFT features
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nbits", [256, 512])
@pytest.mark.parametrize("scale", [1.0, 5.0])
def test_bad_compute_bit_exact(torch_cuda, oracle, nbits, scale):
    import efb200, efo
    w, h = 1280, 720
    img = oracle.synth_frame(util.SEED + 21, 0, w, h)
    k = efo.stress_keypoints(w, h, 20000, seed=5)
    bad = efb200.BAD.create(scale, 100 if nbits == 512 else 101, max_width=w, max_height=h)
    g = bad.compute(img, k)
    o = oracle.bad(img, k, scale, nbits)
    diff = g != o
    assert diff.sum() == 0, f"BAD{nbits} scale {scale}: {diff.any(axis=1).sum()} of {len(k)} descriptors differ"


@pytest.mark.parametrize("nbits", [256, 512])
def test_hashsift_compute(torch_cuda, oracle, nbits):
    import efb200, efo
    w, h = 1280, 720
    img = oracle.synth_frame(util.SEED + 22, 0, w, h)
    k = efo.stress_keypoints(w, h, 20000, seed=6)
    hs = efb200.HashSIFT.create(1.0, 100 if nbits == 512 else 101, max_width=w, max_height=h)
    hs._ef.debugKeepProjection(True)
    g = hs.compute(img, k)
    sift, proj = hs._ef.debugHashSift(len(k))
    feat = oracle.hashsift_features(img, k, 1.0)
    o, oproj = oracle.hashsift(img, k, 1.0, nbits, want_proj=True)
    # 128-vector (u8-valued) identical; bits identical.  Projection: the GPU value is float32(EXACT dot product) (integer
    # tensor cores, one rounding), the oracle's is float32(double accumulation in ascending k): tolerance 1 ULP (the
    # north_star tolerance), and they may differ only through a double-rounding event (< 1e-6 of the outputs)
    assert np.array_equal(sift, feat[:, 1:].astype(np.uint8)), f"{(sift != feat[:, 1:]).any(axis=1).sum()} of {len(k)} SIFT vectors differ"
    ulp = np.abs(proj.view(np.int32).astype(np.int64) - oproj.view(np.int32).astype(np.int64))
    assert ulp.max() <= 1 and (np.sign(proj) == np.sign(oproj)).all(), f"projection differs by up to {ulp.max()} ULP"
    assert (ulp != 0).mean() < 1e-6, f"{(ulp != 0).sum()} projection values differ from the oracle"
    assert np.array_equal(g, o)


def test_compute_rows_forces_size_31(torch_cuda, oracle):
    import efb200, efo
    torch = torch_cuda
    w, h = 960, 540
    img = oracle.synth_frame(util.SEED + 23, 0, w, h)
    d_img = torch.from_numpy(img).cuda()
    ef = make_ef(nfeatures=3000, dtype=efb200.BAD_512, max_width=w, max_height=h)
    kp = ef.detectAsync(d_img)
    desc = ef.computeAsync(d_img, kp.contiguous())
    k = ef.convert(kp)
    k4 = np.stack([k["x"], k["y"], np.full(len(k), 31, np.float32), k["angle"]], axis=1)
    o = oracle.bad(img, k4, 1.0, 512)
    assert np.array_equal(desc.cpu().numpy(), o)


def test_host_api_and_batch(torch_cuda, oracle):
    import efb200, efo
    torch = torch_cuda
    w, h, F = 800, 608, 3
    frames = np.stack([oracle.synth_frame(util.SEED + 31, f, w, h) for f in range(F)])
    ef = make_ef(nfeatures=2500, dtype=efb200.BAD_512, max_width=w, max_height=h, max_batch=F)
    kps, descs = ef._host_call(frames, True)
    kpb, descb, counts = ef.detectAndComputeBatchRaw(torch.from_numpy(frames).cuda())
    torch.cuda.synchronize()
    for f in range(F):
        ok, od, _ = oracle.detect_and_compute(frames[f], oracle.make_params(nfeatures=2500, desc_type=efo.BAD_512))
        g = ef.convert(kps[f])
        util.assert_keypoints_equal(g, util.oracle_to_struct(ok))
        n = int(counts[f].item())
        assert n == len(ok)
        assert np.array_equal(kpb[f][:, :n].cpu().numpy().view(np.uint32), kps[f].view(np.uint32))
        assert np.array_equal(descb[f][:n].cpu().numpy(), descs[f])
        _, go = util.canon_keypoints(g)
        _, oo = util.canon_keypoints(util.oracle_to_struct(ok))
        assert np.array_equal(descs[f][go], od[oo])


def test_errors(torch_cuda):
    import efb200
    torch = torch_cuda
    ef = make_ef(nfeatures=100, max_width=640, max_height=480)
    img = torch.zeros((480, 640), dtype=torch.uint8, device="cuda")
    with pytest.raises(efb200.EfError):
        ef.detectAndComputeRaw(img, useProvidedKeypoints=True)
    with pytest.raises(efb200.EfError):
        ef.detectAndComputeRaw(img.float())
    with pytest.raises(efb200.EfError):
        ef.detectAndComputeRaw(torch.zeros((1000, 1000), dtype=torch.uint8, device="cuda"))
    kp, desc = ef.detectAndComputeAsync(img)  # blank image: zero keypoints, like the reference's release()
    assert kp.shape[1] == 0 and desc.shape[0] == 0


# ---------------------------------------------------------------------------------------------------
# BASELINE.json sizes: 4K / 40 000 requested keypoints against the oracle, 8K through properties
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype_name", ["HASH_SIFT_512", "BAD_512"])
def test_full_size_4k_matches_oracle(torch_cuda, oracle, dtype_name):
    """configs[1]/[2]/[4] geometry: 3840x2160, nfeatures 40000, r 15 -- bit-exact keypoints and descriptors"""
    import efb200, efo
    torch = torch_cuda
    w, h, nfeat = 3840, 2160, 40000
    img = oracle.synth_frame(util.SEED + 41, 0, w, h)
    ef = make_ef(nfeatures=nfeat, dtype=getattr(efb200, dtype_name), max_width=w, max_height=h)
    kp, desc = ef.detectAndComputeAsync(torch.from_numpy(img).cuda())
    g = ef.convert(kp)
    gd = desc.cpu().numpy()
    ok, od, _ = oracle.detect_and_compute(img, oracle.make_params(nfeatures=nfeat, desc_type=getattr(efo, dtype_name)))
    o = util.oracle_to_struct(ok)
    util.assert_keypoints_equal(g, o)
    _, go = util.canon_keypoints(g)
    _, oo = util.canon_keypoints(o)
    assert np.array_equal(gd[go], od[oo])


@pytest.mark.parametrize("dtype_name", ["HASH_SIFT_512", "BAD_512"])
def test_full_size_8k_matches_oracle(torch_cuda, oracle, dtype_name):
    """BASELINE.json configs[3]: 7680x4320, 40 000 keypoints, against the oracle (1-2 s on 16 host threads).  The only
    configuration where EVERY per-level quota binds (the radix select cuts each level) and where a full-frame int32 integral
    would wrap (7680 * 4320 * 255 > 2^31): keypoints and descriptor bytes identical.  The compute-only BAD path
    (vector<KeyPoint>, full-frame integral image like the reference's cv::integral) is checked on the same frame."""
    import efb200, efo
    torch = torch_cuda
    w, h, nfeat = 7680, 4320, 40000
    img = oracle.synth_frame(util.SEED + 43, 0, w, h)
    ef = make_ef(nfeatures=nfeat, dtype=getattr(efb200, dtype_name), max_width=w, max_height=h)
    d_img = torch.from_numpy(img).cuda()
    kp, desc = ef.detectAndComputeAsync(d_img)
    g = ef.convert(kp)
    gd = desc.cpu().numpy()
    ok, od, _ = oracle.detect_and_compute(img, oracle.make_params(nfeatures=nfeat, desc_type=getattr(efo, dtype_name)))
    o = util.oracle_to_struct(ok)
    assert len(o) == nfeat, f"8K noise must fill every quota, oracle delivers {len(o)}"
    util.assert_keypoints_equal(g, o)
    _, go = util.canon_keypoints(g)
    _, oo = util.canon_keypoints(o)
    diff = gd[go] != od[oo]
    assert diff.sum() == 0, f"{dtype_name}@8K: {diff.any(axis=1).sum()} of {len(g)} descriptors differ"
    del ef
    if dtype_name == "BAD_512":
        # compute-only on the level-0 image: keypoints near the bottom-right corner see wrapped int32 prefix sums
        k = np.stack([g["x"], g["y"], g["size"], g["angle"]], axis=1).astype(np.float32)[:20000]
        k = np.concatenate([k, efo.stress_keypoints(w, h, 6000, seed=13)])
        k[-3000:, 0] = np.minimum(k[-3000:, 0] * 0.05 + (w - 400), w - 1)       # bottom-right 400 x 250 px block
        k[-3000:, 1] = np.minimum(k[-3000:, 1] * 0.05 + (h - 250), h - 1)
        gb = efb200.BAD.create(1.0, 100, max_width=w, max_height=h, max_keypoints=len(k)).compute(img, k)
        ob = oracle.bad(img, k, 1.0, 512)
        assert np.array_equal(gb, ob), f"compute-only BAD-512 at 8K: {(gb != ob).any(axis=1).sum()} of {len(k)} descriptors differ"


def nms_property_ok(k, radius, scales):
    """no two keypoints of the same level closer than the radius in level coordinates (checked on the scaled coordinates with the
    rounding slack of scalePoints: |x' - s x| < 1)"""
    for octave in np.unique(k["octave"]):
        s = float(scales[octave])
        m = k[k["octave"] == octave]
        xs, ys = m["x"].astype(np.float64) / s, m["y"].astype(np.float64) / s
        order = np.argsort(ys)
        xs, ys = xs[order], ys[order]
        lim = radius - 2.0  # r minus the rounding slack of both points
        for i in range(len(xs)):
            j = i + 1
            while j < len(xs) and ys[j] - ys[i] < lim:
                if (xs[j] - xs[i]) ** 2 + (ys[j] - ys[i]) ** 2 < lim * lim:
                    return False
                j += 1
    return True


def test_8k_properties_and_determinism(torch_cuda):
    """configs[3] geometry (7680x4320, HashSIFT-512, 40 000 keypoints): size-independent properties instead of the oracle --
    every per-level quota binds at 8K (count == nfeatures), NMS spacing holds, the run is deterministic, and the batched call
    returns for every frame exactly what the single-frame call returns."""
    import efb200
    torch = torch_cuda
    w, h, nfeat = 7680, 4320, 40000
    gen = torch.Generator(device="cpu").manual_seed(1234)
    frames = torch.randint(0, 256, (2, h, w), dtype=torch.uint8, generator=gen).cuda()
    ef = make_ef(nfeatures=nfeat, dtype=efb200.HASH_SIFT_512, max_width=w, max_height=h, max_batch=2)
    kp1, d1 = ef.detectAndComputeAsync(frames[0])
    kp2, d2 = ef.detectAndComputeAsync(frames[0])
    assert kp1.shape[1] == nfeat, f"8K noise must fill every quota, got {kp1.shape[1]}"
    assert torch.equal(kp1.view(torch.int32), kp2.view(torch.int32)) and torch.equal(d1, d2), "run-to-run determinism"
    k = ef.convert(kp1)
    scales = np.float32(1.2) ** np.arange(8, dtype=np.float32)
    assert nms_property_ok(k, 15, scales)
    assert (k["octave"] >= 0).all() and (k["octave"] < 8).all()
    assert np.array_equal(np.sort(np.bincount(k["octave"], minlength=8))[::-1], np.bincount(k["octave"], minlength=8)), "quotas decrease with the level"
    kpb, db, counts = ef.detectAndComputeBatchRaw(frames)
    torch.cuda.synchronize()
    assert int(counts[0]) == nfeat
    assert torch.equal(kpb[0].view(torch.int32), kp1.contiguous().view(torch.int32)) and torch.equal(db[0], d1)
    # frame 1 differs from frame 0 (no cross-frame leakage of workspace slots)
    assert not torch.equal(db[1], db[0])


# ---------------------------------------------------------------------------------------------------
# edge cases: sizes, alignments, parameter boundaries of the tiled kernels (all against the oracle)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("w,h", [(32, 32), (33, 47), (64, 35), (129, 130), (255, 257), (771, 99)])
def test_small_and_odd_sizes(torch_cuda, oracle, w, h):
    """minimum size, widths that are not multiples of 4 / 32 / 64 / 128 (byte-wise load paths, partial tiles, 1-word resize windows)"""
    import efb200, efo
    torch = torch_cuda
    img = oracle.synth_frame(util.SEED + 51, w * 1000 + h, w, h)
    for dtype_name in ("BAD_256", "HASH_SIFT_256"):
        ef = make_ef(nfeatures=500, nlevels=4, dtype=getattr(efb200, dtype_name), max_width=w, max_height=h)
        kp, desc = ef.detectAndComputeAsync(torch.from_numpy(img).cuda())
        ok, od, _ = oracle.detect_and_compute(img, oracle.make_params(nfeatures=500, nlevels=4, desc_type=getattr(efo, dtype_name)))
        g, o = ef.convert(kp), util.oracle_to_struct(ok)
        util.assert_keypoints_equal(g, o)
        if len(g):
            _, go = util.canon_keypoints(g)
            _, oo = util.canon_keypoints(o)
            assert np.array_equal(desc.cpu().numpy()[go], od[oo])


def test_unaligned_caller_image(torch_cuda, oracle):
    """level 0 is the caller's buffer: odd base address and odd pitch (a column slice of a larger tensor)"""
    import efb200, efo
    torch = torch_cuda
    w, h = 517, 389
    img = oracle.synth_frame(util.SEED + 52, 0, w, h)
    big = torch.zeros((h, w + 6), dtype=torch.uint8, device="cuda")
    big[:, 3:3 + w] = torch.from_numpy(img).cuda()
    view = big[:, 3:3 + w]                      # data_ptr % 4 == 3, stride(0) = w + 6 = 523 (odd)
    assert view.data_ptr() % 4 != 0 and view.stride(0) % 4 != 0
    for dtype_name in ("BAD_512", "HASH_SIFT_512"):
        ef = make_ef(nfeatures=1500, dtype=getattr(efb200, dtype_name), max_width=w, max_height=h)
        kp, desc = ef.detectAndComputeAsync(view)
        ok, od, _ = oracle.detect_and_compute(img, oracle.make_params(nfeatures=1500, desc_type=getattr(efo, dtype_name)))
        g, o = ef.convert(kp), util.oracle_to_struct(ok)
        util.assert_keypoints_equal(g, o)
        _, go = util.canon_keypoints(g)
        _, oo = util.canon_keypoints(o)
        assert np.array_equal(desc.cpu().numpy()[go], od[oo])


@pytest.mark.parametrize("radius", [1, 2, 3, 4, 7, 9, 10, 11, 31, 64])
def test_nms_radius_sweep(torch_cuda, oracle, radius):
    """block edge 0 (no suppression), 2, 4 and 8 of the block-maximum NMS, and the widest neighbourhood (r = 64)"""
    import efb200, efo
    torch = torch_cuda
    w, h = 640, 360
    img = oracle.synth_frame(util.SEED + 53, radius, w, h)
    p = dict(nfeatures=4000, scale_factor=1.2, nlevels=5, first_level=0, fast_threshold=25, nonmax_radius=radius)
    ef = make_ef(nfeatures=p["nfeatures"], nlevels=p["nlevels"], fastThreshold=p["fast_threshold"], nonmaxRadius=radius,
                 dtype=efb200.BAD_256, max_width=w, max_height=h)
    g = ef.detect(torch.from_numpy(img).cuda())
    ok, _ = oracle.detect(img, oracle.make_params(desc_type=efo.BAD_256, **p))
    util.assert_keypoints_equal(g, util.oracle_to_struct(ok))


@pytest.mark.parametrize("kw", [dict(scale_factor=1.1, nlevels=12), dict(scale_factor=2.0, nlevels=4), dict(scale_factor=1.3, nlevels=1),
                                dict(fast_threshold=0, nfeatures=3000), dict(fast_threshold=255), dict(nfeatures=1), dict(nfeatures=100000)])
def test_parameter_extremes(torch_cuda, oracle, kw):
    import efb200, efo
    torch = torch_cuda
    w, h = 512, 384
    img = oracle.synth_frame(util.SEED + 54, 0, w, h)
    p = dict(nfeatures=2000, scale_factor=1.2, nlevels=8, first_level=0, fast_threshold=20, nonmax_radius=15)
    p.update(kw)
    ef = make_ef(nfeatures=p["nfeatures"], scaleFactor=p["scale_factor"], nlevels=p["nlevels"], fastThreshold=p["fast_threshold"],
                 dtype=efb200.HASH_SIFT_256, max_width=w, max_height=h)
    kp, desc = ef.detectAndComputeAsync(torch.from_numpy(img).cuda())
    ok, od, _ = oracle.detect_and_compute(img, oracle.make_params(desc_type=efo.HASH_SIFT_256, **p))
    g, o = ef.convert(kp), util.oracle_to_struct(ok)
    util.assert_keypoints_equal(g, o)
    if len(g):
        _, go = util.canon_keypoints(g)
        _, oo = util.canon_keypoints(o)
        assert np.array_equal(desc.cpu().numpy()[go], od[oo])


@pytest.mark.parametrize("value", [0, 255, "gradient", "checker"])
def test_degenerate_images(torch_cuda, oracle, value):
    """constant images (no corner), a smooth ramp (ties everywhere) and a 1-px checkerboard (every interior pixel is a corner,
    all Harris responses of a row equal: exercises the tie rules of the block-maximum NMS and of the top-K select)"""
    import efb200, efo
    torch = torch_cuda
    w, h = 320, 240
    if value == "gradient":
        img = ((np.arange(w)[None, :] + 2 * np.arange(h)[:, None]) % 256).astype(np.uint8)
    elif value == "checker":
        img = (((np.arange(w)[None, :] + np.arange(h)[:, None]) % 2) * 255).astype(np.uint8)
    else:
        img = np.full((h, w), value, np.uint8)
    img = np.ascontiguousarray(img)
    ef = make_ef(nfeatures=1000, nlevels=3, dtype=efb200.BAD_256, max_width=w, max_height=h)
    kp, desc = ef.detectAndComputeAsync(torch.from_numpy(img).cuda())
    ok, od, _ = oracle.detect_and_compute(img, oracle.make_params(nfeatures=1000, nlevels=3, desc_type=efo.BAD_256))
    g, o = ef.convert(kp), util.oracle_to_struct(ok)
    util.assert_keypoints_equal(g, o)
    if len(g):
        _, go = util.canon_keypoints(g)
        _, oo = util.canon_keypoints(o)
        assert np.array_equal(desc.cpu().numpy()[go], od[oo])


def test_single_process_multi_gpu_driver(torch_cuda, oracle):
    """ef_mg_*: frames sharded over every visible device (1 on the CI box, 2+ under gpurun --gpus N) from one process;
    results must equal the single-handle results frame by frame, and the oracle on a sample."""
    import efb200, efo
    torch = torch_cuda
    w, h, F = 640, 480, 7
    frames = np.stack([oracle.synth_frame(util.SEED + 61, f, w, h) for f in range(F)])
    mg = efb200.MultiGpuEfficientFeatures(nfeatures=1200, dtype=efb200.BAD_512, max_width=w, max_height=h, max_batch=2)
    kps, descs = mg.detectAndComputeHost(frames)
    ef = make_ef(nfeatures=1200, dtype=efb200.BAD_512, max_width=w, max_height=h, max_batch=F)
    rk, rd = ef._host_call(frames, True)
    for f in range(F):
        assert np.array_equal(kps[f].view(np.uint32), rk[f].view(np.uint32)) and np.array_equal(descs[f], rd[f]), f"frame {f}"
    ok, od, _ = oracle.detect_and_compute(frames[F - 1], oracle.make_params(nfeatures=1200, desc_type=efo.BAD_512))
    util.assert_keypoints_equal(efb200.EfficientFeatures.convert(kps[F - 1]), util.oracle_to_struct(ok))
    mg.close()


# ---------------------------------------------------------------------------------------------------
# the projection stage alone: tcgen05 == mma.sync (both the exact six-digit integer GEMM) == fp64 CUDA cores
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype_name", ["HASH_SIFT_256", "HASH_SIFT_512"])
def test_projection_paths_agree(torch_cuda, dtype_name):
    import efb200
    torch = torch_cuda
    ef = make_ef(nfeatures=1000, dtype=getattr(efb200, dtype_name), max_width=640, max_height=480, max_keypoints=70000)
    g = torch.Generator(device="cpu").manual_seed(7)
    cases = {
        "uniform": torch.randint(0, 256, (60001, 128), dtype=torch.uint8, generator=g),
        "sparse": (torch.randint(0, 256, (9000, 128), dtype=torch.uint8, generator=g) * (torch.rand((9000, 128), generator=g) < 0.1)).to(torch.uint8),
        "zeros_and_max": torch.cat([torch.zeros((130, 128), dtype=torch.uint8), torch.full((130, 128), 255, dtype=torch.uint8)]),
        "siftlike": torch.clamp((torch.randn((20000, 128), generator=g).abs() * 40), 0, 255).to(torch.uint8),
        "tiny": torch.randint(0, 3, (5, 128), dtype=torch.uint8, generator=g),
    }
    for name, x in cases.items():
        d = x.cuda().contiguous()
        ref = ef.debugProject(d, 3).cpu().numpy()       # double accumulation of the fp32 table (the oracle's definition)
        tc = ef.debugProject(d, 1).cpu().numpy()
        imma = ef.debugProject(d, 2).cpu().numpy()
        assert np.array_equal(imma, ref), f"{name}: mma.sync path differs from fp64 in {(imma != ref).any(axis=1).sum()} rows"
        assert np.array_equal(tc, ref), f"{name}: tcgen05 path differs from fp64 in {(tc != ref).any(axis=1).sum()} rows"
        assert np.array_equal(ef.debugProject(d, 0).cpu().numpy(), ref)


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configs[2]: compute-only, BAD512 + HashSIFT512 on 40 000 precomputed keypoints at 4K
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype_name", ["BAD_512", "HASH_SIFT_512"])
def test_compute_only_40k_keypoints_4k(torch_cuda, oracle, dtype_name):
    import efb200, efo
    torch = torch_cuda
    w, h, n = 3840, 2160, 40000
    img = oracle.synth_frame(util.SEED + 51, 0, w, h)
    det = make_ef(nfeatures=n, dtype=efb200.BAD_256, max_width=w, max_height=h)
    kd = det.detect(torch.from_numpy(img).cuda())                      # ~30 000 detector keypoints (sizes 31 * scale, IC angles)
    k = np.stack([kd["x"], kd["y"], kd["size"], kd["angle"]], axis=1).astype(np.float32)
    k = np.concatenate([k, efo.stress_keypoints(w, h, n - len(k), seed=9)])   # + border band, special angles, sizes 31..111
    assert len(k) == n
    if dtype_name == "BAD_512":
        g = efb200.BAD.create(1.0, 100, max_width=w, max_height=h, max_keypoints=n).compute(img, k)
        o = oracle.bad(img, k, 1.0, 512)
    else:
        g = efb200.HashSIFT.create(1.0, 100, max_width=w, max_height=h, max_keypoints=n).compute(img, k)   # default path: tcgen05 projection
        o = oracle.hashsift(img, k, 1.0, 512)
    assert g.shape == (n, 64)
    assert np.array_equal(g, o), f"{(g != o).any(axis=1).sum()} of {n} descriptors differ"


# ---------------------------------------------------------------------------------------------------
# 5 x N GpuMat compute path (window-staging kernels when the image is 16-byte aligned, generic kernels otherwise):
# arbitrary integer positions including the border band and outside the detector's 15-px margin, special angles
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype_name", ["BAD_256", "BAD_512", "HASH_SIFT_256", "HASH_SIFT_512"])
@pytest.mark.parametrize("aligned", [True, False])
def test_compute_rows_staged_and_generic_paths(torch_cuda, oracle, dtype_name, aligned):
    import efb200, efo
    torch = torch_cuda
    w, h, n = 1000, 700, 6000
    img = oracle.synth_frame(util.SEED + 24, 0, w, h)
    rng = np.random.default_rng(11)
    xs = rng.integers(0, w, n).astype(np.int16); ys = rng.integers(0, h, n).astype(np.int16)
    xs[:40] = [0, 1, 2, 14, 15, 16, 22, 23, 24, 26, 27, 28, w - 1, w - 2, w - 15, w - 16, w - 23, w - 24, w - 27, w - 28] * 2
    ys[40:80] = [0, 1, 2, 14, 15, 16, 22, 23, 24, 26, 27, 28, h - 1, h - 2, h - 15, h - 16, h - 23, h - 24, h - 27, h - 28] * 2
    ang = rng.uniform(0, 360, n).astype(np.float32)
    ang[::7] = np.array([-1.0, 0.0, 90.0, 180.0, 270.0, 359.99, 45.0], np.float32)[np.arange(len(ang[::7])) % 7]
    rows = np.zeros((5, n), np.float32)
    rows[0] = np.stack([xs, ys], axis=1).copy().view(np.float32)[:, 0]          # short2 packed into the float row
    rows[2] = ang
    rows[4] = 77.0                                                                # ignored: the GpuMat path forces size 31
    ef = make_ef(nfeatures=100, dtype=getattr(efb200, dtype_name), max_width=w + 8, max_height=h, max_keypoints=n)
    if aligned:
        d_img = torch.from_numpy(img).cuda()                                     # pitch 1000 is not a multiple of 16 -> pad to 1008
        buf = torch.zeros((h, 1008), dtype=torch.uint8, device="cuda")
        buf[:, :w] = d_img
        d_img = buf[:, :w]
        assert d_img.data_ptr() % 16 == 0 and d_img.stride(0) % 16 == 0
    else:
        buf = torch.zeros((h, w + 5), dtype=torch.uint8, device="cuda")
        buf[:, 3:w + 3] = torch.from_numpy(img).cuda()
        d_img = buf[:, 3:w + 3]                                                   # base and pitch unaligned: generic kernels
    desc = ef.computeAsync(d_img, torch.from_numpy(rows).cuda())
    k4 = np.stack([xs.astype(np.float32), ys.astype(np.float32), np.full(n, 31, np.float32), ang], axis=1)
    o = oracle.bad(img, k4, 1.0, 256 if "256" in dtype_name else 512) if dtype_name.startswith("BAD") else oracle.hashsift(img, k4, 1.0, 256 if "256" in dtype_name else 512)
    g = desc.cpu().numpy()
    assert np.array_equal(g, o), f"{(g != o).any(axis=1).sum()} of {n} descriptors differ (first rows {np.nonzero((g != o).any(axis=1))[0][:8]})"
