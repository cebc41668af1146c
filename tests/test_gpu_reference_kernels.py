"""Pins the detector against the reference ITSELF: the reference's own CUDA kernels (cuda_fast.cu, cuda_efficient_features.cu, compiled
unmodified into oracle/_ref/libef_ref_cuda.so) run on this GPU next to the CPU oracle and the product kernels, stage by stage:
FAST-9 corner set, Harris response (bit for bit: the FMA contraction of the canonical arithmetic is what nvcc emits for the reference
source), radius NMS survivors, limitPoints, scalePoints.  IC angle: the reference calls CUDA atan2f (<= 2 ulp, GPU-only); the canonical
value here is atan2 in double rounded once, so angles are compared within 2 ulp and the moments path is exact."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def refcu():
    import efo
    if not efo.ReferenceCuda.available():
        pytest.skip("oracle/_ref/libef_ref_cuda.so not built (needs /root/reference at build time)")
    return efo.ReferenceCuda()


def images(oracle):
    rng = np.random.default_rng(4)
    yield "noise", oracle.synth_frame(util.SEED + 61, 0, 640, 480)
    # smoother content: few corners, plateaus and ties
    base = rng.integers(0, 256, (60, 80), dtype=np.uint8)
    yield "blocks", np.kron(base, np.ones((8, 8), np.uint8))[:480, :640].copy()
    g = (np.add.outer(np.arange(333), np.arange(517)) % 256).astype(np.uint8)
    g[100:140, 200:260] = 255; g[200:203, :] = 0
    yield "ramp", g


def sort_xy(xy, *cols):
    o = np.lexsort((xy[:, 0], xy[:, 1]))
    return (xy[o],) + tuple(c[o] for c in cols)


def test_fast_and_harris_equal_reference_kernels(oracle, refcu):
    import torch
    import efb200
    for name, img in images(oracle):
        h, w = img.shape
        xy = refcu.fast(img)                                           # every corner (capacity = w * h: no overflow)
        resp_ref, ang_ref = refcu.responses_angles(img, xy)
        xy, resp_ref, ang_ref = sort_xy(xy, resp_ref, ang_ref)
        # CPU oracle
        omap, ncorner = oracle.score_map(img, 20)
        oy, ox = np.nonzero(np.isfinite(omap))
        assert ncorner == len(xy), f"{name}: oracle finds {ncorner} corners, the reference kernel {len(xy)}"
        assert np.array_equal(np.stack([ox, oy], 1).astype(np.int16), xy), f"{name}: FAST corner set differs (oracle vs reference kernel)"
        assert np.array_equal(omap[oy, ox].view(np.uint32), resp_ref.view(np.uint32)), f"{name}: Harris response differs (oracle vs reference kernel)"
        # product kernels (level 0 response map)
        ef = efb200.EfficientFeatures.create(nfeatures=1000, nlevels=1, dtype=efb200.BAD_256, max_width=w, max_height=h)
        ef.detectAsync(torch.from_numpy(img).cuda())
        gmap = ef.debugLevelArrays(0, want=("response",))["response"]
        gy, gx = np.nonzero(np.isfinite(gmap))
        assert np.array_equal(np.stack([gx, gy], 1).astype(np.int16), xy), f"{name}: FAST corner set differs (product vs reference kernel)"
        assert np.array_equal(gmap[gy, gx].view(np.uint32), resp_ref.view(np.uint32)), f"{name}: Harris response differs (product vs reference kernel)"
        # IC angle: exact moments, atan2 within 2 ulp of the reference's CUDA atan2f
        if len(xy):
            sel = np.linspace(0, len(xy) - 1, min(len(xy), 3000)).astype(int)
            oang = np.array([oracle.ic_angle(img, int(x), int(y)) for x, y in xy[sel]], np.float32)
            ulp = np.abs(oang.view(np.int32).astype(np.int64) - ang_ref[sel].view(np.int32).astype(np.int64))
            big = ulp > 2
            # 0 vs 360: the only place where 2 ulp of atan2f can wrap
            assert not big.any() or np.all(np.minimum(np.abs(oang[big] - ang_ref[sel][big]), 360 - np.abs(oang[big] - ang_ref[sel][big])) < 1e-3), f"{name}: angle differs by {ulp.max()} ulp"


@pytest.mark.parametrize("radius", [15, 7, 3, 31])
def test_radius_nms_equals_reference_kernel(oracle, refcu, radius):
    import torch
    import efb200
    for name, img in images(oracle):
        h, w = img.shape
        omap, _ = oracle.score_map(img, 20)
        oy, ox = np.nonzero(np.isfinite(omap))
        xy = np.stack([ox, oy], 1).astype(np.int16)
        sxy, sresp = refcu.nms_limit(xy, omap[oy, ox], w, h, float(radius), -1)
        sxy, sresp = sort_xy(sxy, sresp)
        nx, ny, nr = oracle.radius_nms(omap, radius)
        oxy, orr = sort_xy(np.stack([nx, ny], 1).astype(np.int16), nr)
        assert np.array_equal(oxy, sxy) and np.array_equal(orr.view(np.uint32), sresp.view(np.uint32)), f"{name} r={radius}: NMS survivors differ (oracle vs reference kernel): {len(oxy)} vs {len(sxy)}"
        # product: one level, quota larger than the survivor count -> the keypoints ARE the survivors
        ef = efb200.EfficientFeatures.create(nfeatures=max(len(sxy), 1) + 10, nlevels=1, nonmaxRadius=radius, dtype=efb200.BAD_256, max_width=w, max_height=h)
        k = ef.detect(torch.from_numpy(img).cuda())
        gxy, gresp = sort_xy(np.stack([k["x"], k["y"]], 1).astype(np.int16), k["response"])
        assert np.array_equal(gxy, sxy) and np.array_equal(gresp.view(np.uint32), sresp.view(np.uint32)), f"{name} r={radius}: NMS survivors differ (product vs reference kernel)"


def test_limit_points_and_scale_points_equal_reference_kernels(oracle, refcu):
    import torch
    import efb200
    img = oracle.synth_frame(util.SEED + 62, 0, 800, 600)
    h, w = img.shape
    omap, _ = oracle.score_map(img, 20)
    oy, ox = np.nonzero(np.isfinite(omap))
    xy = np.stack([ox, oy], 1).astype(np.int16)
    quota = 200
    lxy, lresp = refcu.nms_limit(xy, omap[oy, ox], w, h, 15.0, quota)            # radiusSuppression + limitPoints (thrust sort, truncate)
    assert len(lxy) == quota
    thr = np.sort(lresp)[0]
    ef = efb200.EfficientFeatures.create(nfeatures=10000, nlevels=1, dtype=efb200.BAD_256, max_width=w, max_height=h)
    # quota of level 0 with nlevels = 1 is nfeatures: ask for exactly `quota`
    ef.setMaxFeatures(quota)
    k = ef.detect(torch.from_numpy(img).cuda())
    assert len(k) == quota
    # same SET (ties at the cut are broken by arrival order in the reference: none on this frame)
    assert (np.sort(lresp) == np.sort(k["response"])).all() and (lresp == thr).sum() == 1
    lxy_s, = sort_xy(lxy)
    gxy_s, = sort_xy(np.stack([k["x"], k["y"]], 1).astype(np.int16))
    assert np.array_equal(lxy_s, gxy_s)
    # scalePoints at every level scale of an 8-level pyramid
    scales = oracle.level_geometry(3840, 2160)[2]
    pts = np.stack([np.arange(15, 3815, 7), (np.arange(15, 3815, 7) * 5) % 2100 + 15], 1).astype(np.int16)
    for octave, s in enumerate(scales):
        sxy, oc, sz = refcu.scale(pts, float(s), octave)
        ex = np.floor(np.float32(s) * pts.astype(np.float32) + np.float32(0.5))          # documentation only; the exact statement follows
        # oracle / product formula: (short)fmaf(scale, x, 0.5f), octave, scale * 31
        fx = np.array([np.float32(np.float64(np.float32(s)) * np.float64(v) + 0.5) for v in pts[:, 0]], np.float32)   # fma = one rounding of the exact sum
        fy = np.array([np.float32(np.float64(np.float32(s)) * np.float64(v) + 0.5) for v in pts[:, 1]], np.float32)
        assert np.array_equal(sxy[:, 0], fx.astype(np.int16)) and np.array_equal(sxy[:, 1], fy.astype(np.int16))
        assert (oc == octave).all() and np.array_equal(sz.view(np.uint32), np.full(len(pts), np.float32(s) * np.float32(31.0), np.float32).view(np.uint32))
