"""ctypes bindings for the CPU oracle (oracle/libef_oracle.so) and, when built, the reference's own
CPU descriptors (oracle/_ref/libef_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing in the product package imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "libef_oracle.so"
REF_SO = HERE / "_ref" / "libef_ref.so"

BAD_256, BAD_512, HASH_SIFT_256, HASH_SIFT_512 = 0, 1, 2, 3
DESC_BYTES = {BAD_256: 32, BAD_512: 64, HASH_SIFT_256: 32, HASH_SIFT_512: 64}
MAX_LEVELS = 16


class Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int),
                ("first_level", C.c_int), ("fast_threshold", C.c_int), ("nonmax_radius", C.c_int),
                ("desc_type", C.c_int)]


KPT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4")])
KEYPOINT_DTYPE = np.dtype([("x", "<i2"), ("y", "<i2"), ("response", "<f4"), ("angle", "<f4"),
                           ("octave", "<i4"), ("size", "<f4"), ("lx", "<i2"), ("ly", "<i2")])
assert KEYPOINT_DTYPE.itemsize == 24


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < (HERE / "ef_oracle.c").stat().st_mtime:
        subprocess.check_call(["make", "-C", str(HERE), "-s", str(ORACLE_SO)])
    if Path("/root/reference/modules/efficient_features/src").is_dir() and (force or not REF_SO.exists()):
        subprocess.check_call(["make", "-C", str(HERE), "-s", "ref"])


_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def _img(img: np.ndarray):
    assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
    return _p(img, _u8p), img.shape[1], img.shape[0], C.c_size_t(img.strides[0])


class Oracle:
    def __init__(self):
        if not ORACLE_SO.exists():
            build()
        L = C.CDLL(str(ORACLE_SO))
        self.L = L
        L.efo_harris_response.restype = C.c_float
        L.efo_ic_angle.restype = C.c_float
        L.efo_score_map.restype = C.c_long
        L.efo_radius_nms.restype = C.c_long
        L.efo_build_pyramid.restype = C.c_size_t
        L.efo_build_pyramid.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, C.c_float, C.c_int, C.c_int, _u8p,
                                        C.POINTER(C.c_size_t)]
        L.efo_level_geometry.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), _f32p]
        L.efo_level_quotas.argtypes = [C.c_int, C.c_float, C.c_int, C.POINTER(C.c_int)]
        L.efo_resize_linear.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, _u8p, C.c_int, C.c_int, C.c_size_t]
        L.efo_gaussian_blur7.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, _u8p, C.c_size_t]
        L.efo_score_map.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, _f32p]
        L.efo_radius_nms.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_short), C.POINTER(C.c_short), _f32p, C.c_long]
        L.efo_harris_response.argtypes = [_u8p, C.c_size_t, C.c_int, C.c_int]
        L.efo_ic_angle.argtypes = [_u8p, C.c_size_t, C.c_int, C.c_int]
        L.efo_fast_is_corner.argtypes = [_u8p, C.c_size_t, C.c_int, C.c_int, C.c_int]
        L.efo_detect.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, C.POINTER(Params), C.c_void_p, C.c_int, C.POINTER(C.c_long)]
        L.efo_detect_and_compute.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, C.POINTER(Params), C.c_void_p, _u8p, C.c_int, C.POINTER(C.c_long)]
        L.efo_bad_compute.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_float, C.c_int, _u8p]
        L.efo_hashsift_features.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_float, _f32p]
        L.efo_hashsift_patch.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_float, _u8p]
        L.efo_hashsift_project.argtypes = [_f32p, C.c_int, C.c_int, _u8p, _f32p]
        L.efo_hashsift_compute.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_float, C.c_int, _u8p, _f32p]
        L.efo_synth_frame.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_size_t, _u8p]

    # ---- runtime -------------------------------------------------------------------------------
    def set_threads(self, n: int) -> None:
        self.L.efo_set_threads(int(n))

    def max_threads(self) -> int:
        return int(self.L.efo_get_max_threads())

    # ---- geometry ------------------------------------------------------------------------------
    def level_geometry(self, w, h, scale_factor=1.2, nlevels=8):
        ws = (C.c_int * nlevels)(); hs = (C.c_int * nlevels)(); sc = (C.c_float * nlevels)()
        self.L.efo_level_geometry(w, h, scale_factor, nlevels, ws, hs, sc)
        return list(ws), list(hs), [np.float32(v) for v in sc]

    def level_quotas(self, nfeatures, scale_factor=1.2, nlevels=8):
        q = (C.c_int * nlevels)()
        self.L.efo_level_quotas(nfeatures, scale_factor, nlevels, q)
        return list(q)

    # ---- image stages --------------------------------------------------------------------------
    def synth_frame(self, seed, frame, w, h):
        out = np.empty((h, w), np.uint8)
        self.L.efo_synth_frame(seed, frame, w, h, w, _p(out, _u8p))
        return out

    def resize_linear(self, src, dw, dh):
        p, w, h, pitch = _img(src)
        dst = np.empty((dh, dw), np.uint8)
        self.L.efo_resize_linear(p, w, h, pitch, _p(dst, _u8p), dw, dh, dw)
        return dst

    def gaussian_blur7(self, src):
        p, w, h, pitch = _img(src)
        dst = np.empty((h, w), np.uint8)
        self.L.efo_gaussian_blur7(p, w, h, pitch, _p(dst, _u8p), w)
        return dst

    def pyramid(self, img, scale_factor=1.2, nlevels=8, blurred=False):
        p, w, h, pitch = _img(img)
        offs = (C.c_size_t * nlevels)()
        total = self.L.efo_build_pyramid(p, w, h, pitch, scale_factor, nlevels, int(blurred), None, offs)
        buf = np.empty(total, np.uint8)
        self.L.efo_build_pyramid(p, w, h, pitch, scale_factor, nlevels, int(blurred), _p(buf, _u8p), offs)
        ws, hs, _ = self.level_geometry(w, h, scale_factor, nlevels)
        return [buf[offs[i]:offs[i] + ws[i] * hs[i]].reshape(hs[i], ws[i]) for i in range(nlevels)]

    def score_map(self, img, threshold=20):
        p, w, h, pitch = _img(img)
        resp = np.empty((h, w), np.float32)
        n = self.L.efo_score_map(p, w, h, pitch, threshold, _p(resp, _f32p))
        return resp, int(n)

    def radius_nms(self, resp, radius=15):
        h, w = resp.shape
        cap = max(int(np.isfinite(resp).sum()), 1)
        xs = np.empty(cap, np.int16); ys = np.empty(cap, np.int16); rs = np.empty(cap, np.float32)
        n = self.L.efo_radius_nms(_p(resp, _f32p), w, h, radius, _p(xs, C.POINTER(C.c_short)), _p(ys, C.POINTER(C.c_short)), _p(rs, _f32p), cap)
        return xs[:n].copy(), ys[:n].copy(), rs[:n].copy()

    def harris(self, img, x, y):
        p, w, h, pitch = _img(img)
        return np.float32(self.L.efo_harris_response(p, pitch, int(x), int(y)))

    def ic_angle(self, img, x, y):
        p, w, h, pitch = _img(img)
        return np.float32(self.L.efo_ic_angle(p, pitch, int(x), int(y)))

    def fast_is_corner(self, img, x, y, th=20):
        p, w, h, pitch = _img(img)
        return bool(self.L.efo_fast_is_corner(p, pitch, int(x), int(y), th))

    # ---- detector ------------------------------------------------------------------------------
    @staticmethod
    def make_params(nfeatures=5000, scale_factor=1.2, nlevels=8, first_level=0, fast_threshold=20,
                    nonmax_radius=15, desc_type=HASH_SIFT_256) -> Params:
        return Params(nfeatures, scale_factor, nlevels, first_level, fast_threshold, nonmax_radius, desc_type)

    def detect(self, img, params: Params):
        p, w, h, pitch = _img(img)
        cap = params.nfeatures
        out = np.zeros(max(cap, 1), KEYPOINT_DTYPE)
        counts = (C.c_long * (3 * params.nlevels))()
        n = self.L.efo_detect(p, w, h, pitch, C.byref(params), out.ctypes.data, cap, counts)
        return out[:n].copy(), np.array(list(counts)).reshape(params.nlevels, 3)

    def detect_and_compute(self, img, params: Params):
        p, w, h, pitch = _img(img)
        cap = params.nfeatures
        out = np.zeros(max(cap, 1), KEYPOINT_DTYPE)
        nb = DESC_BYTES[params.desc_type]
        desc = np.zeros((max(cap, 1), nb), np.uint8)
        counts = (C.c_long * (3 * params.nlevels))()
        n = self.L.efo_detect_and_compute(p, w, h, pitch, C.byref(params), out.ctypes.data, _p(desc, _u8p), cap, counts)
        return out[:n].copy(), desc[:n].copy(), np.array(list(counts)).reshape(params.nlevels, 3)

    # ---- descriptors ---------------------------------------------------------------------------
    @staticmethod
    def as_kpts(kpts) -> np.ndarray:
        k = np.ascontiguousarray(kpts, dtype=np.float32).reshape(-1, 4)
        return k.view(KPT_DTYPE).reshape(-1)

    def bad(self, img, kpts, scale_factor=1.0, nbits=256):
        p, w, h, pitch = _img(img)
        k = self.as_kpts(kpts)
        desc = np.zeros((len(k), nbits // 8), np.uint8)
        self.L.efo_bad_compute(p, w, h, pitch, k.ctypes.data, len(k), scale_factor, nbits, _p(desc, _u8p))
        return desc

    def hashsift_features(self, img, kpts, cropping_scale=1.0):
        p, w, h, pitch = _img(img)
        k = self.as_kpts(kpts)
        resp = np.zeros((len(k), 129), np.float32)
        self.L.efo_hashsift_features(p, w, h, pitch, k.ctypes.data, len(k), cropping_scale, _p(resp, _f32p))
        return resp

    def hashsift_patch(self, img, kpt, cropping_scale=1.0):
        p, w, h, pitch = _img(img)
        k = self.as_kpts(kpt)
        patch = np.zeros((32, 32), np.uint8)
        self.L.efo_hashsift_patch(p, w, h, pitch, k.ctypes.data, cropping_scale, _p(patch, _u8p))
        return patch

    def hashsift_project(self, resp129, nbits=256):
        r = np.ascontiguousarray(resp129, np.float32)
        n = r.shape[0]
        desc = np.zeros((n, nbits // 8), np.uint8); proj = np.zeros((n, nbits), np.float32)
        self.L.efo_hashsift_project(_p(r, _f32p), n, nbits, _p(desc, _u8p), _p(proj, _f32p))
        return desc, proj

    def hashsift(self, img, kpts, cropping_scale=1.0, nbits=256, want_proj=False):
        p, w, h, pitch = _img(img)
        k = self.as_kpts(kpts)
        desc = np.zeros((len(k), nbits // 8), np.uint8)
        proj = np.zeros((len(k), nbits), np.float32) if want_proj else None
        self.L.efo_hashsift_compute(p, w, h, pitch, k.ctypes.data, len(k), cropping_scale, nbits, _p(desc, _u8p),
                                    _p(proj, _f32p) if want_proj else None)
        return (desc, proj) if want_proj else desc


class Reference:
    """The reference's own bad.cpp / hash_sift.cpp (unmodified), built by oracle/Makefile into oracle/_ref."""

    def __init__(self):
        if not REF_SO.exists():
            raise FileNotFoundError(f"{REF_SO} not built (needs /root/reference; run make -C oracle ref)")
        L = C.CDLL(str(REF_SO))
        self.L = L
        L.efref_bad_compute.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_float, C.c_int, _u8p]
        L.efref_hashsift_compute.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_float, C.c_int, _u8p]
        L.efref_hashsift_features.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_float, _f32p]

    @staticmethod
    def available() -> bool:
        return REF_SO.exists()

    def bad(self, img, kpts, scale_factor=1.0, nbits=256):
        p, w, h, pitch = _img(img)
        k = Oracle.as_kpts(kpts)
        desc = np.zeros((len(k), nbits // 8), np.uint8)
        rc = self.L.efref_bad_compute(p, w, h, pitch, k.ctypes.data, len(k), scale_factor, nbits, _p(desc, _u8p))
        assert rc == 0
        return desc

    def hashsift(self, img, kpts, cropping_scale=1.0, nbits=256):
        p, w, h, pitch = _img(img)
        k = Oracle.as_kpts(kpts)
        desc = np.zeros((len(k), nbits // 8), np.uint8)
        rc = self.L.efref_hashsift_compute(p, w, h, pitch, k.ctypes.data, len(k), cropping_scale, nbits, _p(desc, _u8p))
        assert rc == 0
        return desc

    def hashsift_features(self, img, kpts, cropping_scale=1.0):
        p, w, h, pitch = _img(img)
        k = Oracle.as_kpts(kpts)
        resp = np.zeros((len(k), 129), np.float32)
        rc = self.L.efref_hashsift_features(p, w, h, pitch, k.ctypes.data, len(k), cropping_scale, _p(resp, _f32p))
        assert rc == 0
        return resp


REF_CUDA_SO = REF_SO.parent / "libef_ref_cuda.so"


class ReferenceCuda:
    """The reference's own CUDA detector kernels (cuda_fast.cu, cuda_efficient_features.cu, UNMODIFIED), built by oracle/Makefile
    (target ref_cuda) into oracle/_ref/libef_ref_cuda.so against oracle/shim_cuda.  Needs a GPU: `-m gpu` tests only."""

    def __init__(self):
        if not REF_CUDA_SO.exists():
            raise FileNotFoundError(f"{REF_CUDA_SO} not built (needs /root/reference and nvcc; run make -C oracle ref_cuda)")
        L = C.CDLL(str(REF_CUDA_SO))
        self.L = L
        sp = C.POINTER(C.c_short)
        L.efrefcu_fast.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, sp]
        L.efrefcu_responses_angles.argtypes = [_u8p, C.c_int, C.c_int, sp, C.c_int, _f32p, _f32p]
        L.efrefcu_responses_angles.restype = None
        L.efrefcu_nms_limit.argtypes = [sp, _f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, sp, _f32p]
        L.efrefcu_scale.argtypes = [sp, C.c_int, C.c_float, C.c_int, sp, C.POINTER(C.c_int), _f32p]
        L.efrefcu_scale.restype = None
        L.efrefcu_time_detect_levels.argtypes = [C.POINTER(_u8p), C.POINTER(C.c_int), C.POINTER(C.c_int), _f32p, C.POINTER(C.c_int), C.c_int,
                                                 C.c_int, C.c_float, C.c_int, C.POINTER(C.c_int)]
        L.efrefcu_time_detect_levels.restype = C.c_float
        L.efrefcu_time_hashsift.argtypes = [_u8p, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_float, C.c_int, _u8p]
        L.efrefcu_time_hashsift.restype = C.c_float
        L.efrefcu_time_bad.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_float, C.c_int, _u8p]
        L.efrefcu_time_bad.restype = C.c_float

    @staticmethod
    def available() -> bool:
        return REF_CUDA_SO.exists()

    def fast(self, img, threshold=20, border=15, maxpoints=None):
        """createMask + calcKeypoints: n x 2 int16 (x, y) in the kernel's atomic arrival order"""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        cap = int(maxpoints if maxpoints is not None else w * h)
        xy = np.zeros((cap, 2), np.int16)
        n = self.L.efrefcu_fast(_p(img, _u8p), w, h, threshold, border, cap, _p(xy, C.POINTER(C.c_short)))
        return xy[:n].copy()

    def responses_angles(self, img, xy):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        xy = np.ascontiguousarray(xy, np.int16)
        r = np.zeros(len(xy), np.float32); a = np.zeros(len(xy), np.float32)
        self.L.efrefcu_responses_angles(_p(img, _u8p), w, h, _p(xy, C.POINTER(C.c_short)), len(xy), _p(r, _f32p), _p(a, _f32p))
        return r, a

    def nms_limit(self, xy, resp, w, h, radius=15.0, maxpoints=-1):
        xy = np.ascontiguousarray(xy, np.int16); resp = np.ascontiguousarray(resp, np.float32)
        oxy = np.zeros((max(len(xy), 1), 2), np.int16); orr = np.zeros(max(len(xy), 1), np.float32)
        m = self.L.efrefcu_nms_limit(_p(xy, C.POINTER(C.c_short)), _p(resp, _f32p), len(xy), w, h, float(radius), int(maxpoints),
                                     _p(oxy, C.POINTER(C.c_short)), _p(orr, _f32p))
        return oxy[:m].copy(), orr[:m].copy()

    def time_detect_levels(self, levels, scales, quotas, threshold=20, radius=15.0, iters=20):
        """the reference's per-level detector sequence on its own kernels for the level images of one frame:
        (mean ms per frame, keypoints per level)"""
        n = len(levels)
        imgs = [np.ascontiguousarray(a, np.uint8) for a in levels]
        ptrs = (_u8p * n)(*[_p(a, _u8p) for a in imgs])
        ws = (C.c_int * n)(*[a.shape[1] for a in imgs]); hs = (C.c_int * n)(*[a.shape[0] for a in imgs])
        sc = np.ascontiguousarray(scales, np.float32); q = (C.c_int * n)(*[int(v) for v in quotas]); out = (C.c_int * n)()
        ms = self.L.efrefcu_time_detect_levels(ptrs, ws, hs, _p(sc, _f32p), q, n, threshold, float(radius), iters, out)
        return float(ms), [int(v) for v in out]

    def time_hashsift(self, img, kpts4, nbits=512, cropping_scale=1.0, iters=20):
        """the reference's GPU HashSIFT (computePatchSIFTs + cublasSgemm + binarizeDescriptors): (mean ms, descriptors)"""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        k = np.ascontiguousarray(kpts4, np.float32).reshape(-1, 4)
        desc = np.zeros((len(k), nbits // 8), np.uint8)
        ms = self.L.efrefcu_time_hashsift(_p(img, _u8p), w, h, _p(k, _f32p), len(k), nbits, float(cropping_scale), iters, _p(desc, _u8p))
        return float(ms), desc

    def time_bad(self, img, kpts4, nbits=512, scale_factor=1.0, iters=20):
        """the reference's GPU BAD kernel (loadBoxPairParams + computeBAD) on an exact int32 integral image built here: (mean ms, descriptors)"""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        integ = np.zeros((h + 1, w + 1), np.int32)
        integ[1:, 1:] = np.cumsum(np.cumsum(img.astype(np.int64), axis=0), axis=1).astype(np.int32)
        k = np.ascontiguousarray(kpts4, np.float32).reshape(-1, 4)
        desc = np.zeros((len(k), nbits // 8), np.uint8)
        ms = self.L.efrefcu_time_bad(_p(integ, C.POINTER(C.c_int)), w, h, _p(k, _f32p), len(k), nbits, float(scale_factor), iters, _p(desc, _u8p))
        return float(ms), desc

    def scale(self, xy, scale, octave):
        xy = np.ascontiguousarray(xy, np.int16)
        oxy = np.zeros_like(xy); oc = np.zeros(len(xy), np.int32); sz = np.zeros(len(xy), np.float32)
        self.L.efrefcu_scale(_p(xy, C.POINTER(C.c_short)), len(xy), float(scale), int(octave), _p(oxy, C.POINTER(C.c_short)),
                             _p(oc, C.POINTER(C.c_int)), _p(sz, _f32p))
        return oxy, oc, sz


def stress_keypoints(w: int, h: int, n: int, seed: int = 1) -> np.ndarray:
    """Keypoint stress set: uniform positions including the border band,
    special angles, sizes 31..111.  Returns n x 4 float32 (x, y, size, angle)."""
    rng = np.random.default_rng(seed)
    k = np.empty((n, 4), np.float32)
    k[:, 0] = rng.uniform(0, w - 1, n).astype(np.float32)
    k[:, 1] = rng.uniform(0, h - 1, n).astype(np.float32)
    third = n // 3
    k[:third, 0] = np.round(k[:third, 0]); k[:third, 1] = np.round(k[:third, 1])   # integer positions like the detector
    k[:, 2] = np.where(rng.random(n) < 0.5, 31.0, rng.uniform(31, 111, n)).astype(np.float32)
    ang = rng.uniform(0, 360, n).astype(np.float32)
    special = np.array([-1.0, -0.5, 0.0, 90.0, 180.0, 270.0, 359.99, 45.0], np.float32)
    pick = rng.random(n) < 0.25
    ang[pick] = special[rng.integers(0, len(special), pick.sum())]
    k[:, 3] = ang
    return k
