"""efb200 -- Python host side over the C ABI of libef_b200.so (include/ef_b200.h).

Mirrors the reference's operator interface for the detectAndCompute hot path
(modules/cuda_efficient_features/include/cuda_efficient_features.h:28-98 and
cuda_efficient_descriptors.h:27-121): same names, argument meaning and error behaviour, with
torch CUDA tensors standing in for cv::cuda::GpuMat and numpy arrays for cv::Mat.  PyTorch is only
plumbing here (device memory, streams); every computation runs in the hand-written sm_100a kernels
behind the C ABI.  There is no CPU fallback: importing works without a GPU (so that symbols can be
checked), but creating an object raises if the CUDA device or the extension is missing.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG.parent / "libef_b200.so"

MAX_LEVELS = 16
LOCATION_ROW, RESPONSE_ROW, ANGLE_ROW, OCTAVE_ROW, SIZE_ROW, ROWS_COUNT = 0, 1, 2, 3, 4, 5
BAD_256, BAD_512, HASH_SIFT_256, HASH_SIFT_512 = 0, 1, 2, 3
NORM_HAMMING = 6
CV_8U = 0

(PARAM_MAX_FEATURES, PARAM_SCALE_FACTOR, PARAM_NLEVELS, PARAM_FIRST_LEVEL, PARAM_FAST_THRESHOLD,
 PARAM_NONMAX_RADIUS, PARAM_DESCRIPTOR_TYPE, PARAM_DESC_SCALE) = range(8)

EXPORTS = [
    "ef_default_params", "ef_create", "ef_destroy", "ef_set_param", "ef_get_param", "ef_workspace_bytes",
    "ef_descriptor_size", "ef_last_error_string", "ef_version", "ef_detect_and_compute_async",
    "ef_detect_and_compute_batch_async", "ef_compute_async", "ef_compute_rows_async",
    "ef_detect_and_compute_host", "ef_detect_and_compute_host_batch", "ef_debug_level_view",
    "ef_debug_level_counts", "ef_debug_keep_projection", "ef_debug_hashsift_views", "ef_debug_copy_to_host",
    "ef_stage_timing_enable", "ef_stage_times", "ef_kernel_launch_count",
    "ef_mg_create", "ef_mg_destroy", "ef_mg_device_count", "ef_mg_shard_range", "ef_mg_detect_and_compute_host_batch",
    "ef_mg_last_error_string",
    "ef_band_candidate_bytes", "ef_band_detect_async", "ef_band_finish_async", "ef_band_tile_rows", "ef_band_desc_rows",
    "ef_match_scratch_bytes", "ef_match_knn_async", "ef_match_cross_check_async", "ef_match_ratio_cross_async",
    "ef_match_last_error_string", "ef_bgr_to_gray_async", "ef_debug_project_async", "ef_synth_frames_async",
]
STAGE_NAMES = ["pyramid", "score", "nms", "compact", "select", "angle_pack", "blur", "describe", "project"]


class EfError(RuntimeError):
    """Stands in for cv::Exception (CV_Assert / CV_Error in the reference)."""


class ef_params(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int), ("first_level", C.c_int),
                ("fast_threshold", C.c_int), ("nonmax_radius", C.c_int), ("desc_type", C.c_int), ("desc_scale", C.c_float),
                ("max_width", C.c_int), ("max_height", C.c_int), ("max_batch", C.c_int), ("max_keypoints", C.c_int),
                ("device", C.c_int), ("flags", C.c_int)]


FLAG_COMPUTE_ONLY = 1


class ef_level_view(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("scale", C.c_float), ("quota", C.c_int),
                ("d_image", C.c_void_p), ("image_pitch", C.c_size_t), ("d_blurred", C.c_void_p), ("blurred_pitch", C.c_size_t),
                ("d_response", C.c_void_p), ("response_pitch", C.c_size_t)]


_lib = None


def load_library() -> C.CDLL:
    """Load libef_b200.so; raises (loudly) when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise EfError(f"{LIB_PATH} is missing: build it with `make -C {LIB_PATH.parent}` "
                      "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    L = C.CDLL(str(LIB_PATH))
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    L.ef_default_params.argtypes = [C.POINTER(ef_params)]
    L.ef_create.argtypes = [C.POINTER(ef_params), C.POINTER(vp)]
    L.ef_destroy.argtypes = [vp]
    L.ef_destroy.restype = None
    L.ef_set_param.argtypes = [vp, i32, C.c_double]
    L.ef_get_param.argtypes = [vp, i32, C.POINTER(C.c_double)]
    L.ef_workspace_bytes.argtypes = [vp]
    L.ef_workspace_bytes.restype = sz
    L.ef_descriptor_size.argtypes = [vp]
    L.ef_last_error_string.argtypes = [vp]
    L.ef_last_error_string.restype = C.c_char_p
    L.ef_version.restype = C.c_char_p
    L.ef_detect_and_compute_async.argtypes = [vp, vp, sz, i32, i32, vp, sz, vp, sz, vp, vp]
    L.ef_detect_and_compute_batch_async.argtypes = [vp, i32, vp, sz, sz, i32, i32, vp, sz, sz, vp, sz, sz, vp, vp]
    L.ef_compute_async.argtypes = [vp, vp, sz, i32, i32, vp, i32, vp, sz, vp]
    L.ef_compute_rows_async.argtypes = [vp, vp, sz, i32, i32, vp, sz, i32, vp, sz, vp]
    L.ef_detect_and_compute_host.argtypes = [vp, vp, sz, i32, i32, vp, vp, C.POINTER(i32), vp]
    L.ef_detect_and_compute_host_batch.argtypes = [vp, i32, vp, sz, sz, i32, i32, vp, vp, C.POINTER(i32), vp]
    L.ef_debug_level_view.argtypes = [vp, i32, i32, C.POINTER(ef_level_view)]
    L.ef_debug_level_counts.argtypes = [vp, i32, C.POINTER(i32), vp]
    L.ef_debug_keep_projection.argtypes = [vp, i32]
    L.ef_debug_hashsift_views.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.ef_debug_copy_to_host.argtypes = [vp, vp, sz, vp, sz, sz, sz]
    L.ef_stage_timing_enable.argtypes = [vp, i32]
    L.ef_stage_times.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(i32)]
    L.ef_kernel_launch_count.restype = C.c_ulonglong
    L.ef_mg_create.argtypes = [C.POINTER(ef_params), C.POINTER(i32), i32, C.POINTER(vp)]
    L.ef_mg_destroy.argtypes = [vp]
    L.ef_mg_destroy.restype = None
    L.ef_mg_device_count.argtypes = [vp]
    L.ef_mg_shard_range.argtypes = [i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    L.ef_mg_shard_range.restype = None
    L.ef_mg_detect_and_compute_host_batch.argtypes = [vp, i32, vp, sz, sz, i32, i32, vp, vp, C.POINTER(i32)]
    L.ef_mg_last_error_string.argtypes = [vp]
    L.ef_mg_last_error_string.restype = C.c_char_p
    L.ef_band_candidate_bytes.argtypes = [vp]
    L.ef_band_candidate_bytes.restype = sz
    L.ef_band_detect_async.argtypes = [vp, i32, i32, i32, vp, sz, sz, i32, i32, vp, vp]
    L.ef_band_finish_async.argtypes = [vp, i32, i32, i32, vp, vp, sz, sz, vp, sz, sz, vp, vp]
    L.ef_band_tile_rows.argtypes = [i32, i32, i32, i32] + [C.POINTER(i32)] * 4
    L.ef_band_tile_rows.restype = None
    L.ef_band_desc_rows.argtypes = [i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    L.ef_band_desc_rows.restype = None
    L.ef_match_scratch_bytes.argtypes = [i32, i32]
    L.ef_match_scratch_bytes.restype = sz
    L.ef_match_knn_async.argtypes = [vp, sz, i32, vp, sz, i32, i32, i32, vp, vp, vp, vp]
    L.ef_match_cross_check_async.argtypes = [vp, sz, i32, vp, sz, i32, i32, vp, vp, vp, vp]
    L.ef_match_ratio_cross_async.argtypes = [vp, vp, i32, vp, vp, i32, C.c_double, vp, vp]
    L.ef_match_last_error_string.restype = C.c_char_p
    L.ef_bgr_to_gray_async.argtypes = [vp, sz, i32, i32, i32, vp, sz, vp]
    L.ef_debug_project_async.argtypes = [vp, vp, i32, i32, vp, sz, vp]
    L.ef_synth_frames_async.argtypes = [vp, sz, sz, i32, i32, i32, C.c_uint, C.c_uint, vp]
    _lib = L
    return L


def _torch():
    import torch
    return torch


def _desc_bytes(dtype: int) -> int:
    return 32 if dtype in (BAD_256, HASH_SIFT_256) else 64


KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4")])


def _stream_ptr(stream, device=None) -> int:
    """cudaStream_t of `stream`, or of torch's current stream ON `device` (the handle's / the tensors' device, which need
    not be torch's current device in a single-process multi-GPU program)."""
    torch = _torch()
    if stream is None:
        stream = torch.cuda.current_stream(device)
    elif device is not None and stream.device != torch.device(device):
        raise EfError(f"stream belongs to {stream.device}, the call runs on {device}")
    return int(stream.cuda_stream)


class _Handle:
    """Owns one ef_handle (one per concurrent stream, like the reference object)."""

    def __init__(self, **kw):
        L = load_library()
        torch = _torch()
        if not torch.cuda.is_available():
            raise EfError("no CUDA device: the detectAndCompute path exists only as sm_100a kernels (no CPU fallback)")
        p = ef_params()
        L.ef_default_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        if p.device < 0:
            p.device = torch.cuda.current_device()
        self.L = L
        self.params = p
        self.h = C.c_void_p()
        rc = L.ef_create(C.byref(p), C.byref(self.h))
        if rc != 0:
            raise EfError(f"ef_create failed with status {rc}")
        self.device = torch.device("cuda", p.device)

    def check(self, rc: int):
        if rc != 0:
            raise EfError(f"status {rc}: {self.L.ef_last_error_string(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.ef_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, pid, v):
        self.check(self.L.ef_set_param(self.h, pid, float(v)))

    def get(self, pid):
        out = C.c_double()
        self.check(self.L.ef_get_param(self.h, pid, C.byref(out)))
        return out.value


def _check_image_tensor(img, device=None):
    torch = _torch()
    if not (isinstance(img, torch.Tensor) and img.is_cuda and img.dtype == torch.uint8 and img.dim() == 2 and img.stride(1) == 1):
        raise EfError("image must be a 2-D uint8 CUDA tensor with unit column stride (CV_8UC1 GpuMat)")
    if device is not None and img.device != device:
        raise EfError(f"image lives on {img.device}, the object was created on {device} (one object per device)")


class EfficientFeatures:
    """cv::cuda::EfficientFeatures (cuda_efficient_features.h:28-98)."""

    LOCATION_ROW, RESPONSE_ROW, ANGLE_ROW, OCTAVE_ROW, SIZE_ROW, ROWS_COUNT = 0, 1, 2, 3, 4, 5
    BAD_256, BAD_512, HASH_SIFT_256, HASH_SIFT_512 = 0, 1, 2, 3

    def __init__(self, nfeatures=5000, scaleFactor=1.2, nlevels=8, firstLevel=0, fastThreshold=20, nonmaxRadius=15,
                 dtype=HASH_SIFT_256, max_width=3840, max_height=2160, max_batch=1, max_keypoints=0, device=-1, flags=0):
        self._h = _Handle(nfeatures=nfeatures, scale_factor=scaleFactor, nlevels=nlevels, first_level=firstLevel,
                          fast_threshold=fastThreshold, nonmax_radius=nonmaxRadius, desc_type=dtype,
                          max_width=max_width, max_height=max_height, max_batch=max_batch,
                          max_keypoints=max_keypoints, device=device, flags=flags)
        self._out = None

    @staticmethod
    def create(nfeatures=5000, scaleFactor=1.2, nlevels=8, firstLevel=0, fastThreshold=20, nonmaxRadius=15,
               dtype=HASH_SIFT_256, **capacity) -> "EfficientFeatures":
        return EfficientFeatures(nfeatures, scaleFactor, nlevels, firstLevel, fastThreshold, nonmaxRadius, dtype, **capacity)

    # ---- Feature2D-shaped synchronous API (cuda_efficient_features.cpp:197-213) -----------------
    def detect(self, image, mask=None):
        kp = self.detectAsync(image, mask)
        return self.convert(kp)

    def compute(self, image, keypoints):
        """keypoints: structured array from convert() / n x 4 (x, y, size, angle) array (std::vector<KeyPoint> path:
        the caller's size and angle are used, cuda_efficient_features.cpp:116-128)."""
        torch = _torch()
        k = keypoints
        if isinstance(k, np.ndarray) and k.dtype.names:
            k = np.stack([k["x"], k["y"], k["size"], k["angle"]], axis=1)
        k = np.ascontiguousarray(k, np.float32).reshape(-1, 4)
        dev_img = self._as_device_image(image)
        if len(k) == 0:
            return np.zeros((0, self.descriptorSize()), np.uint8)
        dk = torch.from_numpy(k).to(self._h.device)
        desc = torch.empty((len(k), self.descriptorSize()), dtype=torch.uint8, device=self._h.device)
        self._h.check(self._h.L.ef_compute_async(self._h.h, dev_img.data_ptr(), dev_img.stride(0), dev_img.shape[1], dev_img.shape[0],
                                                 dk.data_ptr(), len(k), desc.data_ptr(), desc.stride(0), _stream_ptr(None, self._h.device)))
        out = desc.cpu().numpy() if isinstance(image, np.ndarray) else desc
        return out

    def detectAndCompute(self, image, mask=None, useProvidedKeypoints=False):
        """Host (cv::Mat) or device image; returns (keypoints structured array, descriptors)."""
        if isinstance(image, np.ndarray):
            kp5, desc = self._host_call(image[None], True)
            return self.convert(kp5[0]), desc[0]
        kp, desc = self.detectAndComputeAsync(image, mask, useProvidedKeypoints)
        return self.convert(kp), desc

    # ---- *Async API over device buffers (cuda_efficient_features.h:60,69,72-73) -----------------
    def detectAsync(self, image, mask=None, stream=None):
        kp, _ = self.detectAndComputeAsync(image, mask, False, stream, want_descriptors=False)
        return kp

    def computeAsync(self, image, keypoints, stream=None):
        """keypoints: 5 x N float32 CUDA tensor (GpuMat path: LOCATION and ANGLE rows only, size forced to 31,
        cuda_efficient_features.cu:250-263)."""
        torch = _torch()
        _check_image_tensor(image, self._h.device)
        if not (isinstance(keypoints, torch.Tensor) and keypoints.is_cuda and keypoints.dtype == torch.float32 and keypoints.device == image.device
                and keypoints.dim() == 2 and keypoints.shape[0] == ROWS_COUNT and keypoints.stride(1) == 1):
            raise EfError("keypoints must be a 5 x N float32 CUDA tensor")  # CV_Assert(tmp.rows == 5 && tmp.type() == CV_32F)
        n = keypoints.shape[1]
        desc = torch.empty((n, self.descriptorSize()), dtype=torch.uint8, device=image.device)
        if n == 0:
            return desc
        self._h.check(self._h.L.ef_compute_rows_async(self._h.h, image.data_ptr(), image.stride(0), image.shape[1], image.shape[0],
                                                      keypoints.data_ptr(), keypoints.stride(0) * 4, n, desc.data_ptr(), desc.stride(0),
                                                      _stream_ptr(stream, self._h.device)))
        return desc

    def detectAndComputeAsync(self, image, mask=None, useProvidedKeypoints=False, stream=None, want_descriptors=True):
        """Returns (keypoints 5 x N, descriptors N x B) CUDA tensors.  Like the reference (GpuMat::create to the
        exact size) this needs N on the host, i.e. ONE synchronisation at the end; use
        detectAndComputeRaw() for the fully asynchronous fixed-capacity form."""
        kp, desc, count = self.detectAndComputeRaw(image, mask, useProvidedKeypoints, stream, want_descriptors)
        n = int(count.item())
        return kp[:, :n], (desc[:n] if desc is not None else None)

    def detectAndComputeRaw(self, image, mask=None, useProvidedKeypoints=False, stream=None, want_descriptors=True):
        """Fully asynchronous: returns (5 x nfeatures keypoints, nfeatures x B descriptors, 1-element count),
        all on the device; columns/rows [0, count) are valid.  `mask` is accepted and ignored exactly like the
        reference (cuda_efficient_features.cpp:225-250 never reads it)."""
        torch = _torch()
        if useProvidedKeypoints:
            raise EfError("useProvidedKeypoints must be false")  # CV_Assert(!useProvidedKeypoints), :229
        _check_image_tensor(image, self._h.device)  # CV_Assert(_image.type() == CV_8U), :228
        nf = int(self._h.get(PARAM_MAX_FEATURES))
        kp = torch.empty((ROWS_COUNT, nf), dtype=torch.float32, device=image.device)
        desc = torch.empty((nf, self.descriptorSize()), dtype=torch.uint8, device=image.device) if want_descriptors else None
        count = torch.zeros(1, dtype=torch.int32, device=image.device)
        self._h.check(self._h.L.ef_detect_and_compute_async(
            self._h.h, image.data_ptr(), image.stride(0), image.shape[1], image.shape[0],
            kp.data_ptr(), kp.stride(0) * 4, desc.data_ptr() if desc is not None else None,
            desc.stride(0) if desc is not None else 0, count.data_ptr(), _stream_ptr(stream, self._h.device)))
        return kp, desc, count

    def detectAndComputeBatchRaw(self, images, stream=None, want_descriptors=True, out=None):
        """New (not in the reference): F x H x W uint8 CUDA tensor -> (F x 5 x nfeatures, F x nfeatures x B, F counts)."""
        torch = _torch()
        if not (isinstance(images, torch.Tensor) and images.is_cuda and images.dtype == torch.uint8 and images.dim() == 3 and images.stride(2) == 1
                and images.device == self._h.device):
            raise EfError(f"images must be an F x H x W uint8 CUDA tensor on {self._h.device}")
        F, H, W = images.shape
        nf = int(self._h.get(PARAM_MAX_FEATURES))
        if out is None:
            kp = torch.empty((F, ROWS_COUNT, nf), dtype=torch.float32, device=images.device)
            desc = torch.empty((F, nf, self.descriptorSize()), dtype=torch.uint8, device=images.device) if want_descriptors else None
            counts = torch.zeros(F, dtype=torch.int32, device=images.device)
        else:
            kp, desc, counts = out
        self._h.check(self._h.L.ef_detect_and_compute_batch_async(
            self._h.h, F, images.data_ptr(), images.stride(0), images.stride(1), W, H,
            kp.data_ptr(), kp.stride(0) * 4, kp.stride(1) * 4,
            desc.data_ptr() if desc is not None else None, desc.stride(0) if desc is not None else 0,
            desc.stride(1) if desc is not None else 0, counts.data_ptr(), _stream_ptr(stream, self._h.device)))
        return kp, desc, counts

    # ---- one oversized frame over several GPUs (ef_band_*; see efb200/tiling.py for the collective plumbing) ----
    def bandCandidateBytes(self) -> int:
        return int(self._h.L.ef_band_candidate_bytes(self._h.h))

    def bandDetect(self, images, shard: int, nshards: int, stream=None):
        """Phase 1 on band `shard` of `nshards`: pyramid, FAST/Harris + radius NMS on the band, local top-quota.
        images: F x H x W uint8 CUDA tensor (the WHOLE frame on every GPU).  Returns the packed candidates
        (F x bandCandidateBytes() uint8 CUDA tensor) to be all-gathered in shard order."""
        torch = _torch()
        if not (isinstance(images, torch.Tensor) and images.is_cuda and images.dtype == torch.uint8 and images.dim() == 3 and images.stride(2) == 1):
            raise EfError("images must be an F x H x W uint8 CUDA tensor")
        F, H, W = images.shape
        cand = torch.empty((F, self.bandCandidateBytes()), dtype=torch.uint8, device=images.device)
        self._h.check(self._h.L.ef_band_detect_async(self._h.h, shard, nshards, F, images.data_ptr(), images.stride(0), images.stride(1),
                                                     W, H, cand.data_ptr(), _stream_ptr(stream, self._h.device)))
        self._band_images = images  # level 0 of the pyramid aliases the caller's image until bandFinish
        return cand

    def bandFinish(self, all_cand, shard: int, nshards: int, stream=None, want_descriptors=True, out=None):
        """Phase 2: all_cand = nshards x F x bandCandidateBytes() (all-gathered).  Returns (F x 5 x nfeatures keypoints --
        complete and identical on every GPU --, F x nfeatures x B descriptors of which only this GPU's block of output rows
        (efb200.tiling.band_desc_rows) is filled, F counts)."""
        torch = _torch()
        if not (isinstance(all_cand, torch.Tensor) and all_cand.is_cuda and all_cand.dtype == torch.uint8 and all_cand.is_contiguous()
                and all_cand.dim() == 3 and all_cand.shape[0] == nshards and all_cand.shape[2] == self.bandCandidateBytes()):
            raise EfError("all_cand must be a contiguous nshards x F x bandCandidateBytes() uint8 CUDA tensor")
        F = all_cand.shape[1]
        nf = int(self._h.get(PARAM_MAX_FEATURES))
        if out is None:
            kp = torch.empty((F, ROWS_COUNT, nf), dtype=torch.float32, device=all_cand.device)
            desc = torch.empty((F, nf, self.descriptorSize()), dtype=torch.uint8, device=all_cand.device) if want_descriptors else None
            counts = torch.zeros(F, dtype=torch.int32, device=all_cand.device)
        else:
            kp, desc, counts = out
        self._h.check(self._h.L.ef_band_finish_async(
            self._h.h, shard, nshards, F, all_cand.data_ptr(), kp.data_ptr(), kp.stride(0) * 4, kp.stride(1) * 4,
            desc.data_ptr() if desc is not None else None, desc.stride(0) if desc is not None else 0,
            desc.stride(1) if desc is not None else 0, counts.data_ptr(), _stream_ptr(stream, self._h.device)))
        return kp, desc, counts

    def _host_call(self, images: np.ndarray, want_desc: bool):
        """cv::Mat path: host frames in, host results out, H2D/D2H inside (ef_detect_and_compute_host_batch)."""
        if images.dtype != np.uint8 or images.ndim != 3 or images.strides[2] != 1:
            raise EfError("image must be uint8 (CV_8UC1)")
        F, H, W = images.shape
        nf = int(self._h.get(PARAM_MAX_FEATURES))
        db = self.descriptorSize()
        kp = np.zeros((F, ROWS_COUNT, nf), np.float32)
        desc = np.zeros((F, nf, db), np.uint8) if want_desc else None
        counts = (C.c_int * F)()
        self._h.check(self._h.L.ef_detect_and_compute_host_batch(
            self._h.h, F, images.ctypes.data, images.strides[0], images.strides[1], W, H, kp.ctypes.data,
            desc.ctypes.data if want_desc else None, counts, _stream_ptr(None, self._h.device)))
        kps = [kp[f][:, :counts[f]] for f in range(F)]
        descs = [desc[f][:counts[f]] if want_desc else None for f in range(F)]
        return kps, descs

    def _as_device_image(self, image):
        torch = _torch()
        if isinstance(image, np.ndarray):
            if image.dtype != np.uint8 or image.ndim != 2:
                raise EfError("image must be uint8 (CV_8UC1)")
            return torch.from_numpy(np.ascontiguousarray(image)).to(self._h.device)
        _check_image_tensor(image, self._h.device)
        return image

    # ---- convert (cuda_efficient_features.cpp:323-349) --------------------------------------------
    @staticmethod
    def convert(gpu_keypoints) -> np.ndarray:
        """5 x N keypoint matrix (CUDA tensor or numpy) -> structured array of KeyPoint fields."""
        k = gpu_keypoints
        if not isinstance(k, np.ndarray):
            k = k.detach().cpu().numpy()
        k = np.ascontiguousarray(k, np.float32)
        n = k.shape[1]
        out = np.zeros(n, KEYPOINT_DTYPE)
        loc = k[LOCATION_ROW].view(np.int16).reshape(n, 2)
        out["x"] = loc[:, 0]; out["y"] = loc[:, 1]
        out["response"] = k[RESPONSE_ROW]; out["angle"] = k[ANGLE_ROW]
        out["octave"] = k[OCTAVE_ROW].view(np.int32); out["size"] = k[SIZE_ROW]
        return out

    # ---- descriptor info (cuda_efficient_features.cpp:351-353) ------------------------------------
    def descriptorSize(self) -> int:
        return int(self._h.L.ef_descriptor_size(self._h.h))

    def descriptorType(self) -> int:
        return CV_8U

    def defaultNorm(self) -> int:
        return NORM_HAMMING

    # ---- the 7 setter/getter pairs (cuda_efficient_features.h:78-97) ------------------------------
    def setMaxFeatures(self, v): self._h.set(PARAM_MAX_FEATURES, v)
    def getMaxFeatures(self): return int(self._h.get(PARAM_MAX_FEATURES))
    def setScaleFactor(self, v): self._h.set(PARAM_SCALE_FACTOR, v)
    def getScaleFactor(self): return float(np.float32(self._h.get(PARAM_SCALE_FACTOR)))
    def setNLevels(self, v): self._h.set(PARAM_NLEVELS, v)
    def getNLevels(self): return int(self._h.get(PARAM_NLEVELS))
    def setFirstLevel(self, v): self._h.set(PARAM_FIRST_LEVEL, v)
    def getFirstLevel(self): return int(self._h.get(PARAM_FIRST_LEVEL))
    def setFastThreshold(self, v): self._h.set(PARAM_FAST_THRESHOLD, v)
    def getFastThreshold(self): return int(self._h.get(PARAM_FAST_THRESHOLD))
    def setNonmaxRadius(self, v): self._h.set(PARAM_NONMAX_RADIUS, v)
    def getNonmaxRadius(self): return int(self._h.get(PARAM_NONMAX_RADIUS))
    def setDescriptorType(self, v): self._h.set(PARAM_DESCRIPTOR_TYPE, v)
    def getDescriptorType(self): return int(self._h.get(PARAM_DESCRIPTOR_TYPE))

    # ---- measurement support ------------------------------------------------------------------------
    def stageTimingEnable(self, enable=True):
        self._h.check(self._h.L.ef_stage_timing_enable(self._h.h, int(enable)))

    def stageTimes(self):
        """(dict stage -> summed ms since the last call, ncalls); synchronises."""
        ms = (C.c_float * len(STAGE_NAMES))()
        n = C.c_int()
        self._h.check(self._h.L.ef_stage_times(self._h.h, ms, C.byref(n)))
        return {k: float(v) for k, v in zip(STAGE_NAMES, ms)}, int(n.value)

    @staticmethod
    def kernelLaunchCount() -> int:
        return int(load_library().ef_kernel_launch_count())

    def hostBatch(self, frames: np.ndarray, want_descriptors=True):
        """cv::Mat-style batched call: HOST frames (F x H x W uint8) in, host keypoints/descriptors out; the H2D
        and D2H copies happen inside (ef_detect_and_compute_host_batch)."""
        return self._host_call(frames, want_descriptors)

    def hostBatchInto(self, frames, kp_out, desc_out, stream=None):
        """Same as hostBatch() but with caller-provided (ideally pinned) HOST buffers, no allocation:
        frames F x H x W uint8, kp_out F x 5 x nfeatures float32, desc_out F x nfeatures x B uint8 (or None).
        Accepts numpy arrays or CPU torch tensors.  Returns the list of per-frame counts."""
        def ptr(a):
            return a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr()
        F, H, W = frames.shape
        st = frames.strides if isinstance(frames, np.ndarray) else tuple(frames.stride())
        counts = (C.c_int * F)()
        self._h.check(self._h.L.ef_detect_and_compute_host_batch(
            self._h.h, F, ptr(frames), st[0], st[1], W, H, ptr(kp_out), ptr(desc_out) if desc_out is not None else None,
            counts, _stream_ptr(stream, self._h.device)))
        return list(counts)

    def workspaceBytes(self) -> int:
        return int(self._h.L.ef_workspace_bytes(self._h.h))

    # ---- stage introspection for parity tests -----------------------------------------------------
    def debugLevel(self, level: int, frame: int = 0):
        """(image, blurred, response) of a pyramid level of the last call as CUDA tensors (copies)."""
        v = ef_level_view()
        self._h.check(self._h.L.ef_debug_level_view(self._h.h, frame, level, C.byref(v)))
        return v

    def _copy_2d(self, ptr, pitch, width, height, elem, dtype):
        out = np.empty((height, width), dtype)
        self._h.check(self._h.L.ef_debug_copy_to_host(self._h.h, out.ctypes.data, width * elem, ptr, pitch, width * elem, height))
        return out

    def debugLevelArrays(self, level: int, frame: int = 0, want=("image", "blurred", "response")):
        v = self.debugLevel(level, frame)
        out = {"width": v.width, "height": v.height, "scale": np.float32(v.scale), "quota": v.quota}
        if "image" in want:
            out["image"] = self._copy_2d(v.d_image, v.image_pitch, v.width, v.height, 1, np.uint8)
        if "blurred" in want:
            out["blurred"] = self._copy_2d(v.d_blurred, v.blurred_pitch, v.width, v.height, 1, np.uint8)
        if "response" in want:
            out["response"] = self._copy_2d(v.d_response, v.response_pitch * 4, v.width, v.height, 4, np.float32)
        return out

    def debugLevelCounts(self, frame: int = 0) -> np.ndarray:
        n = self.getNLevels()
        buf = (C.c_int * (3 * n))()
        self._h.check(self._h.L.ef_debug_level_counts(self._h.h, frame, buf, _stream_ptr(None, self._h.device)))
        return np.array(list(buf)).reshape(n, 3)

    def debugProject(self, sift128, path=0):
        """The projection stage alone: n x 128 uint8 CUDA tensor -> n x descriptorSize() bits; path 1 = tcgen05, 2 = mma.sync, 3 = fp64."""
        torch = _torch()
        assert sift128.is_cuda and sift128.dtype == torch.uint8 and sift128.dim() == 2 and sift128.shape[1] == 128 and sift128.is_contiguous()
        desc = torch.empty((sift128.shape[0], self.descriptorSize()), dtype=torch.uint8, device=sift128.device)
        self._h.check(self._h.L.ef_debug_project_async(self._h.h, sift128.data_ptr(), sift128.shape[0], path, desc.data_ptr(), desc.stride(0), _stream_ptr(None, self._h.device)))
        return desc

    def debugKeepProjection(self, keep=True):
        self._h.check(self._h.L.ef_debug_keep_projection(self._h.h, int(keep)))

    def debugHashSift(self, n: int):
        a, b = C.c_void_p(), C.c_void_p()
        self._h.check(self._h.L.ef_debug_hashsift_views(self._h.h, C.byref(a), C.byref(b)))
        nbits = self.descriptorSize() * 8
        sift = self._copy_2d(a.value, 128, 128, n, 1, np.uint8)
        proj = self._copy_2d(b.value, nbits * 4, nbits, n, 4, np.float32) if b.value else None
        return sift, proj


class MultiGpuEfficientFeatures:
    """Single-process multi-GPU driver (ef_mg_*): one handle, host thread and stream per device, frames sharded in contiguous
    blocks, no cross-GPU exchange on the data path.  Host (numpy / pinned torch) buffers in and out."""

    def __init__(self, devices=None, nfeatures=5000, scaleFactor=1.2, nlevels=8, firstLevel=0, fastThreshold=20, nonmaxRadius=15,
                 dtype=HASH_SIFT_256, max_width=3840, max_height=2160, max_batch=4):
        L = load_library()
        torch = _torch()
        if not torch.cuda.is_available():
            raise EfError("no CUDA device: the detectAndCompute path exists only as sm_100a kernels (no CPU fallback)")
        devices = list(range(torch.cuda.device_count())) if devices is None else list(devices)
        p = ef_params()
        L.ef_default_params(C.byref(p))
        p.nfeatures, p.scale_factor, p.nlevels, p.first_level = nfeatures, scaleFactor, nlevels, firstLevel
        p.fast_threshold, p.nonmax_radius, p.desc_type = fastThreshold, nonmaxRadius, dtype
        p.max_width, p.max_height, p.max_batch = max_width, max_height, max_batch
        self.L, self.nfeatures, self.desc_bytes, self.devices = L, nfeatures, _desc_bytes(dtype), devices
        self.h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        rc = L.ef_mg_create(C.byref(p), arr, len(devices), C.byref(self.h))
        if rc != 0:
            raise EfError(f"ef_mg_create failed with status {rc}")

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.ef_mg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def detectAndComputeHost(self, frames: np.ndarray, want_descriptors=True):
        """frames: F x H x W uint8 (host).  Returns (list of 5 x N_f keypoint matrices, list of N_f x B descriptors)."""
        if frames.dtype != np.uint8 or frames.ndim != 3 or frames.strides[2] != 1:
            raise EfError("frames must be F x H x W uint8")
        F, H, W = frames.shape
        kp = np.zeros((F, ROWS_COUNT, self.nfeatures), np.float32)
        desc = np.zeros((F, self.nfeatures, self.desc_bytes), np.uint8) if want_descriptors else None
        counts = (C.c_int * F)()
        rc = self.L.ef_mg_detect_and_compute_host_batch(self.h, F, frames.ctypes.data, frames.strides[0], frames.strides[1], W, H,
                                                        kp.ctypes.data, desc.ctypes.data if want_descriptors else None, counts)
        if rc != 0:
            raise EfError(f"status {rc}: {self.L.ef_mg_last_error_string(self.h).decode()}")
        return [kp[f][:, :counts[f]] for f in range(F)], [desc[f][:counts[f]] if want_descriptors else None for f in range(F)]


class _Describer:
    """cv::cuda::EfficientDescriptorsAsync (cuda_efficient_descriptors.h:27-57)."""

    SIZE_512_BITS, SIZE_256_BITS = 100, 101

    def __init__(self, dtype, scale, max_width, max_height, max_keypoints, device):
        # compute-only handle: tables and per-keypoint scratch, no detection workspace (the reference describers hold only their tables)
        self._ef = EfficientFeatures(nfeatures=1, dtype=dtype, max_width=max_width, max_height=max_height,
                                     max_keypoints=max_keypoints, device=device, flags=FLAG_COMPUTE_ONLY)
        self._ef._h.set(PARAM_DESC_SCALE, scale)

    def compute(self, image, keypoints):
        return self._ef.compute(image, keypoints)

    def computeAsync(self, image, keypoints, stream=None):
        return self._ef.computeAsync(image, keypoints, stream)

    def descriptorSize(self): return self._ef.descriptorSize()
    def descriptorType(self): return CV_8U
    def defaultNorm(self): return NORM_HAMMING


class BAD(_Describer):
    """cv::cuda::BAD (cuda_efficient_descriptors.h:67-90)."""

    @staticmethod
    def create(scaleFactor, nbits=101, max_width=3840, max_height=2160, max_keypoints=100000, device=-1) -> "BAD":
        if nbits not in (100, 101):
            raise EfError("n_bits should be either SIZE_512_BITS or SIZE_256_BITS")
        return BAD(BAD_512 if nbits == 100 else BAD_256, scaleFactor, max_width, max_height, max_keypoints, device)


class HashSIFT(_Describer):
    """cv::cuda::HashSIFT (cuda_efficient_descriptors.h:101-121)."""

    @staticmethod
    def create(croppingScale, nbits=101, max_width=3840, max_height=2160, max_keypoints=100000, device=-1) -> "HashSIFT":
        if nbits not in (100, 101):
            raise EfError("n_bits should be either SIZE_512_BITS or SIZE_256_BITS")
        return HashSIFT(HASH_SIFT_512 if nbits == 100 else HASH_SIFT_256, croppingScale, max_width, max_height, max_keypoints, device)


def synth_frames(nframes: int, height: int, width: int, seed: int, first_frame: int = 0, device=None, out=None, stream=None):
    """Synthetic uniform-noise frames of the benchmark generated ON THE DEVICE (ef_synth_frames_async): the pinned counter-based
    generator SURVEY 8d prescribes, bit-identical to the CPU arm's (oracle synth_frame(seed, frame, w, h)).  -> F x H x W uint8."""
    torch = _torch()
    L = load_library()
    if out is None:
        out = torch.empty((nframes, height, width), dtype=torch.uint8, device=device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    if not (out.is_cuda and out.dtype == torch.uint8 and out.dim() == 3 and out.stride(2) == 1 and tuple(out.shape) == (nframes, height, width)):
        raise EfError("out must be an F x H x W uint8 CUDA tensor")
    with torch.cuda.device(out.device):
        rc = L.ef_synth_frames_async(out.data_ptr(), out.stride(1), out.stride(0), width, height, nframes, seed & 0xffffffff, first_frame,
                                     _stream_ptr(stream, out.device))
    if rc != 0:
        raise EfError(f"ef_synth_frames_async failed with status {rc}")
    return out


# ---- callers either side of the path: matcher and colour conversion (efb200/matching.py) ----
from .matching import BFMatcher, DMATCH_DTYPE, cvtColorToGray, ratio_cross_filter  # noqa: E402,F401
