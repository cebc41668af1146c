"""Host mirror of the callers either side of the path: cv::BFMatcher(NORM_HAMMING) over the descriptors that
detectAndCompute left on the device, and convertToGray (samples/sample_common.cpp:35-45).  Same method names and argument
meaning as OpenCV; torch CUDA tensors stand in for GpuMat, numpy arrays for Mat.  No CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

NORM_HAMMING = 6
DMATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("distance", "<f4")])


def _lib():
    from . import load_library
    return load_library()


def _torch():
    import torch
    return torch


def _check(rc):
    if rc != 0:
        from . import EfError
        raise EfError(f"status {rc}: {_lib().ef_match_last_error_string().decode()}")


def _as_device_desc(d):
    torch = _torch()
    from . import EfError
    if isinstance(d, np.ndarray):
        if d.dtype != np.uint8 or d.ndim != 2:
            raise EfError("descriptors must be a 2-D uint8 matrix (CV_8U)")
        if not torch.cuda.is_available():
            raise EfError("no CUDA device: the matcher exists only as sm_100a kernels (no CPU fallback)")
        d = torch.from_numpy(np.ascontiguousarray(d)).cuda()
    if not (isinstance(d, torch.Tensor) and d.is_cuda and d.dtype == torch.uint8 and d.dim() == 2 and (d.shape[0] == 0 or d.stride(1) == 1)):
        raise EfError("descriptors must be an N x B uint8 CUDA tensor with unit column stride")
    if d.shape[1] not in (32, 64):
        raise EfError("descriptor size must be 32 or 64 bytes")
    return d


class BFMatcher:
    """cv::BFMatcher for NORM_HAMMING (create(normType, crossCheck), match, knnMatch)."""

    def __init__(self, normType=NORM_HAMMING, crossCheck=False):
        from . import EfError
        if normType != NORM_HAMMING:
            raise EfError("only NORM_HAMMING (defaultNorm() of the path's descriptors) is implemented")
        self.crossCheck = bool(crossCheck)
        self._scratch = {}      # one scratch buffer per (device, stream): concurrent streams never share one

    @staticmethod
    def create(normType=NORM_HAMMING, crossCheck=False) -> "BFMatcher":
        return BFMatcher(normType, crossCheck)

    def _scratch_for(self, nq, nt, device, stream):
        """Scratch of the call, owned per (device, stream).  The kernels run on `stream`, which need not be torch's current stream:
        every tensor they touch is handed to the caching allocator with record_stream(), so that memory freed while the kernels are
        still in flight (a replaced scratch, the results, temporaries made from host descriptors) is not recycled under them."""
        torch = _torch()
        need = int(_lib().ef_match_scratch_bytes(nq, nt))
        st = stream if stream is not None else torch.cuda.current_stream(device)
        key = (device, int(st.cuda_stream))
        sc = self._scratch.get(key)
        if sc is None or sc.numel() < need:
            sc = torch.empty(need, dtype=torch.uint8, device=device)
            self._scratch[key] = sc
        sc.record_stream(st)
        return sc, st

    # ---- device forms: int32 CUDA tensors, no host synchronisation ----
    def knnMatchAsync(self, query, train, k=2, stream=None):
        """-> (idx nq x k, dist nq x k) int32 CUDA tensors; idx -1 where train has fewer than k rows."""
        torch = _torch()
        from . import _stream_ptr
        q, t = _as_device_desc(query), _as_device_desc(train)
        nq, nt = q.shape[0], t.shape[0]
        idx = torch.empty((nq, k), dtype=torch.int32, device=q.device)
        dist = torch.empty((nq, k), dtype=torch.int32, device=q.device)
        if t.device != q.device:
            from . import EfError
            raise EfError("query and train descriptors must live on the same device")
        sc, st = self._scratch_for(nq, nt, q.device, stream)
        for a in (q, t, idx, dist):
            a.record_stream(st)
        _check(_lib().ef_match_knn_async(q.data_ptr(), q.stride(0) if nq else 0, nq, t.data_ptr(), t.stride(0) if nt else 0, nt, q.shape[1], k,
                                         idx.data_ptr(), dist.data_ptr(), sc.data_ptr(), _stream_ptr(st, q.device)))
        return idx, dist

    def matchAsync(self, query, train, stream=None):
        """-> (trainIdx nq, dist nq) int32 CUDA tensors; with crossCheck, trainIdx is -1 for queries without a mutual match."""
        torch = _torch()
        from . import _stream_ptr
        q, t = _as_device_desc(query), _as_device_desc(train)
        nq, nt = q.shape[0], t.shape[0]
        if not self.crossCheck:
            idx, dist = self.knnMatchAsync(q, t, 1, stream)
            return idx[:, 0], dist[:, 0]
        idx = torch.empty(nq, dtype=torch.int32, device=q.device)
        dist = torch.empty(nq, dtype=torch.int32, device=q.device)
        if t.device != q.device:
            from . import EfError
            raise EfError("query and train descriptors must live on the same device")
        sc, st = self._scratch_for(nq, nt, q.device, stream)
        for a in (q, t, idx, dist):
            a.record_stream(st)
        _check(_lib().ef_match_cross_check_async(q.data_ptr(), q.stride(0) if nq else 0, nq, t.data_ptr(), t.stride(0) if nt else 0, nt, q.shape[1],
                                                 idx.data_ptr(), dist.data_ptr(), sc.data_ptr(), _stream_ptr(st, q.device)))
        return idx, dist

    # ---- OpenCV-shaped forms: DMatch records on the host ----
    def match(self, queryDescriptors, trainDescriptors) -> np.ndarray:
        """std::vector<DMatch> as a structured array (queryIdx, trainIdx, distance), in query order; unmatched queries omitted."""
        idx, dist = self.matchAsync(queryDescriptors, trainDescriptors)
        idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
        keep = np.nonzero(idx >= 0)[0]
        out = np.zeros(len(keep), DMATCH_DTYPE)
        out["queryIdx"] = keep; out["trainIdx"] = idx[keep]; out["distance"] = dist[keep]
        return out

    def knnMatch(self, queryDescriptors, trainDescriptors, k=2) -> list:
        """std::vector<std::vector<DMatch>>: one structured array of <= k records per query row."""
        idx, dist = self.knnMatchAsync(queryDescriptors, trainDescriptors, k)
        idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
        out = []
        for q in range(idx.shape[0]):
            m = idx[q] >= 0
            r = np.zeros(int(m.sum()), DMATCH_DTYPE)
            r["queryIdx"] = q; r["trainIdx"] = idx[q][m]; r["distance"] = dist[q][m]
            out.append(r)
        return out


def ratio_cross_filter(idx12, dist12, idx21, dist21, uniqueness=0.9, stream=None):
    """The match filter of samples/sample_image_sequence.cpp:121-137 on the device: knn (k = 2) results of both directions ->
    int32 CUDA tensor, entry q = matched train row or -1."""
    torch = _torch()
    from . import _stream_ptr
    nq, nt = idx12.shape[0], idx21.shape[0]
    for a in (idx12, dist12, idx21, dist21):
        assert a.is_cuda and a.dtype == torch.int32 and a.is_contiguous() and a.dim() == 2 and a.shape[1] == 2
    out = torch.empty(nq, dtype=torch.int32, device=idx12.device)
    st = stream if stream is not None else torch.cuda.current_stream(idx12.device)
    for a in (idx12, dist12, idx21, dist21, out):
        a.record_stream(st)
    _check(_lib().ef_match_ratio_cross_async(idx12.data_ptr(), dist12.data_ptr(), nq, idx21.data_ptr(), dist21.data_ptr(), nt,
                                             float(uniqueness), out.data_ptr(), _stream_ptr(st, idx12.device)))
    return out


def cvtColorToGray(image, stream=None):
    """convertToGray (samples/sample_common.cpp:35-45): H x W (returned as is), H x W x 3 (BGR) or H x W x 4 (BGRA) uint8 CUDA tensor
    -> H x W uint8 CUDA tensor with OpenCV's COLOR_BGR2GRAY arithmetic."""
    torch = _torch()
    from . import EfError, _stream_ptr
    if not (isinstance(image, torch.Tensor) and image.is_cuda and image.dtype == torch.uint8):
        raise EfError("image must be a uint8 CUDA tensor")
    if image.dim() == 2:
        return image
    if image.dim() != 3 or image.shape[2] not in (3, 4) or image.stride(2) != 1 or image.stride(1) != image.shape[2]:
        raise EfError("Image should be 8UC1, 8UC3 or 8UC4")  # CV_Error(StsBadArg, ...), sample_common.cpp:44
    H, W, cn = image.shape
    gray = torch.empty((H, W), dtype=torch.uint8, device=image.device)
    st = stream if stream is not None else torch.cuda.current_stream(image.device)
    image.record_stream(st); gray.record_stream(st)
    rc = _lib().ef_bgr_to_gray_async(image.data_ptr(), image.stride(0), W, H, cn, gray.data_ptr(), gray.stride(0), _stream_ptr(st, image.device))
    if rc != 0:
        raise EfError(f"ef_bgr_to_gray_async failed with status {rc}")
    return gray
