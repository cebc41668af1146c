"""CPU oracle of the callers either side of the path (TEST INFRASTRUCTURE ONLY: imported by tests/ and nothing else).

  knn_match / cross_check_match   cv::BFMatcher(NORM_HAMMING) as called in samples/sample_image_sequence.cpp:115-116 and
                                  samples/sample_feature_matching.cpp:99-101.  cv::BFMatcher is third-party (OpenCV >= 4.6,
                                  features2d matchers.cpp + core batch_distance.cpp, not vendored under /root/reference); its
                                  algorithm is restated here and PINNED against the real OpenCV 4.13 (`cv2`) in
                                  tests/test_matcher_cpu.py and by the committed golden vectors tests/golden/match_golden.npz
                                  (tools/make_match_golden.py).
  ratio_cross_filter              the loop of samples/sample_image_sequence.cpp:121-137.
  bgr_to_gray                     cv::cvtColor(COLOR_BGR2GRAY) for CV_8U (samples/sample_common.cpp:39-42): OpenCV's 15-bit fixed
                                  point, checked against cv2 on all 2^24 colours.
"""
import numpy as np

INT_MAX = 2**31 - 1


def hamming_matrix(q, t):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    lut = np.unpackbits(np.arange(256, dtype=np.uint8)[:, None], axis=1).sum(1).astype(np.int32)
    out = np.zeros((len(q), len(t)), np.int32)
    for i in range(len(q)):            # row at a time: bounded memory
        out[i] = lut[q[i][None, :] ^ t].sum(1)
    return out


def knn_match(q, t, k=2):
    """(idx nq x k, dist nq x k): the k lexicographically smallest (distance, trainIdx); -1 / INT_MAX where nt < k
    (batchDistance scans the train rows in order and replaces on strict <)."""
    D = hamming_matrix(q, t)
    nq, nt = D.shape
    idx = np.full((nq, k), -1, np.int32); dist = np.full((nq, k), INT_MAX, np.int32)
    for i in range(nq):
        order = np.lexsort((np.arange(nt), D[i]))[:k]
        idx[i, :len(order)] = order; dist[i, :len(order)] = D[i, order]
    return idx, dist


def cross_check_match(q, t):
    """trainIdx per query or -1: j = nearest train of i and i = nearest query of j, ties to the lower index on both sides."""
    D = hamming_matrix(q, t)
    nq, nt = D.shape
    idx = np.full(nq, -1, np.int32); dist = np.full(nq, INT_MAX, np.int32)
    if nq == 0 or nt == 0:
        return idx, dist
    fwd = np.array([np.lexsort((np.arange(nt), D[i]))[0] for i in range(nq)])
    bwd = np.array([np.lexsort((np.arange(nq), D[:, j]))[0] for j in range(nt)])
    for i in range(nq):
        if bwd[fwd[i]] == i:
            idx[i] = fwd[i]; dist[i] = D[i, fwd[i]]
    return idx, dist


def ratio_cross_filter(idx12, dist12, idx21, dist21, uniqueness=0.9):
    out = np.full(len(idx12), -1, np.int32)
    for q in range(len(idx12)):
        t = idx12[q, 0]
        if t < 0 or idx12[q, 1] < 0 or idx21[t, 1] < 0:
            continue
        if float(np.float32(dist12[q, 0])) > uniqueness * float(np.float32(dist12[q, 1])):
            continue
        if float(np.float32(dist21[t, 0])) > uniqueness * float(np.float32(dist21[t, 1])):
            continue
        if idx21[t, 0] != q:
            continue
        out[q] = t
    return out


def bgr_to_gray(img):
    a = img.astype(np.uint32)
    return ((a[..., 0] * 3735 + a[..., 1] * 19235 + a[..., 2] * 9798 + (1 << 14)) >> 15).astype(np.uint8)
