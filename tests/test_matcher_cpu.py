"""CPU tests of the matcher / colour-conversion oracle (oracle/match_oracle.py): against the golden vectors generated with the real
OpenCV (tests/golden/match_golden.npz, tools/make_match_golden.py) and, where cv2 is importable, against cv2 live."""
from pathlib import Path

import numpy as np
import pytest

import match_oracle as mo

GOLD = Path(__file__).resolve().parent / "golden" / "match_golden.npz"
CASES = ["ties32", "rand64", "dups64", "one_train"]


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_opencv_golden(gold, name):
    q, t = gold[f"{name}_q"], gold[f"{name}_t"]
    idx, dist = mo.knn_match(q, t, 2)
    assert np.array_equal(idx, gold[f"{name}_knn_idx"])
    assert np.array_equal(dist, gold[f"{name}_knn_dist"])
    cidx, cdist = mo.cross_check_match(q, t)
    keep = np.nonzero(cidx >= 0)[0]
    got = np.stack([keep, cidx[keep], cdist[keep]], axis=1).astype(np.int32).reshape(-1, 3)
    assert np.array_equal(got, gold[f"{name}_cross"])


def test_gray_oracle_matches_opencv_golden(gold):
    assert np.array_equal(mo.bgr_to_gray(gold["bgr"]), gold["bgr_gray"])
    assert np.array_equal(mo.bgr_to_gray(gold["bgra"]), gold["bgra_gray"])


def test_oracle_against_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for bits, lo in [(32, 3), (64, 256), (64, 2)]:
        q = rng.integers(0, lo, (60, bits), dtype=np.uint8); t = rng.integers(0, lo, (85, bits), dtype=np.uint8)
        knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
        idx, dist = mo.knn_match(q, t, 2)
        assert [[m.trainIdx for m in ms] for ms in knn] == idx.tolist()
        assert [[int(m.distance) for m in ms] for ms in knn] == dist.tolist()
        cm = cv2.BFMatcher(cv2.NORM_HAMMING, True).match(q, t)
        cidx, _ = mo.cross_check_match(q, t)
        assert [(m.queryIdx, m.trainIdx) for m in cm] == [(i, int(j)) for i, j in enumerate(cidx) if j >= 0]
    # every colour: 2^24 BGR triples
    v = np.arange(1 << 24, dtype=np.uint32)
    bgr = np.stack([v & 255, (v >> 8) & 255, v >> 16], axis=1).astype(np.uint8).reshape(4096, 4096, 3)
    assert np.array_equal(mo.bgr_to_gray(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))


def test_ratio_cross_filter_is_the_sample_loop():
    rng = np.random.default_rng(9)
    q = rng.integers(0, 256, (70, 32), dtype=np.uint8)
    t = np.concatenate([q[:40] ^ (rng.integers(0, 256, (40, 32), dtype=np.uint8) & 1), rng.integers(0, 256, (30, 32), dtype=np.uint8)])
    i12, d12 = mo.knn_match(q, t, 2); i21, d21 = mo.knn_match(t, q, 2)
    out = mo.ratio_cross_filter(i12, d12, i21, d21, 0.9)
    exp = []
    for qi in range(len(q)):   # samples/sample_image_sequence.cpp:121-137 written out
        m12 = [(int(i12[qi, k]), float(np.float32(d12[qi, k]))) for k in range(2)]
        m21 = [(int(i21[m12[0][0], k]), float(np.float32(d21[m12[0][0], k]))) for k in range(2)]
        if m12[0][1] > 0.9 * m12[1][1] or m21[0][1] > 0.9 * m21[1][1] or m21[0][0] != qi:
            exp.append(-1)
        else:
            exp.append(m12[0][0])
    assert out.tolist() == exp and sum(e >= 0 for e in exp) >= 30
