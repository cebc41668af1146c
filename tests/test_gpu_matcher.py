"""GPU tests of the Hamming matcher and the colour conversion (ef_match_*, ef_bgr_to_gray_async) against the oracle pinned to
OpenCV (oracle/match_oracle.py, tests/test_matcher_cpu.py), through the C ABI."""
from pathlib import Path

import numpy as np
import pytest

import match_oracle as mo

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden" / "match_golden.npz"


@pytest.mark.parametrize("name", ["ties32", "rand64", "dups64", "one_train"])
def test_matcher_equals_opencv_golden(name):
    import efb200
    g = np.load(GOLD)
    q, t = g[f"{name}_q"], g[f"{name}_t"]
    bf = efb200.BFMatcher.create(efb200.NORM_HAMMING)
    idx, dist = bf.knnMatchAsync(q, t, 2)
    assert np.array_equal(idx.cpu().numpy(), g[f"{name}_knn_idx"])
    valid = g[f"{name}_knn_idx"] >= 0
    assert np.array_equal(dist.cpu().numpy()[valid], g[f"{name}_knn_dist"][valid])
    m = efb200.BFMatcher.create(efb200.NORM_HAMMING, True).match(q, t)
    got = np.stack([m["queryIdx"], m["trainIdx"], m["distance"].astype(np.int32)], axis=1).reshape(-1, 3)
    assert np.array_equal(got, g[f"{name}_cross"])
    knn = bf.knnMatch(q, t, 2)
    assert [len(r) for r in knn] == valid.sum(1).tolist()


@pytest.mark.parametrize("nq,nt,nbytes,lo", [(1000, 1300, 64, 256), (777, 2500, 32, 4), (3, 5000, 64, 256), (4100, 129, 32, 256), (300, 300, 64, 2)])
def test_matcher_equals_oracle(nq, nt, nbytes, lo):
    import torch
    import efb200
    rng = np.random.default_rng(nq + nt)
    q = rng.integers(0, lo, (nq, nbytes), dtype=np.uint8); t = rng.integers(0, lo, (nt, nbytes), dtype=np.uint8)
    dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    bf = efb200.BFMatcher.create()
    i12, d12 = bf.knnMatchAsync(dq, dt, 2); i21, d21 = bf.knnMatchAsync(dt, dq, 2)
    oi12, od12 = mo.knn_match(q, t, 2); oi21, od21 = mo.knn_match(t, q, 2)
    assert np.array_equal(i12.cpu().numpy(), oi12) and np.array_equal(d12.cpu().numpy(), od12)
    assert np.array_equal(i21.cpu().numpy(), oi21) and np.array_equal(d21.cpu().numpy(), od21)
    i1, d1 = bf.matchAsync(dq, dt)
    assert np.array_equal(i1.cpu().numpy(), oi12[:, 0]) and np.array_equal(d1.cpu().numpy(), od12[:, 0])
    ci, cd = efb200.BFMatcher.create(efb200.NORM_HAMMING, True).matchAsync(dq, dt)
    oci, ocd = mo.cross_check_match(q, t)
    assert np.array_equal(ci.cpu().numpy(), oci)
    assert np.array_equal(cd.cpu().numpy()[oci >= 0], ocd[oci >= 0])
    f = efb200.ratio_cross_filter(i12, d12, i21, d21, 0.9)
    assert np.array_equal(f.cpu().numpy(), mo.ratio_cross_filter(oi12, od12, oi21, od21, 0.9))


def test_matcher_pitched_unaligned_and_empty():
    import torch
    import efb200
    rng = np.random.default_rng(3)
    q = rng.integers(0, 256, (200, 64), dtype=np.uint8); t = rng.integers(0, 256, (333, 64), dtype=np.uint8)
    bq = torch.zeros((200, 100), dtype=torch.uint8).cuda(); bt = torch.zeros((333, 71), dtype=torch.uint8).cuda()
    bq[:, 3:67] = torch.from_numpy(q).cuda(); bt[:, 5:69] = torch.from_numpy(t).cuda()
    bf = efb200.BFMatcher.create()
    idx, dist = bf.knnMatchAsync(bq[:, 3:67], bt[:, 5:69], 2)     # rows neither 16-byte aligned nor densely packed
    oi, od = mo.knn_match(q, t, 2)
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(dist.cpu().numpy(), od)
    e = torch.zeros((0, 64), dtype=torch.uint8).cuda()
    idx, _ = bf.knnMatchAsync(bq[:, 3:67], e, 2)
    assert (idx.cpu().numpy() == -1).all() and len(bf.match(q, np.zeros((0, 64), np.uint8))) == 0
    assert bf.knnMatchAsync(e, bt[:, 5:69], 2)[0].shape == (0, 2)
    with pytest.raises(efb200.EfError):
        bf.knnMatchAsync(torch.zeros((4, 48), dtype=torch.uint8).cuda(), bt[:, 5:69])


def test_matcher_full_size_properties():
    """40k x 40k x 512 bit (the BASELINE.json keypoint budget): permutation recovery, symmetry, distance of the reported pair."""
    import torch
    import efb200
    g = torch.Generator(device="cpu").manual_seed(11)
    t = torch.randint(0, 256, (40000, 64), dtype=torch.uint8, generator=g).cuda()
    perm = torch.randperm(40000, generator=g).cuda()
    q = t[perm].clone()
    q[:, 0] ^= 1                                  # distance 1 to its source row, ~256 to every other row
    bf = efb200.BFMatcher.create()
    idx, dist = bf.knnMatchAsync(q, t, 2)
    assert torch.equal(idx[:, 0].long(), perm) and bool((dist[:, 0] == 1).all()) and bool((dist[:, 1] > 150).all())
    j = idx[:, 1].long()
    d = (torch.bitwise_xor(q, t[j]).view(torch.int32)).cpu().numpy().view(np.uint32)
    pop = np.unpackbits(d.view(np.uint8), axis=1).sum(1)
    assert np.array_equal(pop, dist[:, 1].cpu().numpy())
    ci, _ = efb200.BFMatcher.create(efb200.NORM_HAMMING, True).matchAsync(q, t)
    assert torch.equal(ci.long(), perm)


def test_bgr_to_gray():
    import torch
    import efb200
    g = np.load(GOLD)
    for k in ("bgr", "bgra"):
        out = efb200.cvtColorToGray(torch.from_numpy(g[k]).cuda())
        assert np.array_equal(out.cpu().numpy(), g[f"{k}_gray"])
    rng = np.random.default_rng(1)
    for (h, w, cn) in [(2160, 3840, 3), (101, 67, 3), (64, 130, 4), (5, 3, 3)]:
        img = rng.integers(0, 256, (h, w, cn), dtype=np.uint8)
        assert np.array_equal(efb200.cvtColorToGray(torch.from_numpy(img).cuda()).cpu().numpy(), mo.bgr_to_gray(img))
    gray = torch.zeros((8, 8), dtype=torch.uint8).cuda()
    assert efb200.cvtColorToGray(gray) is gray
    with pytest.raises(efb200.EfError):
        efb200.cvtColorToGray(torch.zeros((8, 8, 2), dtype=torch.uint8).cuda())
