"""One oversized frame cut into horizontal bands over several GPUs (ef_band_*): the band-sharded result must be
bit-identical to the single-GPU result (which the parity tests pin against the oracle).  The collectives are emulated on one
GPU here (concatenation = both all-gathers); tests/test_multirank_cpu.py covers the host partition and the collectives under
gloo, and bench.py / tools/tiled_parity_check.py exercise the real NCCL path."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _run(torch, oracle, w, h, nfeat, dtype_name, nshards, frames=1, allow_empty=False, **kw):
    import efb200
    from efb200 import tiling
    dt = getattr(efb200, dtype_name)
    imgs = np.stack([oracle.synth_frame(util.SEED + 77 + nshards, f, w, h) for f in range(frames)])
    d = torch.from_numpy(imgs).cuda()
    single = efb200.EfficientFeatures.create(nfeatures=nfeat, dtype=dt, max_width=w, max_height=h, max_batch=frames, **kw)
    kp0, desc0, cnt0 = single.detectAndComputeBatchRaw(d)
    torch.cuda.synchronize()
    counts0 = np.stack([single.debugLevelCounts(f) for f in range(frames)])
    efs = [efb200.EfficientFeatures.create(nfeatures=nfeat, dtype=dt, max_width=w, max_height=h, max_batch=frames, **kw) for _ in range(nshards)]
    kp, desc, cnt, outs = tiling.detect_and_compute_tiled_emulated(efs, d)
    torch.cuda.synchronize()
    assert np.array_equal(cnt.cpu().numpy(), cnt0.cpu().numpy())
    for f in range(frames):
        n = int(cnt0[f])
        assert n > 0 or allow_empty
        for g, o in enumerate(outs):   # the keypoint matrix is complete and identical on every band owner
            assert np.array_equal(o[0][f, :, :n].cpu().numpy().view(np.uint32), kp0[f, :, :n].cpu().numpy().view(np.uint32)), f"keypoints differ on shard {g}"
            assert np.array_equal(efs[g].debugLevelCounts(f), counts0[f]), f"per-level counts differ on shard {g}"
        a, b = desc[f, :n].cpu().numpy(), desc0[f, :n].cpu().numpy()
        assert np.array_equal(a, b), f"{(a != b).any(axis=1).sum()} of {n} descriptors differ"
        # every shard filled exactly its own block of output rows
        for g, o in enumerate(outs):
            row0, c = tiling.band_desc_rows(nfeat, g, nshards)
            lo, hi = min(row0, n), min(row0 + c, n)
            assert np.array_equal(o[1][f, lo:hi].cpu().numpy(), b[lo:hi]), f"shard {g} did not fill its rows [{lo}, {hi})"
    return int(cnt0.sum())


@pytest.mark.parametrize("nshards", [1, 2, 3, 8])
@pytest.mark.parametrize("dtype_name", ["BAD_512", "HASH_SIFT_256"])
def test_band_sharded_equals_single_gpu(oracle, dtype_name, nshards):
    import torch
    _run(torch, oracle, 1280, 720, 3000, dtype_name, nshards)


def test_band_sharded_more_shards_than_tile_rows(oracle):
    import torch
    # level 7 of a 300x200 frame has 2 tile rows: most of the 8 bands own nothing there
    _run(torch, oracle, 300, 200, 500, "BAD_256", 8)


@pytest.mark.parametrize("kw", [dict(nonmaxRadius=31), dict(nonmaxRadius=3), dict(nonmaxRadius=64, nlevels=4), dict(firstLevel=2)])
def test_band_sharded_other_parameters(oracle, kw):
    import torch
    _run(torch, oracle, 900, 700, 2000, "BAD_256", 3, **kw)


def test_band_sharded_batch_and_8k(oracle):
    import torch
    _run(torch, oracle, 800, 608, 1500, "HASH_SIFT_512", 2, frames=3)
    n = _run(torch, oracle, 7680, 4320, 40000, "HASH_SIFT_512", 4)
    assert n == 40000   # every per-level quota binds at 8K


@pytest.mark.parametrize("nfeat,nlevels", [(7, 8), (37, 16), (1, 8)])
def test_band_sharded_tiny_nfeatures(oracle, nfeat, nlevels):
    """per-level quotas whose sum exceeds nfeatures (7 -> 8, 37 with 16 levels -> 39): the candidate slots are laid out by the
    prefix of the quotas and must hold their sum (the buffer is sized from it, not from nfeatures)"""
    import torch
    _run(torch, oracle, 1280, 720, nfeat, "BAD_256", 3, nlevels=nlevels)


import os

N_BAND_FUZZ = int(os.environ.get("EF_FUZZ_BAND_CASES", "12"))


@pytest.mark.parametrize("case", range(N_BAND_FUZZ))
def test_band_sharded_random_configuration(oracle, case):
    """seeded random frame size, pyramid shape, FAST threshold, NMS radius, keypoint budget, descriptor and band count"""
    import torch
    rng = np.random.default_rng(0xEFB2B000 + case)
    w, h = int(rng.integers(120, 1100)), int(rng.integers(120, 800))
    sf = float(rng.choice([1.1, 1.2, 1.41, 2.0]))
    nlevels = int(rng.integers(1, 10))
    while nlevels > 1 and min(w, h) / sf ** (nlevels - 1) < 40.0:
        nlevels -= 1
    _run(torch, oracle, w, h, int(rng.choice([3, 50, 700, 2500])), str(rng.choice(["BAD_256", "BAD_512", "HASH_SIFT_256", "HASH_SIFT_512"])),
         int(rng.integers(2, 9)), frames=int(rng.choice([1, 1, 2])), allow_empty=True, scaleFactor=sf, nlevels=nlevels,
         fastThreshold=int(rng.choice([5, 20, 50])), nonmaxRadius=int(rng.choice([0, 2, 7, 15, 16, 33, 64])))
