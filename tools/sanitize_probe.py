#!/usr/bin/env python3
"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck): detectAndCompute (BAD + HashSIFT incl. the tcgen05
projection), compute-only API, band-sharded path, matcher, colour conversion."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "cuda-efficient-features_b200"))
import torch
import efb200
from efb200 import tiling

g = torch.Generator(device="cpu").manual_seed(3)
h, nf = 301, 700
descs = {}
# 403: caller image not 16-byte aligned (level 0 through the non-TMA loaders, internal levels through TMA); 448: TMA everywhere;
# HASH_SIFT_256 = the seven-digit tcgen05 projection, HASH_SIFT_512 the six-digit one
for w, dt in ((403, efb200.BAD_256), (448, efb200.HASH_SIFT_512), (448, efb200.HASH_SIFT_256), (448, efb200.BAD_512)):
    img = efb200.synth_frames(2, h, w, 0xEFB20004, first_frame=w)
    ef = efb200.EfficientFeatures.create(nf, dtype=dt, max_width=w, max_height=h, max_batch=2, max_keypoints=2000)
    kp, desc, cnt = ef.detectAndComputeBatchRaw(img)
    n = int(cnt[0])
    descs[dt] = desc[0, :n].clone()
    k1, d1 = ef.detectAndComputeAsync(img[0])
    ef.computeAsync(img[0], k1)
    efs = [efb200.EfficientFeatures.create(nf, dtype=dt, max_width=w, max_height=h, max_batch=2) for _ in range(3)]
    kp2, desc2, cnt2, _ = tiling.detect_and_compute_tiled_emulated(efs, img)
    assert torch.equal(cnt, cnt2) and torch.equal(desc[0, :n], desc2[0, :n])
    if dt in (efb200.HASH_SIFT_512, efb200.HASH_SIFT_256):
        x = torch.randint(0, 256, (300, 128), dtype=torch.uint8, generator=g).cuda()
        assert torch.equal(ef.debugProject(x, 1), ef.debugProject(x, 3))
bf = efb200.BFMatcher.create()
for d in descs.values():
    i, dd = bf.knnMatchAsync(d, d.flip(0).contiguous(), 2)
    efb200.BFMatcher.create(efb200.NORM_HAMMING, True).matchAsync(d, d)
efb200.cvtColorToGray(torch.randint(0, 256, (37, 53, 3), dtype=torch.uint8, generator=g).cuda())
torch.cuda.synchronize()
print("sanitize probe OK")
