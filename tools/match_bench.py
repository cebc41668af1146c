#!/usr/bin/env python3
"""Times the Hamming matcher (ef_match_knn_async, k = 2) on N x N descriptors with CUDA events; prints one JSON line."""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "cuda-efficient-features_b200"))
import torch, efb200

def main():
    out = []
    for n, nbytes in [(40000, 64), (40000, 32), (5000, 64)]:
        g = torch.Generator(device="cpu").manual_seed(n)
        q = torch.randint(0, 256, (n, nbytes), dtype=torch.uint8, generator=g).cuda()
        t = torch.randint(0, 256, (n, nbytes), dtype=torch.uint8, generator=g).cuda()
        bf = efb200.BFMatcher.create()
        for _ in range(3): bf.knnMatchAsync(q, t, 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): bf.knnMatchAsync(q, t, 2)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out.append({"n": n, "bits": nbytes * 8, "ms": ms, "pairs_per_s": n * n / ms * 1e3, "bit_ops_per_s": n * n * nbytes * 8 / ms * 1e3})
    print(json.dumps({"matcher_knn2": out}))

if __name__ == "__main__":
    main()
