#!/usr/bin/env python3
"""torchrun script: band-sharded detectAndCompute over real NCCL (broadcast + all-gather of candidates + all-gather of descriptor row blocks) must equal the
single-GPU result on every rank, bit for bit.  Prints one line per rank; exit code != 0 on mismatch."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "cuda-efficient-features_b200"))

import torch
import torch.distributed as dist

import efb200
from efb200 import tiling


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for (w, h, nf, dt) in [(3840, 2160, 40000, efb200.HASH_SIFT_512), (1920, 1080, 5000, efb200.BAD_256), (7680, 4320, 40000, efb200.BAD_512), (1280, 720, 3001, efb200.HASH_SIFT_256)]:
        g = torch.Generator(device="cpu").manual_seed(1234 + w)
        img = torch.randint(0, 256, (1, h, w), dtype=torch.uint8, generator=g)
        d = img.cuda() if rank == 0 else torch.zeros_like(img).cuda()
        ef = efb200.EfficientFeatures.create(nfeatures=nf, dtype=dt, max_width=w, max_height=h, device=local)
        kp, desc, cnt = tiling.detect_and_compute_tiled(ef, d, src=0)
        kp0, desc0, cnt0 = ef.detectAndComputeBatchRaw(d)   # d now holds the broadcast frame on every rank
        torch.cuda.synchronize()
        n = int(cnt0[0])
        same = (int(cnt[0]) == n and torch.equal(kp[0, :, :n].view(torch.int32), kp0[0, :, :n].view(torch.int32))
                and torch.equal(desc[0, :n], desc0[0, :n]))
        print(f"rank {rank}/{world} {w}x{h} dtype {dt}: {n} keypoints, band-sharded == single GPU: {same}", flush=True)
        ok = ok and same
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
