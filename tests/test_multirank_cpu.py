"""world_size-2 gloo test of the multi-GPU host logic (frame sharding, max-over-ranks timing, count gather).
The data path itself has no collective (frames are independent)."""
import os
import sys
from pathlib import Path

import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT / "cuda-efficient-features_b200"))
    sys.path.insert(0, str(ROOT / "oracle"))
    import numpy as np
    import torch.distributed as dist
    from efb200.sharding import gather_counts, reduce_max_time, shard_range
    import efo
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nframes = 5
    b, e = shard_range(nframes, rank, world)
    o = efo.Oracle()
    # each rank "processes" its own frames (CPU oracle stands in for the device path in this host-logic test)
    counts = []
    for f in range(b, e):
        img = o.synth_frame(0xEFB20000 + 3, f, 320, 240)
        kp, _ = o.detect(img, o.make_params(nfeatures=300, desc_type=efo.BAD_256))
        counts.append(len(kp))
    allc = gather_counts(counts)
    tmax = reduce_max_time(10.0 + rank)
    q.put((rank, (b, e), allc, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    world, port = 2, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == (0, 3) and res[1][1] == (3, 5)
    assert res[0][2] == res[1][2] and len(res[0][2]) == 5      # every rank sees all per-frame counts in frame order
    assert res[0][3] == res[1][3] == 11.0                       # max over ranks
    # sharded result == single-process result
    sys.path.insert(0, str(ROOT / "oracle"))
    import efo
    o = efo.Oracle()
    single = [len(o.detect(o.synth_frame(0xEFB20000 + 3, f, 320, 240), o.make_params(nfeatures=300, desc_type=efo.BAD_256))[0]) for f in range(5)]
    assert single == res[0][2]
