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


# ---------------------------------------------------------------------------------------------------
# one oversized frame over several ranks (efb200/tiling.py): band partition + the two collectives
# ---------------------------------------------------------------------------------------------------
def test_band_partition_covers_every_tile_row_once():
    sys.path.insert(0, str(ROOT / "cuda-efficient-features_b200"))
    from efb200.tiling import band_tile_rows
    for tiles_y in (1, 2, 7, 19, 68, 135):
        for n in (1, 2, 3, 4, 8):
            for halo in (0, 1, 3):
                owned = []
                for g in range(n):
                    o0, on, s0, sn = band_tile_rows(tiles_y, g, n, halo)
                    owned += list(range(o0, o0 + on))
                    if on == 0:
                        assert sn == 0
                    else:   # score rows = owned rows + halo, clipped to the level
                        assert s0 == max(0, o0 - halo) and s0 + sn == min(tiles_y, o0 + on + halo)
                assert owned == list(range(tiles_y)), (tiles_y, n)


class _StubFeatures:
    """Stands in for EfficientFeatures on CPU tensors: checks what the collectives deliver."""

    def __init__(self, rank, world, nf=10):
        self.rank, self.world, self.nf = rank, world, nf

    def bandDetect(self, images, shard, nshards, stream=None):
        import torch
        assert (shard, nshards) == (self.rank, self.world)
        return torch.full((images.shape[0], 48), shard + 1, dtype=torch.uint8) + images[:, 0, :48]

    def bandFinish(self, all_cand, shard, nshards, want_descriptors=True, out=None):
        import torch
        assert tuple(all_cand.shape) == (nshards, 2, 48)
        for g in range(nshards):   # shard order, every rank's candidates
            assert int(all_cand[g, 0, 0]) == g + 1 + 7
        from efb200.tiling import band_desc_rows
        desc = torch.full((2, self.nf, 4), 255, dtype=torch.uint8)        # rows of other ranks: undefined (here 255)
        row0, c = band_desc_rows(self.nf, shard, nshards)
        desc[:, row0:row0 + c] = 100 + shard
        return torch.zeros((2, 5, self.nf)), desc, torch.full((2,), self.nf, dtype=torch.int32)


def _band_worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT / "cuda-efficient-features_b200"))
    import torch
    import torch.distributed as dist
    from efb200.tiling import detect_and_compute_tiled
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    images = torch.full((2, 64, 64), 7 if rank == 0 else 0, dtype=torch.uint8)   # only the source rank has the frame
    kp, desc, counts = detect_and_compute_tiled(_StubFeatures(rank, world, 10), images, src=0)
    kp, desc_odd, counts = detect_and_compute_tiled(_StubFeatures(rank, world, 7), images, src=0)   # 7 rows: not divisible, padded path
    q.put((rank, desc[0, :, 0].tolist() + desc_odd[1, :, 3].tolist(), int(images[1, 5, 5])))
    dist.barrier()
    dist.destroy_process_group()


def test_band_collectives_gloo():
    world, port = 2, 29741
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_band_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rows, px in res:
        assert px == 7                                      # image broadcast from rank 0
        assert rows == [100] * 5 + [101] * 5 + [100] * 4 + [101] * 3   # all-gather of the equal row blocks (in place / padded)
