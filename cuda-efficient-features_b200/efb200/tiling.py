"""One oversized frame over several GPUs: horizontal bands with halo (SURVEY 8e "one oversized frame").

Every rank holds the whole image (rank `src` broadcasts it over NCCL/NVLink -- the only image movement), builds the whole
pyramid, and runs FAST/Harris + radius NMS + compaction on its band of every level.  Two small collectives complete the
frame: an all-gather of the per-band top-quota candidates (<= 8 B x nfeatures per rank) before the global per-level
selection, and an all-gather of the descriptor matrix in equal blocks of output rows (rank g describes rows
[g C, (g + 1) C), C = ceil(nfeatures / world): 1/world of the bytes per rank, no reduction anywhere).  The result on every rank
is bit-identical to the single-GPU result.
"""
from __future__ import annotations

import ctypes as C

COLLECTIVES = "NCCL broadcast of the image, all-gather of band candidates, all-gather of equal descriptor row blocks (no reduction)"


def band_tile_rows(tiles_y: int, shard: int, nshards: int, halo_tiles: int = 1):
    """(own0, own_n, score0, score_n): tile rows owned by `shard` and the rows its score stage covers (host arithmetic of
    ef_band_tile_rows; callable without a GPU)."""
    from . import load_library
    L = load_library()
    v = [C.c_int() for _ in range(4)]
    L.ef_band_tile_rows(tiles_y, shard, nshards, halo_tiles, *[C.byref(x) for x in v])
    return tuple(x.value for x in v)


def band_desc_rows(nfeatures: int, shard: int, nshards: int):
    """(row0, nrows): the block of descriptor rows `shard` fills (host arithmetic of ef_band_desc_rows; callable without a GPU)."""
    from . import load_library
    L = load_library()
    a, b = C.c_int(), C.c_int()
    L.ef_band_desc_rows(nfeatures, shard, nshards, C.byref(a), C.byref(b))
    return a.value, b.value


def detect_and_compute_tiled(ef, images, group=None, src=None, want_descriptors=True, out=None):
    """detectAndCompute of F whole frames cut into world_size bands; call on every rank of `group` with an
    EfficientFeatures created on the rank's GPU.  `images`: F x H x W uint8 CUDA tensor; with `src` given it is
    broadcast from that rank first (the other ranks pass a buffer of the same shape).  Returns (keypoints F x 5 x nfeatures,
    descriptors F x nfeatures x B, counts F), complete on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world > 1 and src is not None:
        dist.broadcast(images, src=src, group=group)
    cand = ef.bandDetect(images, rank, world)
    if world > 1:
        all_cand = torch.empty((world,) + tuple(cand.shape), dtype=torch.uint8, device=cand.device)
        dist.all_gather_into_tensor(all_cand.view(world * cand.shape[0], cand.shape[1]), cand, group=group)  # rank-major concatenation
    else:
        all_cand = cand[None]
    kp, desc, counts = ef.bandFinish(all_cand, rank, world, want_descriptors=want_descriptors, out=out)
    if world > 1 and desc is not None:
        # every rank filled rows [rank C, (rank + 1) C) of every frame: all-gather of the equal blocks.  In place when the matrix
        # holds world * C rows (nfeatures divisible by world); otherwise through a padded buffer.
        F, nf, B = desc.shape
        row0, c = band_desc_rows(nf, rank, world)
        for f in range(F):
            if world * c == nf and desc[f].is_contiguous():
                dist.all_gather_into_tensor(desc[f].view(-1), desc[f, row0:row0 + c].reshape(-1), group=group)
            else:
                mine = torch.zeros((c, B), dtype=desc.dtype, device=desc.device)
                n_own = max(0, min(c, nf - row0))
                mine[:n_own] = desc[f, row0:row0 + n_own]
                full = torch.empty((world * c, B), dtype=desc.dtype, device=desc.device)
                dist.all_gather_into_tensor(full.view(-1), mine.view(-1), group=group)
                desc[f].copy_(full[:nf])
    return kp, desc, counts


def detect_and_compute_tiled_emulated(efs, images, want_descriptors=True):
    """The same data flow on ONE GPU with len(efs) handles standing in for the ranks (tests, and a reference for the
    collective plumbing): concatenation replaces both all-gathers."""
    import torch
    n = len(efs)
    cands = [ef.bandDetect(images, g, n) for g, ef in enumerate(efs)]
    all_cand = torch.stack(cands, 0).contiguous()
    outs = [ef.bandFinish(all_cand, g, n, want_descriptors=want_descriptors) for g, ef in enumerate(efs)]
    kp, desc, counts = outs[0]
    if desc is not None:
        nf = desc.shape[1]
        desc = desc.clone()
        for g, o in enumerate(outs):
            row0, c = band_desc_rows(nf, g, n)
            desc[:, row0:row0 + c] = o[1][:, row0:row0 + c]
    return kp, desc, counts, outs
