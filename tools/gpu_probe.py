#!/usr/bin/env python3
"""Verbose stage-by-stage comparison GPU vs oracle (prints statistics instead of asserting); used during
bring-up on the GPU box:  python tools/gpu_probe.py [--w 640 --h 480]"""
import argparse, sys, time, traceback
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "oracle", ROOT / "cuda-efficient-features_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import numpy as np
import torch
import efb200, efo, util

ap = argparse.ArgumentParser()
ap.add_argument("--w", type=int, default=1000); ap.add_argument("--h", type=int, default=700)
ap.add_argument("--nfeat", type=int, default=3000); ap.add_argument("--quick", action="store_true")
a = ap.parse_args()
o = efo.Oracle(); o.set_threads(min(16, o.max_threads()))
img = o.synth_frame(util.SEED + 5, 0, a.w, a.h)
d_img = torch.from_numpy(img).cuda()
print("device", torch.cuda.get_device_name(0), "lib", efb200.load_library().ef_version())

def section(name):
    print(f"\n=== {name} ===", flush=True)

try:
    section("stages (BAD_256 pipeline)")
    ef = efb200.EfficientFeatures.create(nfeatures=a.nfeat, dtype=efb200.BAD_256, max_width=a.w, max_height=a.h)
    kp, desc, cnt = ef.detectAndComputeRaw(d_img); torch.cuda.synchronize()
    print("count", int(cnt.item()), "workspace MB", ef.workspaceBytes() / 1e6)
    L = ef.getNLevels()
    pyr = o.pyramid(img, 1.2, L, False); bpyr = o.pyramid(img, 1.2, L, True)
    gc = ef.debugLevelCounts(); _, oc = o.detect(img, o.make_params(nfeatures=a.nfeat, desc_type=efo.BAD_256))
    print("gpu counts\n", gc.T, "\noracle counts\n", oc.T)
    for l in range(L):
        A = ef.debugLevelArrays(l)
        resp, nc = o.score_map(pyr[l], 20)
        fin_g, fin_o = np.isfinite(A["response"]), np.isfinite(resp)
        both = fin_g & fin_o
        print(f"L{l} {A['width']}x{A['height']} pyr_mism {(A['image'] != pyr[l]).sum()} blur_mism {(A['blurred'] != bpyr[l]).sum()} "
              f"corners gpu {fin_g.sum()} ora {fin_o.sum()} xor {(fin_g ^ fin_o).sum()} resp_bit_mism {(A['response'][both].view(np.uint32) != resp[both].view(np.uint32)).sum()}")
except Exception:
    traceback.print_exc()

for name in ("BAD_256", "BAD_512", "HASH_SIFT_256", "HASH_SIFT_512"):
    try:
        section(f"detectAndCompute {name}")
        ef = efb200.EfficientFeatures.create(nfeatures=a.nfeat, dtype=getattr(efb200, name), max_width=a.w, max_height=a.h)
        kp, desc = ef.detectAndComputeAsync(d_img)
        g = ef.convert(kp); gd = desc.cpu().numpy()
        ok, od, _ = o.detect_and_compute(img, o.make_params(nfeatures=a.nfeat, desc_type=getattr(efo, name)))
        os_ = util.oracle_to_struct(ok)
        print("n gpu", len(g), "oracle", len(os_))
        if len(g) == len(os_):
            gs, go = util.canon_keypoints(g); oss, oo = util.canon_keypoints(os_)
            for f in ("x", "y", "octave"):
                print(" ", f, "mism", (gs[f] != oss[f]).sum())
            for f in ("response", "angle", "size"):
                print(" ", f, "bit mism", (gs[f].view(np.uint32) != oss[f].view(np.uint32)).sum())
            d = gd[go] != od[oo]
            print("  descriptor rows differing", d.any(axis=1).sum(), "bytes", d.sum(), "of", d.size)
        else:
            sg = set(zip(g["octave"].tolist(), g["y"].tolist(), g["x"].tolist())); so = set(zip(os_["octave"].tolist(), os_["y"].tolist(), os_["x"].tolist()))
            print("  only gpu", len(sg - so), "only oracle", len(so - sg), list(sg - so)[:5], list(so - sg)[:5])
    except Exception:
        traceback.print_exc()

if not a.quick:
    k = efo.stress_keypoints(a.w, a.h, 5000, seed=9)
    for nbits in (256, 512):
        for scale in (1.0, 5.0):
            try:
                section(f"compute-only BAD{nbits} scale {scale}")
                bad = efb200.BAD.create(scale, 100 if nbits == 512 else 101, max_width=a.w, max_height=a.h)
                g = bad.compute(img, k); ob = o.bad(img, k, scale, nbits)
                d = g != ob
                print("rows differing", d.any(axis=1).sum(), "of", len(k), "first rows", np.nonzero(d.any(axis=1))[0][:8], k[np.nonzero(d.any(axis=1))[0][:4]])
            except Exception:
                traceback.print_exc()
        try:
            section(f"compute-only HashSIFT{nbits}")
            hs = efb200.HashSIFT.create(1.0, 100 if nbits == 512 else 101, max_width=a.w, max_height=a.h)
            hs._ef.debugKeepProjection(True)
            g = hs.compute(img, k)
            sift, proj = hs._ef.debugHashSift(len(k))
            feat = o.hashsift_features(img, k, 1.0); od, oproj = o.hashsift(img, k, 1.0, nbits, want_proj=True)
            fd = sift != feat[:, 1:].astype(np.uint8)
            rows = fd.any(axis=1)
            print("sift rows differing", rows.sum(), "of", len(k), "max abs diff", np.abs(sift.astype(int) - feat[:, 1:].astype(int)).max())
            print("first differing keypoints", k[np.nonzero(rows)[0][:5]])
            same = ~rows
            print("proj bit mism on identical rows", (proj[same].view(np.uint32) != oproj[same].view(np.uint32)).sum(), "desc rows differing overall", (g != od).any(axis=1).sum())
            for i in np.nonzero(rows)[0][:3]:
                pg = None
                po = o.hashsift_patch(img, k[i], 1.0)
                print("  kp", k[i], "nz diffs", np.nonzero(fd[i])[0][:10], sift[i][fd[i]][:10], feat[i, 1:][fd[i]][:10])
        except Exception:
            traceback.print_exc()
print("\nprobe done")
