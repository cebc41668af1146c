#!/usr/bin/env python3
"""Shared-memory wavefronts, instructions and stall samples per CUDA source line of one kernel in an ncu report (-lineinfo + --import-source on):
   python tools/ncu_smem_lines.py gpurun_out/prof.ncu-rep ef_hashsift_pipe [min_pct]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", f"regex:{pat}"],
                         capture_output=True, text=True).stdout
    hdr, acc, src_of, kernels = None, defaultdict(lambda: [0.0, 0.0, 0.0, 0.0]), {}, 0
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "Kernel Name":
            kernels += 1
            if kernels > 1:
                break
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 2 or r[2] != "-":
            continue
        d = dict(zip(hdr[4:], r[4:]))

        def f(k):
            try:
                return float(d.get(k, "0") or 0)
            except ValueError:
                return 0.0
        a = acc[r[0]]
        a[0] += f("L1 Wavefronts Shared"); a[1] += f("L1 Wavefronts Shared Ideal"); a[2] += f("Instructions Executed"); a[3] += f("# Samples")
        src_of[r[0]] = r[1].strip()
    tw = sum(a[0] for a in acc.values()) or 1.0
    ti = sum(a[2] for a in acc.values()) or 1.0
    ts = sum(a[3] for a in acc.values()) or 1.0
    print(f"kernel {pat}: {tw:.4g} shared wavefronts ({sum(a[1] for a in acc.values()):.4g} ideal), {ti:.4g} warp instructions, {ts:.0f} samples")
    for ln in sorted(acc, key=lambda k: int(k) if k.isdigit() else 0):
        w, wi, n, s = acc[ln]
        if 100 * w / tw >= min_pct or 100 * n / ti >= min_pct or 100 * s / ts >= min_pct:
            print(f"{ln:>5} wave {100 * w / tw:5.1f}% (x{w / wi if wi else 0:4.2f} ideal) inst {100 * n / ti:5.1f}% smp {100 * s / ts:5.1f}% | {src_of[ln][:110]}")


if __name__ == "__main__":
    main()
