#!/usr/bin/env python3
"""Per-source-line summary of an ncu report (needs -lineinfo + --import-source on):
   python tools/ncu_lines.py gpurun_out/prof.ncu-rep ef_hashsift_pipe [min_pct]
Prints, for the first launch of the kernel matching the regex, every CUDA source line that holds at least
min_pct % of the stall samples or of the executed warp instructions."""
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", f"regex:{pat}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, lines, seen_kernel = None, None, [], 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Kernel Name":
            seen_kernel += 1
            if seen_kernel > 1:
                break
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 2:
            continue
        if r[2] != "-":  # SASS row
            continue
        d = dict(zip(hdr[4:], r[4:]))
        def f(k):
            try:
                return float(d.get(k, "0") or 0)
            except ValueError:
                return 0.0
        stalls = {k[6:]: f(k) for k in d if k.startswith("stall_") and "Not Issued" not in k}
        lines.append((cur_file, r[0], r[1], f("# Samples"), f("Instructions Executed"), stalls,
                      f("L1 Wavefronts Shared Excessive"), f("L2 Theoretical Sectors Global Excessive")))
    ts = sum(l[3] for l in lines) or 1.0
    ti = sum(l[4] for l in lines) or 1.0
    print(f"kernel {pat}: {ts:.0f} samples, {ti:.3e} warp instructions")
    for fl, ln, src, s, n, st, smx, l2x in lines:
        if 100 * s / ts >= min_pct or 100 * n / ti >= min_pct:
            top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            tops = " ".join(f"{k}:{v / max(s, 1):.0%}" for k, v in top if v > 0)
            print(f"{fl}:{ln:>4} smp {100 * s / ts:5.1f}% inst {100 * n / ti:5.1f}%  {tops:40s} | {src.strip()[:100]}")


if __name__ == "__main__":
    main()
