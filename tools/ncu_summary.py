#!/usr/bin/env python3
"""Summarise an ncu report into a text table (for profiles/) and, optionally, the per-stage DRAM traffic JSON that
bench.py attaches to its roofline objects:

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--traffic profiles/traffic.json] > profiles/rNN_summary.txt

Per launch: duration, DRAM read/write bytes, DRAM / SM / issue / occupancy percentages, registers, warp instructions,
shared-memory wavefronts and bank conflicts, tensor-pipe activity.  Times under ncu are cold-cache and serialised: compare
shares, not absolutes."""
import csv
import json
import subprocess
import sys

STAGE_OF = [("ef_resize", "pyramid"), ("ef_score", "score"), ("ef_nms", "nms"), ("ef_compact", "compact"), ("ef_select", "select"),
            ("ef_angle_pack", "angle_pack"), ("ef_blur", "blur"), ("ef_hashsift_pipe", "describe"), ("ef_bad_pipe", "describe_bad"),
            ("ef_hashsift_project", "project")]
COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp_inst"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conflicts"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst")]


def to_bytes(value, unit):
    v = float(value)
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def main():
    rep = sys.argv[1]
    traffic_path = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# " + " | ".join(["kernel", "grid"] + [c[1] for c in COLS]))
    traffic, winst, wavef = {}, {}, {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        if name.startswith("void "):
            name = name[5:]
        cells = [name, r[idx["Grid Size"]]]
        for key, _ in COLS:
            if key in idx:
                cells.append(f"{r[idx[key]]} {units[idx[key]]}".strip())
            else:
                cells.append("n/a")
        print(" | ".join(cells))
        for pat, stage in STAGE_OF:
            if name.startswith(pat) and "dram__bytes_read.sum" in idx:
                b = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                    to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
                traffic.setdefault(stage, []).append(b)
                if "smsp__inst_executed.sum" in idx:
                    winst.setdefault(stage, []).append(float(r[idx["smsp__inst_executed.sum"]]))
                if "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum" in idx:
                    wavef.setdefault(stage, []).append(float(r[idx["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]]))
                break
    if traffic_path:
        # pyramid = 7 launches per step (sum); other stages = one launch per step (mean over the captured launches)
        res = {}
        for stage, v in traffic.items():
            res[stage] = sum(v) if stage == "pyramid" else sum(v) / len(v)
        # executed warp instructions and shared-memory wavefronts per launch group: bench.py turns them into issue-slot and
        # shared-memory-pipe fractions with its own live timing (the bounds that matter for these kernels)
        res["_warp_inst"] = {st: (sum(v) if st == "pyramid" else sum(v) / len(v)) for st, v in winst.items()}
        res["_smem_wavefronts"] = {st: (sum(v) if st == "pyramid" else sum(v) / len(v)) for st, v in wavef.items()}
        res["_frames"] = int(sys.argv[sys.argv.index("--frames") + 1]) if "--frames" in sys.argv else 8
        # hash of the kernel sources the capture was taken from: bench.py reports the capture only while it matches
        import hashlib
        from pathlib import Path
        hsh = hashlib.sha256()
        for f in sorted((Path(__file__).resolve().parent.parent / "cuda-efficient-features_b200" / "csrc").glob("*.cu*")):
            hsh.update(f.name.encode()); hsh.update(f.read_bytes())
        res["_src_sha"] = hsh.hexdigest()[:16]
        res["_note"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch group (one step of the captured bench command, _frames frames), from " + rep
        with open(traffic_path, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
