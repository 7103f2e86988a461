"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`) of a
bench.py run: launches, total time and share per kernel of THIS library (everything else - torch's workload
generation, the cuBLASLt probe, NCCL - is listed as one line).

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launch_summary.txt
"""
from __future__ import annotations

import csv
import re
import sys
from collections import defaultdict

OURS = ("match_top2", "top2_", "row_sqnorm", "db_tile", "db_norm", "pack_u8", "hough_", "group_scatter",
        "exclusive_scan", "scan_tile_sums", "pose_bin_index", "compact_", "affine_", "valid_records",
        "xchg_", "adjacency_kernel", "label_")


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|sod::", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("void ", "").strip()


def main(path: str) -> None:
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for rec in csv.DictReader(lines):
        if rec.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(rec["Metric Value"].replace(",", ""))
        unit = rec.get("Metric Unit", "ns")
        ms = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit.replace("second", "s"), 1e-6)
        rows.append((short(rec["Kernel Name"]), ms))
    per = defaultdict(lambda: [0.0, 0])
    other = [0.0, 0]
    for name, ms in rows:
        tgt = per[name] if any(k in name for k in OURS) else other
        tgt[0] += ms
        tgt[1] += 1
    total = sum(v[0] for v in per.values())
    n = sum(v[1] for v in per.values())
    print(f"our kernels: {n} launches, {total:.3f} ms total")
    print(f"{'ms':>10} {'share':>6} {'launches':>8}  kernel")
    for name, (ms, cnt) in sorted(per.items(), key=lambda kv: -kv[1][0]):
        print(f"{ms:10.3f} {100 * ms / total:5.1f}% {cnt:8d}  {name}")
    print(f"{other[0]:10.3f} {'':>6} {other[1]:8d}  (other: torch / cuBLASLt / NCCL kernels of the harness)")


if __name__ == "__main__":
    main(sys.argv[1])
