#!/usr/bin/env python
"""Per-kernel table from an `ncu --metrics ... --csv` log of whole steps (tools/gpu_r2n.sh): for every kernel the
launches captured, mean duration, tensor-pipe utilisation (convolutions) and DRAM bytes / achieved DRAM GB/s
(bandwidth kernels) against the measured HBM peak of MEASURED_PEAKS.json.

    python tools/ncu_step_table.py gpurun_out/r2n_step_metrics_yolo_voc.csv > profiles/r2n_step_metrics_yolo_voc.txt
"""
import csv
import json
import re
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def num(v):
    return float(v.replace(",", ""))


def main():
    path = sys.argv[1]
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm = peaks.get("hbm_gbs", 6539.9)
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    launches = OrderedDict()   # id -> {name, grid, metrics}
    for r in csv.DictReader(lines):
        i = r["ID"]
        d = launches.setdefault(i, {"name": re.sub(r"\(.*$", "", r["Kernel Name"]).replace("void ", "").replace("y2::", ""),
                                    "grid": r["Grid Size"], "m": {}})
        v = num(r["Metric Value"])
        u = r.get("Metric Unit", "")
        n = r["Metric Name"]
        if n == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1e-3)
        if n.startswith("dram__bytes") or n.startswith("lts__t_bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        d["m"][n] = v
    agg = OrderedDict()
    for d in launches.values():
        a = agg.setdefault(d["name"], [])
        a.append(d["m"])
    print(f"# {path}: {len(launches)} launches; duration = gpu__time_duration (ncu: cold caches, serialised, "
          f"--clock-control none); HBM peak {hbm:.0f} GB/s (MEASURED_PEAKS.json)")
    print(f"{'kernel':52s} {'n':>3s} {'us':>8s} {'tensor%':>8s} {'dram MB':>9s} {'GB/s':>8s} {'of HBM':>7s} {'L2 MB':>8s} {'GHz':>5s}")
    tot = 0.0
    rows = []
    for name, ms in agg.items():
        n = len(ms)
        us = sum(m.get("gpu__time_duration.sum", 0) for m in ms) / n
        tp = sum(m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0) for m in ms) / n
        dr = sum(m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0) for m in ms) / n
        l2 = sum(m.get("lts__t_bytes.sum", 0) for m in ms) / n
        ghz = sum(m.get("sm__cycles_elapsed.avg.per_second", 0) for m in ms) / n
        if ghz > 1e6:
            ghz /= 1e9
        gbs = dr / (us * 1e-6) / 1e9 if us else 0
        rows.append((us * n, name, n, us, tp, dr, gbs, l2, ghz))
        tot += us * n
    for _, name, n, us, tp, dr, gbs, l2, ghz in sorted(rows, reverse=True):
        print(f"{name[:52]:52s} {n:3d} {us:8.1f} {tp:8.1f} {dr / 1e6:9.1f} {gbs:8.0f} {gbs / hbm:7.2f} {l2 / 1e6:8.1f} {ghz:5.2f}")
    print(f"# total {tot:.0f} us over the captured launches")


if __name__ == "__main__":
    main()
