#!/usr/bin/env python
"""Text summary of one launch of an `ncu -i x.ncu-rep --page raw --csv` export: every metric of the families that
matter for the roofline argument, as `name [unit] = value` lines (the .ncu-rep itself stays in gpurun_out/, scratch).

    python tools/ncu_raw_summary.py raw.csv [launch_index] > profiles/<round>_ncu_<kernel>.txt
"""
import csv
import sys

KEEP = ("gpu__time", "dram__bytes", "dram__throughput", "dram__cycles_elapsed", "lts__t_bytes", "lts__t_sectors.sum",
        "lts__t_sectors_srcunit_tex", "lts__throughput", "l1tex__m_", "l1tex__data_bank", "l1tex__data_pipe_lsu_wavefronts",
        "l1tex__throughput", "sm__cycles", "sm__pipe_tensor", "sm__inst_executed.sum", "sm__inst_executed_pipe_tensor",
        "sm__throughput", "smsp__inst_executed.sum", "smsp__issue_active", "smsp__warp_issue_stalled", "launch__",
        "smsp__pipe_tensor", "sm__warps_active", "smsp__cycles_active")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hdr, units, row = rows[0], rows[1], rows[2 + idx]
    name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"# {name[:120]}")
    print(f"# launch {idx} of {len(rows) - 2} in {sys.argv[1]}; ncu --set full --clock-control none --import-source on")
    for h, u, v in sorted(zip(hdr, units, row)):
        if h.startswith(KEEP) and v not in ("", "n/a"):
            print(f"{h} [{u}] = {v}")


if __name__ == "__main__":
    main()
