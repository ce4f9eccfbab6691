#!/usr/bin/env python3
"""Summarise an `ncu --set full` report: `ncu -i X.ncu-rep --page raw --csv | python tools/ncu_full_summary.py > out.json`.
One record per captured launch: duration, DRAM bytes, tensor-pipe activity, occupancy, registers."""
import csv
import json
import re
import sys

WANT = {
    "gpu__time_duration.sum": "duration_ns",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__t_bytes.sum": "l2_bytes",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct_of_peak",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_instructions",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return x


def main():
    rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
    header, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(header, r))
        u = dict(zip(header, units))
        name = re.sub(r"\(.*", "", d.get("Kernel Name", ""))
        name = re.sub(r"mmfn::|tc::|<unnamed>::|\(anonymous namespace\)::|void ", "", name)
        rec = {"kernel": name, "grid": d.get("Grid Size"), "block": d.get("Block Size")}
        for k, short in WANT.items():
            if k in d and d[k] != "":
                v = num(d[k])
                unit = u.get(k, "")
                if isinstance(v, float):
                    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}.get(unit, 1.0)
                    v *= scale
                rec[short] = v
        if "dram_read_bytes" in rec and "dram_write_bytes" in rec:
            rec["traffic_bytes"] = rec["dram_read_bytes"] + rec["dram_write_bytes"]
        out.append(rec)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
