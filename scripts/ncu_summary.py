#!/usr/bin/env python
"""Condenses `ncu -i X.ncu-rep --page raw --csv` output into the few numbers the roofline discussion needs.

    python scripts/ncu_summary.py gpurun_out/prof_gemm.raw.csv [more.csv ...]  > profiles/ncu/SUMMARY.md
"""
import csv
import json
import os
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active, % of elapsed"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active, % of SM-active cycles"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__cycles_active.avg", "SM active cycles (avg)"),
    ("gpc__cycles_elapsed.max", "elapsed cycles"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "instructions"),
]


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        out.append((d, u))
    return out


def main():
    peaks = {}
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    print("# ncu captures (`ncu --set full --clock-control none --import-source on`, one launch each, 1 GPU)\n")
    print("Raw reports: the `.ncu-rep` files next to this summary (open with `ncu -i <file> --page raw|source`).\n"
          "Durations here are under the profiler (serialised, cold caches) and are NOT benchmark numbers; the\n"
          "benchmark numbers are CUDA-event timings in `profiles/README.md`.\n")
    for path in sys.argv[1:]:
        for d, u in load(path):
            name = d.get("Kernel Name", "?")
            print(f"## `{os.path.basename(path).replace('.raw.csv', '.ncu-rep')}` — `{name[:110]}`\n")
            print("| metric | value |")
            print("|---|---|")
            for key, label in KEYS:
                if key in d and d[key] != "":
                    print(f"| {label} (`{key}`) | {d[key]} {u.get(key, '')} |")
            try:
                rd = float(d["dram__bytes_read.sum"].replace(",", ""))
                wr = float(d["dram__bytes_write.sum"].replace(",", ""))
                scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
                tot = rd * scale.get(u["dram__bytes_read.sum"], 1.0) + wr * scale.get(u["dram__bytes_write.sum"], 1.0)
                print(f"| DRAM traffic (read + write) | {tot / 1e6:.1f} MB |")
            except Exception:
                pass
            print()


if __name__ == "__main__":
    main()
