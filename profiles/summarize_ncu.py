#!/usr/bin/env python
"""Turns `ncu -i <rep> --page raw --csv` into the per-launch summary kept under profiles/.
usage: ncu -i gpurun_out/prof.ncu-rep --page raw --csv | python profiles/summarize_ncu.py > profiles/<name>.csv"""
import csv
import sys

WANT = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic"]
rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
out = csv.writer(sys.stdout)
out.writerow([f"{w} [{units[i]}]" if units[i] else w for w, i in cols])
for d in data:
    out.writerow([d[i][:60] for _, i in cols])
