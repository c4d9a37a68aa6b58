#!/usr/bin/env python
"""profiles/ncu_traffic.json (read by bench.py for roofline.traffic) from a summarised conv-chain capture.
usage: python profiles/make_ncu_traffic.py profiles/ncu_r2_conv_chain_full.csv > profiles/ncu_traffic.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[1:]


def col(prefix):
    i = [k for k, h in enumerate(hdr) if h.startswith(prefix)][0]
    unit = hdr[i][hdr[i].index("[") + 1:hdr[i].index("]")] if "[" in hdr[i] else ""
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(unit, 1.0)
    return [float(d[i]) * scale for d in data]


rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
t = col("gpu__time_duration.sum")
tp = col("sm__pipe_tensor_cycles_active")
n = len(data)
print(json.dumps({
    "dram_bytes_per_launch": (sum(rd) + sum(wr)) / n,
    "launches": n,
    "dram_read_bytes_per_eval": sum(rd),
    "dram_write_bytes_per_eval": sum(wr),
    "time_weighted_tensor_pipe_active_pct": sum(a * b for a, b in zip(t, tp)) / sum(t),
    "serialised_us_per_eval": sum(t),
    "source": sys.argv[1] + ": ncu --set full --clock-control none over the %d tensor-core launches (conv_row_kernel / "
              "conv_gemm2_kernel / conv_gemm2c_kernel) of one UNet evaluation at cfg 2 (scripts/profile_ncu.sh; "
              "dram__bytes_read.sum + dram__bytes_write.sum; ncu flushes the caches before every launch, so reads are "
              "cold and write-backs that drain after a launch are not counted)" % n,
}, indent=1))
