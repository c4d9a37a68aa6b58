#!/bin/bash
# The ncu passes behind profiles/ncu_r2_*.csv and profiles/ncu_traffic.json (run on the GPU box, one GPU, after the
# same commands have exited 0 without ncu). Outputs go to gpurun_out/; profiles/summarize_ncu.py and
# profiles/make_ncu_traffic.py turn them into the committed summaries.
set -u
OUT=${1:-gpurun_out}
# 1. launch list of the bench command (cold-cache, serialised per-launch times: shares of the step, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-aggregation > $OUT/ncu_launches.log 2>&1
# 2. full capture of the 28 tensor-core launches of the third UNet evaluation at cfg 2
ncu --set full --clock-control none --import-source on -k regex:'conv_(row|gemm2|gemm2c)_kernel' --launch-skip 56 -c 28 \
    -f -o $OUT/ncu_conv_chain python scripts/diag_forward.py 16 256 > $OUT/ncu_conv_chain.log 2>&1
ncu -i $OUT/ncu_conv_chain.ncu-rep --page raw --csv 2>/dev/null | python profiles/summarize_ncu.py > $OUT/ncu_conv_chain_full.csv
