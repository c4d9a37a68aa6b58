#!/bin/bash
DRS_V2_VERBOSE=1 python scripts/diag_layer_timeline.py 2>&1 | grep "drs\]" > gpurun_out/y6_verbose.log
python bench.py --steps 30 --warmup 5 --no-cpu --no-aggregation --layers gpurun_out/y6_layers.json > gpurun_out/y6_bench.json 2> gpurun_out/y6_bench.err; head -c 300 gpurun_out/y6_bench.json
