#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/y1_tests.log 2>&1; tail -3 gpurun_out/y1_tests.log
python bench.py --steps 30 --warmup 5 --no-cpu --no-aggregation --layers gpurun_out/y1_layers.json > gpurun_out/y1_bench.json 2> gpurun_out/y1_bench.err; cat gpurun_out/y1_bench.json | head -c 600
