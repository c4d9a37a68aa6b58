#!/bin/bash
python -m pytest tests/test_gpu_conv_layers.py tests/test_gpu_kernel_variants.py tests/test_gpu_full_size.py -x -q > gpurun_out/y4_tests.log 2>&1; tail -5 gpurun_out/y4_tests.log
python bench.py --steps 30 --warmup 5 --no-cpu --no-aggregation --layers gpurun_out/y4_layers.json > gpurun_out/y4_bench.json 2> gpurun_out/y4_bench.err; head -c 400 gpurun_out/y4_bench.json
