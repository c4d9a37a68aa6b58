set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_sampler.py -x -q > gpurun_out/r2s_unet.log 2>&1
echo "rc=$?" >> gpurun_out/r2s_unet.log
timeout 900 python -m pytest tests/test_gpu_full_size.py -x -q > gpurun_out/r2s_full.log 2>&1
echo "rc=$?" >> gpurun_out/r2s_full.log
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu --no-aggregation --layers gpurun_out/r2s_layers.json > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
DRS_V2_TIMELINE=4 timeout 300 python scripts/diag_graph_spans.py > gpurun_out/r2s_spans.log 2>&1
