set -x
mkdir -p gpurun_out
DRS_ROW=force timeout 900 python -m pytest tests/test_gpu_conv_layers.py tests/test_gpu_unet.py -x -q > gpurun_out/r2h_force.log 2>&1
echo "rc=$?" >> gpurun_out/r2h_force.log
DRS_ROW=force DRS_ROW_PIPES=1 timeout 900 python -m pytest tests/test_gpu_conv_layers.py tests/test_gpu_unet.py -x -q > gpurun_out/r2h_force1.log 2>&1
echo "rc=$?" >> gpurun_out/r2h_force1.log
for l in up_convs.2 conv_blocks.0.conv1; do
  DRS_V2_TIMELINE=1 DRS_V2_TIMELINE_LAYER=$l DRS_TL_PAIRS=24 timeout 300 python scripts/diag_layer_timeline.py > gpurun_out/r2h_tl_$l.log 2>&1
done
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu --no-aggregation --layers gpurun_out/r2h_layers.json > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
DRS_ROW_PIPES=1 timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu --no-aggregation --layers gpurun_out/r2h_layers_p1.json > gpurun_out/r2h_bench_p1.json 2> gpurun_out/r2h_bench_p1.err
DRS_ROW=force DRS_ROW_PIPES=2 timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu --no-aggregation --layers gpurun_out/r2h_layers_p2.json > gpurun_out/r2h_bench_p2.json 2> gpurun_out/r2h_bench_p2.err
