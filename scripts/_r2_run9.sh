set -x
mkdir -p gpurun_out
DRS_ROW=force timeout 900 python -m pytest tests/test_gpu_conv_layers.py tests/test_gpu_unet.py -x -q > gpurun_out/r2i_force.log 2>&1
echo "rc=$?" >> gpurun_out/r2i_force.log
for l in up_convs.2 conv_blocks.0.conv1 conv_blocks.0.conv2; do
  DRS_V2_TIMELINE=1 DRS_V2_TIMELINE_LAYER=$l timeout 300 python scripts/diag_row_timeline.py > gpurun_out/r2i_tl_$l.log 2>&1
done
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu --no-aggregation --layers gpurun_out/r2i_layers.json > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
