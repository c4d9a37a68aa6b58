set -x
mkdir -p gpurun_out
DRS_ROW=force timeout 900 python -m pytest tests/test_gpu_conv_layers.py tests/test_gpu_unet.py -x -q > gpurun_out/r2e_force.log 2>&1
echo "rc=$?" >> gpurun_out/r2e_force.log
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu --no-aggregation --layers gpurun_out/r2e_layers.json > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
timeout 300 python scripts/diag_blend.py 4 > gpurun_out/r2e_blend_serial.log 2>&1
DRS_BLEND_BATCH=1 timeout 300 python scripts/diag_blend.py 4 > gpurun_out/r2e_blend_batch.log 2>&1
for l in up_convs.2 conv_blocks.0.conv1; do
  DRS_V2_TIMELINE=4 timeout 300 python scripts/diag_graph_spans.py > gpurun_out/r2e_spans.log 2>&1
done
