set -x
mkdir -p gpurun_out
DRS_ROW=force DRS_V2_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_conv_layers.py -x -q > gpurun_out/r2d_layers_force.log 2>&1
echo "rc=$?" >> gpurun_out/r2d_layers_force.log
DRS_ROW=force timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_sampler.py -x -q > gpurun_out/r2d_unet_force.log 2>&1
echo "rc=$?" >> gpurun_out/r2d_unet_force.log
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu --no-aggregation --layers gpurun_out/r2d_layers.json > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
DRS_ROW=0 timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu --no-aggregation --layers gpurun_out/r2d_layers_norow.json > gpurun_out/r2d_bench_norow.json 2> gpurun_out/r2d_bench_norow.err
timeout 300 python scripts/diag_blend.py 4 > gpurun_out/r2d_blend.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:blend_gather4 -c 1 -o gpurun_out/r2d_blend python scripts/diag_blend.py 2 > gpurun_out/r2d_blend_ncu.log 2>&1
ncu -i gpurun_out/r2d_blend.ncu-rep --page raw --csv > gpurun_out/r2d_blend_raw.csv 2>/dev/null
