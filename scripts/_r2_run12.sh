set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_kernel_variants.py -x -q > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
timeout 300 python scripts/diag_sample_overhead.py 50 > gpurun_out/r2l_overhead.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-aggregation > gpurun_out/r2l_ncu.log 2>&1
