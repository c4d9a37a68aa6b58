set -x
mkdir -p gpurun_out
timeout 600 python scripts/diag_batch_sweep.py 31 > gpurun_out/r2c_sweep.log 2>&1
timeout 600 python -m pytest tests/test_gpu_aggregation.py -x -q > gpurun_out/r2c_pytest.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2c_ncu.log 2>&1
