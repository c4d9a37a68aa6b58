set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2t_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
timeout 1200 ncu --set full --clock-control none -k regex:"conv_row_kernel|conv_gemm2" -c 28 -o gpurun_out/r2t_chain python bench.py --steps 1 --warmup 3 --no-cpu --no-aggregation > gpurun_out/r2t_ncu.log 2>&1
ncu -i gpurun_out/r2t_chain.ncu-rep --page raw --csv > gpurun_out/r2t_chain_raw.csv 2>/dev/null
rm -f gpurun_out/r2t_chain.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-aggregation > gpurun_out/r2t_ncu2.log 2>&1
