set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 4000 --warmup 5 --no-cpu --no-aggregation > gpurun_out/r2m_bench_long.json 2> gpurun_out/r2m_bench_long.err
