set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2q_bench_2gpu.json 2> gpurun_out/r2q_bench_2gpu.err
echo "rc=$?" >> gpurun_out/r2q_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 scripts/run_aggregation_dist.py > gpurun_out/r2q_agg_2gpu.log 2>&1
