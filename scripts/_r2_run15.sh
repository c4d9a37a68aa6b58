set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2o_bench_8gpu.json 2> gpurun_out/r2o_bench_8gpu.err
echo "rc=$?" >> gpurun_out/r2o_bench_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2o_bench_2gpu.json 2> gpurun_out/r2o_bench_2gpu.err
echo "rc=$?" >> gpurun_out/r2o_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/r2o_bench_4gpu.json 2> gpurun_out/r2o_bench_4gpu.err
echo "rc=$?" >> gpurun_out/r2o_bench_4gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2o_ref_8gpu.json 2> gpurun_out/r2o_ref_8gpu.err
