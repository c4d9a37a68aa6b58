set -x
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 scripts/diag_nccl_p2p.py > gpurun_out/r2r_nccl.log 2>&1
nvidia-smi topo -m > gpurun_out/r2r_topo.log 2>&1
