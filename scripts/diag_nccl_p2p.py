"""Diagnostic: NCCL send/recv and all_reduce bandwidth between the ranks of a torchrun job + the transports NCCL picked.
    NCCL_DEBUG=INFO python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/diag_nccl_p2p.py"""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
x = torch.empty(94 << 20 >> 2, device=dev)   # 94 MB fp32
for it in range(4):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    if rank == 0:
        reqs = [dist.irecv(torch.empty_like(x), src=r) for r in range(1, world)]
        for q in reqs: q.wait()
    else:
        dist.send(x, dst=0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if rank == 0: print(f"gather of {world - 1} x 94 MB to rank 0: {(t1 - t0) * 1e3:.2f} ms = {(world - 1) * 94e-3 / (t1 - t0):.1f} GB/s", flush=True)
for it in range(3):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); dist.all_reduce(x); torch.cuda.synchronize(); t1 = time.perf_counter()
    if rank == 0: print(f"all_reduce 94 MB: {(t1 - t0) * 1e3:.2f} ms", flush=True)
if rank == 0:
    print("can_device_access_peer(0,1):", torch.cuda.can_device_access_peer(0, 1) if torch.cuda.device_count() > 1 else None)
dist.barrier(); dist.destroy_process_group()
