"""Aggregation sampling sharded over the ranks of a torchrun job (one process per GPU, NCCL gather to rank 0),
checked on rank 0 against the same scene sampled on a single GPU with the same injected noise.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/run_aggregation_dist.py
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import common as T
import diffusionremotesensing_b200 as D

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
steps, P, stride, k, LR = 20, 64, 32, 2, 256           # 49 patches of 64 -> 128, scene 256 -> 512
model, _ = T.default_init_model("superres")
model.to(dev)
diff = D.Diffusion("cosine", model, "/nonexistent", noise_steps=steps, device=str(dev), magnification_factor=k,
                   image_size=P * k, Degradation_type="DownBlur")
img = T.np_rand(77, 1, 3, LR, LR).to(dev)
x_T = lambda p: T.np_randn(7000 + p, 1, 3, P * k, P * k)                  # noqa: E731
noise = lambda p, i: T.np_randn(9000 + 100 * p + i, 1, 3, P * k, P * k)   # noqa: E731
agg = D.split_aggregation_sampling(img, P, stride, k, diff, str(dev), patch_batch=8)
torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
sharded = agg.aggregation_sampling(noise=noise, x_T=x_T)
torch.cuda.synchronize(); t1 = time.time()
if rank == 0:
    single = D.split_aggregation_sampling(img, P, stride, k, diff, str(dev), patch_batch=8)
    dist_ok = True
    # single-GPU run of the same scene: bypass the process group by sampling all patches locally
    patches = single.sample_patches(range(len(single.patches_lr)), noise, x_T)
    ref, _ = D.blend_patches(patches, single.patches_sr_infos, single.weight[0, 0], LR * k, LR * k)
    print(f"world={world} patches={len(agg.patches_lr)} blocks={D.partition_blocks(len(agg.patches_lr), world)} "
          f"sharded {t1 - t0:.2f}s  bit-identical to single GPU: {bool(torch.equal(sharded, ref))}  "
          f"max diff {float((sharded - ref).abs().max()):.3e}")
dist.barrier()
dist.destroy_process_group()
