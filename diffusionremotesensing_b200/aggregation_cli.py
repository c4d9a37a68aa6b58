"""Command line of aggregation sampling, same flags as the reference script (Aggregation_Sampling.py:207-227):

    python -m diffusionremotesensing_b200.aggregation_cli --model_name <name> --UNet_type "Residual Attention UNet" \
        --Degradation_type DownBlur --magnification_factor 2 --patch_size 128 --stride 64 --model_input_size 256 \
        --img_lr_path scene.png --destination_path scene_sr.png

Launched under torchrun (one process per GPU) the patch list is sharded over the ranks; rank 0 writes the image.
"""
from __future__ import annotations

import argparse
import os


def main(argv=None):
    parser = argparse.ArgumentParser(description=" ")
    parser.add_argument("--noise_schedule", type=str, default="cosine")
    parser.add_argument("--snapshot_name", type=str, default="snapshot.pt")
    parser.add_argument("--noise_steps", type=int, default=1500)
    parser.add_argument("--model_input_size", type=int, default=512)
    parser.add_argument("--model_name", type=str)
    parser.add_argument("--UNet_type", type=str)
    parser.add_argument("--Degradation_type", type=str)
    parser.add_argument("--device", type=str, default="cuda")
    parser.add_argument("--magnification_factor", type=int)
    parser.add_argument("--inp_out_channels", type=int, default=3)
    parser.add_argument("--patch_size", type=int, default=64)
    parser.add_argument("--stride", type=int, default=32)
    parser.add_argument("--destination_path", type=str)
    parser.add_argument("--img_lr_path", type=str)
    args = parser.parse_args(argv)
    args.snapshot_folder_path = os.path.join(os.curdir, "models_run", args.model_name, "weights")
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        args.device = f"cuda:{local}"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from .entrypoints import launch
    launch(args)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
