"""One-call entry points on top of the CUDA sampling path.

    super_resolver            superres_and_NDVIgen.py:14-51   (model-name parsing, cosine / 1500 steps, clamp)
    SAR_to_NDVI_generator     superres_and_NDVIgen.py:85-119  (SAR range normalisation [-1, 1] -> [0, 1])
    generate_per_class        generate_new_imgs/imgs_generator.py:27-45 (one sample per class, clamped)
    load_scene / save_scene   Aggregation_Sampling.py:170-191,204-205 (PIL open, Image.BICUBIC square resize to the
                              nearest size of the fixed list, ToTensor; ToPILImage().save)
    prepare_scene             the same square-resize rule for a scene that is already a tensor
    aggregation_super_resolver  Aggregation_Sampling.py:193-203 (patch split, batched / sharded sampling, blend)
    launch                    Aggregation_Sampling.py:140-205 (the CLI body: file in -> file out)

Same arguments and return values as the reference functions; `noise_steps`, `snapshot_root` and the noise hooks are
optional keyword extensions (the reference hard-codes 1500 steps and ./models_run).
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from .aggregation import split_aggregation_sampling
from .diffusion import Diffusion, Diffusion_SAR_TO_NDVI, Diffusion_generation
from .unet import (Residual_Attention_UNet_SAR_TO_NDVI, Residual_Attention_UNet_generation,
                   Residual_Attention_UNet_superres)

SCENE_SIZES = [64, 128, 256, 512, 1024, 2048, 4096, 8192, 10000]  # Aggregation_Sampling.py:178


def parse_model_name(model_name: str):
    """('..._magnification2_LRimgsize128_...') -> (magnification, HR image size); superres_and_NDVIgen.py:30-31."""
    parts = model_name.split("_")
    mags = [p[len("magnification"):] for p in parts if p.startswith("magnification")]
    sizes = [p[len("LRimgsize"):] for p in parts if p.startswith("LRimgsize")]
    if not mags or not sizes:
        raise IndexError("model_name must contain 'magnification<k>' and 'LRimgsize<n>' fields separated by '_'")
    k = int(mags[0])
    return k, int(sizes[0]) * k


def super_resolver(lr_img, device, model_name, *, noise_steps=1500, snapshot_root="models_run", x_T=None, noise=None):
    k, image_size = parse_model_name(model_name)
    channels = lr_img.shape[0]
    model = Residual_Attention_UNet_superres(channels, channels, device).to(device)
    snapshot_path = os.path.join(snapshot_root, model_name, "weights", "snapshot.pt")
    print(f"HR Image size: {image_size}, LR Image size: {image_size // k} Magnification factor: {k}, Channels: {channels}")
    diffusion = Diffusion(noise_schedule="cosine", model=model, snapshot_path=snapshot_path, noise_steps=noise_steps,
                          beta_start=1e-4, beta_end=0.02, magnification_factor=k, device=device, image_size=image_size,
                          model_name=model_name, Degradation_type="DownBlur")
    sr = diffusion.sample(n=1, model=model, lr_img=lr_img, input_channels=channels, generate_video=False, x_T=x_T,
                          noise=noise)
    return torch.clamp(sr, 0, 1)


def normalise_sar(SAR_img: torch.Tensor) -> torch.Tensor:
    """superres_and_NDVIgen.py:106-109."""
    lo = SAR_img.min()
    if lo < 0 and lo > -1:
        return (SAR_img + 1) / 2
    if lo < -1 or SAR_img.max() > 1:
        raise ValueError("SAR image values are not in the range [-1, 1]")
    return SAR_img


def SAR_to_NDVI_generator(SAR_img_path, device, n_generations=1, *, noise_steps=1500, snapshot_root="models_run",
                          x_T=None, noise=None):
    model_name = "Residual_Attention_UNet_EMA_imgsize128_SAR_TO_NDVI"
    model = Residual_Attention_UNet_SAR_TO_NDVI(2, 1, device).to(device)
    snapshot_path = os.path.join(snapshot_root, model_name, "weights", "snapshot.pt")
    image_size = int([p[len("imgsize"):] for p in model_name.split("_") if p.startswith("imgsize")][0])
    print(f"Image size: {image_size}, SAR channels: 2, NDVI channels: 1")
    SAR_img = SAR_img_path if torch.is_tensor(SAR_img_path) else torch.load(SAR_img_path)
    SAR_img = normalise_sar(SAR_img)
    diffusion = Diffusion_SAR_TO_NDVI(noise_schedule="cosine", model=model, snapshot_path=snapshot_path,
                                      noise_steps=noise_steps, beta_start=1e-4, beta_end=0.02, device=device,
                                      image_size=image_size, model_name=model_name, multiple_gpus=False,
                                      ema_smoothing=False)
    return diffusion.sample(n=n_generations, model=model, SAR_img=SAR_img, NDVI_channels=1, generate_video=False,
                            x_T=x_T, noise=noise)


def generate_per_class(model: Residual_Attention_UNet_generation, diffusion: Diffusion_generation, num_classes: int,
                       cfg_scale=3, input_channels=3, *, x_T=None, noise=None):
    """One image per class, clamped to [0, 1]: generate_new_imgs/imgs_generator.py:38-40 calls
    ``sample(n=1, target_class=tensor([i]))`` then ``clamp(0, 1)`` for i = 0..num_classes-1. Here the classes run as
    ONE batch; the start states are drawn one ``randn((1, C, S, S))`` per class from the CPU default generator, in
    class order, which is exactly the sequence of x_T tensors the reference's loop draws. (Its per-step device noise
    stream cannot be reproduced by concurrent chains; ``noise(class_index, step) -> [1, C, S, S]`` injects it.)
    Returns [num_classes, C, S, S]."""
    S = diffusion.image_size
    if x_T is None:
        x_T = torch.cat([torch.randn((1, input_channels, S, S)) for _ in range(num_classes)], dim=0)
    labels = torch.arange(num_classes, dtype=torch.long)
    nz = None if noise is None else (lambda step: torch.cat([noise(i, step) for i in range(num_classes)], dim=0))
    out = diffusion.sample(n=num_classes, model=model, target_class=labels, cfg_scale=cfg_scale,
                           input_channels=input_channels, x_T=x_T, noise=nz)
    return out.clamp(0, 1)


def nearest_scene_size(width: int, height: int) -> int:
    """Aggregation_Sampling.py:176-185: the size of the fixed list with the smallest |s - w| + |s - h| (first wins)."""
    dist = [abs(s - width) + abs(s - height) for s in SCENE_SIZES]
    return SCENE_SIZES[dist.index(min(dist))]


def load_scene(img_lr_path) -> torch.Tensor:
    """Image file -> [1, C, N, N] fp32 in [0, 1] the way Aggregation_Sampling.py:170-191 does it: PIL open, a
    non-square image is resized with ``Image.BICUBIC`` (PIL's uint8 filter, a = -0.5 -- not ``F.interpolate``) to the
    nearest size of the fixed list, then torchvision's ``ToTensor``. Host-side, one-off."""
    from PIL import Image
    from torchvision import transforms
    img = Image.open(img_lr_path)
    if img.size[0] != img.size[1]:
        n = nearest_scene_size(img.size[0], img.size[1])
        print(f"The image must be square but it is {img.size[0], img.size[1]}! It will be resized to {n}x{n}")
        img = img.resize((n, n), Image.BICUBIC)
    return transforms.ToTensor()(img).unsqueeze(0)


def save_scene(final_pred: torch.Tensor, destination_path) -> None:
    """[1, C, H, W] in [0, 1] -> image file (Aggregation_Sampling.py:204-205: ToPILImage on the CPU tensor, save)."""
    from torchvision import transforms
    transforms.ToPILImage()(final_pred.squeeze(0).cpu()).save(destination_path)


def prepare_scene(img: torch.Tensor) -> torch.Tensor:
    """[C, H, W] fp32 in [0, 1] -> [1, C, N, N]. A non-square scene goes through the same PIL path as `load_scene`
    (8-bit quantisation, ``Image.BICUBIC``), because that is what the reference's CLI applies to its input file."""
    c, h, w = img.shape
    if h == w:
        return img.unsqueeze(0)
    from PIL import Image
    from torchvision import transforms
    n = nearest_scene_size(w, h)
    print(f"The image must be square but it is {(w, h)}! It will be resized to {n}x{n}")
    pil = transforms.ToPILImage()(img.cpu().clamp(0, 1)).resize((n, n), Image.BICUBIC)
    return transforms.ToTensor()(pil).unsqueeze(0)


def aggregation_super_resolver(img_lr: torch.Tensor, model, diffusion: Diffusion, patch_size: int, stride: int,
                               device, patch_batch: int = 128, *, noise=None, x_T=None) -> torch.Tensor:
    """Whole-scene super-resolution: split, sample every patch on this rank's share, gather, blend."""
    scene = (prepare_scene(img_lr) if img_lr.dim() == 3 else img_lr).to(device)
    agg = split_aggregation_sampling(scene, patch_size, stride, diffusion.magnification_factor, diffusion, device,
                                     patch_batch=patch_batch)
    return agg.aggregation_sampling(noise=noise, x_T=x_T)


def launch(args, *, noise=None, x_T=None, patch_batch: int = 128) -> torch.Tensor:
    """Body of the reference CLI (Aggregation_Sampling.py:140-205): `args` carries the same attributes
    (snapshot_folder_path, snapshot_name, magnification_factor, inp_out_channels, noise_schedule, device,
    model_input_size, noise_steps, model_name, Degradation_type, patch_size, stride, destination_path, img_lr_path,
    UNet_type). Reads the scene, samples every patch (sharded over the ranks when torch.distributed is initialised),
    blends and writes the image (rank 0). Returns the blended scene."""
    snapshot_path = os.path.join(args.snapshot_folder_path, args.snapshot_name)
    channels = args.inp_out_channels
    if args.UNet_type.lower() != "residual attention unet":
        raise ValueError("UNet_type must be 'Residual Attention UNet'")
    model = Residual_Attention_UNet_superres(channels, channels, args.device).to(args.device)
    print(f"You are using {args.UNet_type} model")
    img_lr = load_scene(args.img_lr_path).to(args.device)
    diffusion = Diffusion(noise_schedule=args.noise_schedule, model=model, snapshot_path=snapshot_path,
                          noise_steps=args.noise_steps, beta_start=1e-4, beta_end=0.02,
                          magnification_factor=args.magnification_factor, device=args.device,
                          image_size=args.model_input_size, model_name=args.model_name,
                          Degradation_type=args.Degradation_type, multiple_gpus=False, ema_smoothing=False)
    agg = split_aggregation_sampling(img_lr, args.patch_size, args.stride, args.magnification_factor, diffusion,
                                     args.device, patch_batch=patch_batch)
    final_pred = agg.aggregation_sampling(noise=noise, x_T=x_T)
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0:
        save_scene(final_pred, args.destination_path)
    return final_pred
