"""One-call entry points on top of the CUDA sampling path.

    super_resolver            superres_and_NDVIgen.py:14-51   (model-name parsing, cosine / 1500 steps, clamp)
    SAR_to_NDVI_generator     superres_and_NDVIgen.py:85-119  (SAR range normalisation [-1, 1] -> [0, 1])
    generate_per_class        generate_new_imgs/imgs_generator.py:27-45 (one sample per class)
    prepare_scene             Aggregation_Sampling.py:171-191 (non-square scene -> nearest size of the fixed list)
    aggregation_super_resolver  Aggregation_Sampling.py:193-203 (patch split, batched / sharded sampling, blend)

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
                       cfg_scale=3, input_channels=3):
    """One image per class in ONE batched sample() call (the reference script loops sample(n=1, ...) per class,
    generate_new_imgs/imgs_generator.py:27-39); returns [num_classes, C, S, S]."""
    labels = torch.arange(num_classes, dtype=torch.long)
    return diffusion.sample(n=num_classes, model=model, target_class=labels, cfg_scale=cfg_scale,
                            input_channels=input_channels)


def prepare_scene(img: torch.Tensor) -> torch.Tensor:
    """[C, H, W] in [0, 1] -> [1, C, N, N]; a non-square scene is resized (bicubic) to the nearest size of the fixed
    list like Aggregation_Sampling.py:171-188 does with PIL."""
    c, h, w = img.shape
    if h != w:
        dist = [abs(s - w) + abs(s - h) for s in SCENE_SIZES]
        n = SCENE_SIZES[dist.index(min(dist))]
        print(f"The image must be square but it is {(w, h)}! It will be resized to {n}x{n}")
        img = torch.nn.functional.interpolate(img.unsqueeze(0), size=(n, n), mode="bicubic", align_corners=False)[0]
        img = img.clamp(0, 1)
    return img.unsqueeze(0)


def aggregation_super_resolver(img_lr: torch.Tensor, model, diffusion: Diffusion, patch_size: int, stride: int,
                               device, patch_batch: int = 32) -> torch.Tensor:
    """Whole-scene super-resolution: split, sample every patch on this rank's share, gather, blend."""
    scene = prepare_scene(img_lr).to(device)
    agg = split_aggregation_sampling(scene, patch_size, stride, diffusion.magnification_factor, diffusion, device,
                                     patch_batch=patch_batch)
    return agg.aggregation_sampling()
