"""B200-native reverse-diffusion sampling path of AdrianoEttari/DiffusionRemoteSensing.

Python surface = the reference's (same class names, constructor and call signatures); arithmetic = hand-written
sm_100a CUDA behind the C ABI in include/drs_b200.h (libdrs_b200.so, built in-tree by __graft_entry__.build()).
"""
from .unet import (Residual_Attention_UNet_superres, Residual_Attention_UNet_SAR_TO_NDVI,
                   Residual_Attention_UNet_generation)
from .diffusion import Diffusion, Diffusion_SAR_TO_NDVI, Diffusion_generation
from .aggregation import split_aggregation_sampling, partition_blocks, gather_blocks, blend_patches

from .entrypoints import (super_resolver, SAR_to_NDVI_generator, generate_per_class, prepare_scene,
                          aggregation_super_resolver, parse_model_name, normalise_sar, load_scene, save_scene,
                          nearest_scene_size, launch)

Diffusion_superres = Diffusion

__all__ = [
    "Residual_Attention_UNet_superres", "Residual_Attention_UNet_SAR_TO_NDVI", "Residual_Attention_UNet_generation",
    "Diffusion", "Diffusion_superres", "Diffusion_SAR_TO_NDVI", "Diffusion_generation",
    "split_aggregation_sampling", "partition_blocks", "gather_blocks", "blend_patches",
    "super_resolver", "SAR_to_NDVI_generator", "generate_per_class", "prepare_scene", "aggregation_super_resolver",
    "parse_model_name", "normalise_sar", "load_scene", "save_scene", "nearest_scene_size", "launch",
]
