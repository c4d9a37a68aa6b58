"""ORACLE -- test infrastructure only. Imports the UNMODIFIED reference modules from /root/reference.

Only usable in the build container (the GPU box has no /root/reference); tests that need it skip otherwise.
matplotlib and imageio are not installed and not touched on the sampling path, so inert stand-ins are placed in
sys.modules before the import. The generation package shadows the top-level `utils` module, so
each family is imported under its own sys.path / sys.modules state.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DRS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "UNet_model_superres.py"))


def _stub(name: str) -> None:
    if name not in sys.modules:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)


_cache = {}


def load(family: str):
    """Returns (UNet class, Diffusion class) of the reference for family in {superres, sar, generation}."""
    if family in _cache:
        return _cache[family]
    if not available():
        raise RuntimeError("reference checkout not found at " + REFERENCE_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "imageio"):
        _stub(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    saved_path = list(sys.path)
    for mod in ("utils", "UNet_model_generation", "train_diffusion_generation"):
        sys.modules.pop(mod, None)
    try:
        if family == "generation":
            sys.path.insert(0, os.path.join(REFERENCE_ROOT, "generate_new_imgs"))
            unet = importlib.import_module("UNet_model_generation").Residual_Attention_UNet_generation
            diff = importlib.import_module("train_diffusion_generation").Diffusion
        elif family == "sar":
            sys.path.insert(0, REFERENCE_ROOT)
            unet = importlib.import_module("UNet_model_SAR_TO_NDVI").Residual_Attention_UNet_SAR_TO_NDVI
            diff = importlib.import_module("train_diffusion_SAR_TO_NDVI").Diffusion
        elif family == "superres":
            sys.path.insert(0, REFERENCE_ROOT)
            unet = importlib.import_module("UNet_model_superres").Residual_Attention_UNet_superres
            diff = importlib.import_module("train_diffusion_superres").Diffusion
        else:
            raise ValueError(family)
    finally:
        sys.path[:] = saved_path
        sys.modules.pop("utils", None)
    _cache[family] = (unet, diff)
    return _cache[family]


def load_aggregation():
    if "agg" in _cache:
        return _cache["agg"]
    for name in ("matplotlib", "matplotlib.pyplot", "imageio"):
        _stub(name)
    saved_path = list(sys.path)
    try:
        sys.path.insert(0, REFERENCE_ROOT)
        cls = importlib.import_module("Aggregation_Sampling").split_aggregation_sampling
    finally:
        sys.path[:] = saved_path
    _cache["agg"] = cls
    return cls


@contextlib.contextmanager
def injected_noise(x_T, noise_fn):
    """Makes the reference's Diffusion.sample() consume prepared noise: torch.randn(shape) returns x_T and the k-th
    torch.randn_like call returns noise_fn(step) for step = noise_steps-1, noise_steps-2, ... (the reference draws
    in that order, train_diffusion_superres.py:230,243-245)."""
    import torch
    real_randn, real_randn_like = torch.randn, torch.randn_like
    state = {"step": None}

    def fake_randn(*size, **kw):
        shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
        assert tuple(x_T.shape) == shape, (x_T.shape, shape)
        return x_T.clone()

    def fake_randn_like(t, **kw):
        out = noise_fn(state["step"]).to(t.device)
        state["step"] -= 1
        return out

    def arm(noise_steps):
        state["step"] = noise_steps - 1

    torch.randn, torch.randn_like = fake_randn, fake_randn_like
    try:
        yield arm
    finally:
        torch.randn, torch.randn_like = real_randn, real_randn_like
