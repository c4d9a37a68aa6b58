"""Synthetic weights, inputs and parity metrics of the sampling path: what `bench.py`,
`__graft_entry__.smoke()` and the test-suite feed the kernels. There are no pre-trained weights (the reference's
snapshot.pt files are absent), so everything is random-init with randomised BatchNorm statistics."""
from __future__ import annotations

import numpy as np
import torch

FAMILIES = ("superres", "sar", "generation")

def build_model(family: str, device="cpu", num_classes=10):
    """Our drop-in module of a family with default ctor arguments."""
    from . import unet as D
    if family == "superres":
        return D.Residual_Attention_UNet_superres(3, 3, device)
    if family == "sar":
        return D.Residual_Attention_UNet_SAR_TO_NDVI(2, 1, device)
    return D.Residual_Attention_UNet_generation(3, 3, num_classes, device)


def synthetic_state_dict(model: torch.nn.Module, seed: int) -> dict:
    """Deterministic, torch-version-independent weights for every state_dict entry of `model`: numpy PCG64 streams
    keyed by (seed, entry index). Conv / Linear weights ~ U(-b, b) with b = sqrt(3 / fan_in) (unit-gain, so
    activations neither vanish nor explode through the 30 layers), biases ~ U(-0.1, 0.1); BatchNorm weight ~ U(0.5,
    1.5), bias ~ N(0, 0.1), running_mean ~ N(0, 0.1), running_var ~ U(0.5, 1.5) so the affine fold is exercised
    (SURVEY.md section 0)."""
    out = {}
    ref = model.state_dict()
    for idx, (k, v) in enumerate(ref.items()):
        if k in out:
            continue
        rng = np.random.Generator(np.random.PCG64([seed, idx]))
        shape = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros((), dtype=torch.long)
            continue
        leaf = k.rsplit(".", 1)[-1]
        is_bn = leaf in ("running_mean", "running_var") or (k.rsplit(".", 1)[0] + ".running_mean") in ref
        if is_bn:
            if leaf in ("weight", "running_var"):
                a = rng.uniform(0.5, 1.5, size=shape)
            else:
                a = rng.normal(0.0, 0.1, size=shape)
        elif k == "label_emb.weight":
            a = rng.normal(0.0, 1.0, size=shape)
        elif leaf == "weight":
            fan_in = int(np.prod(shape[1:])) if "transform" not in k else shape[0] * 9 // 4 + 1
            a = rng.uniform(-1.0, 1.0, size=shape) * np.sqrt(3.0 / fan_in)
        else:
            a = rng.uniform(-0.1, 0.1, size=shape)
        out[k] = torch.from_numpy(a.astype(np.float32))
    # aliased BatchNorm entries (batch_norm1 <-> conv1.1 ...) must carry identical values
    model.load_state_dict(out)
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def default_init_model(family: str, seed: int = 0, bn_seed: int = 1):
    """SURVEY.md section 8d synthetic weights: the constructor's default initialisation under torch.manual_seed(seed),
    then BatchNorm weight ~ U(0.5, 1.5), bias ~ N(0, 0.1), running_mean ~ N(0, 0.1), running_var ~ U(0.5, 1.5)."""
    torch.manual_seed(seed)
    m = build_model(family)
    g = torch.Generator().manual_seed(bn_seed)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.uniform_(0.5, 1.5, generator=g)
                mod.bias.normal_(0.0, 0.1, generator=g)
                mod.running_mean.normal_(0.0, 0.1, generator=g)
                mod.running_var.uniform_(0.5, 1.5, generator=g)
    return m, {k: v.detach().clone() for k, v in m.state_dict().items()}


def np_randn(seed, *shape) -> torch.Tensor:
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.standard_normal(size=shape).astype(np.float32))


def np_rand(seed, *shape) -> torch.Tensor:
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.random(size=shape).astype(np.float32))


def max_rel_err(test: torch.Tensor, ref: torch.Tensor) -> float:
    """max|test - ref| / max|ref| -- the per-step UNet metric of BASELINE.json's north_star (tolerance 2e-2)."""
    return ((test.double().cpu() - ref.double().cpu()).abs().max() / ref.double().abs().max().clamp_min(1e-30)).item()


def psnr_ref_range(test: torch.Tensor, ref: torch.Tensor) -> float:
    """PSNR with peak := max(1, max(ref) - min(ref)) on the unclamped sample() output."""
    ref = ref.double().cpu()
    mse = ((test.double().cpu() - ref) ** 2).mean().item()
    peak = max(1.0, (ref.max() - ref.min()).item())
    if mse == 0:
        return float("inf")
    return 10.0 * np.log10(peak * peak / mse)
