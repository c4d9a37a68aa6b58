"""ORACLE -- test infrastructure only. Nothing under diffusionremotesensing_b200/ may import this module.

CPU fp32 restatement (plain torch functional ops on a state_dict, no nn.Module) of the reference's sampling hot path:

    noise schedules            train_diffusion_superres.py:128-169 (identical in the SAR / generation trainers)
    pos_encoding               UNet_model_superres.py:328-335
    UNet forward, 3 families   UNet_model_superres.py:337-379, UNet_model_SAR_TO_NDVI.py:333-370,
                               generate_new_imgs/UNet_model_generation.py:296-329
    Diffusion.sample           train_diffusion_superres.py:207-255, train_diffusion_SAR_TO_NDVI.py:204-249,
                               generate_new_imgs/train_diffusion_generation.py:206-259
    patchifier / weights /     Aggregation_Sampling.py:30-74, 118-138, 76-116
    aggregation blend

Parity status: the reference has no tests, golden vectors or fixtures of its own, so this
restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF: tests/test_oracle_vs_reference.py imports the
unmodified reference modules from /root/reference (available in the build container only) and asserts bit-equality
on CPU fp32, and oracle/make_golden.py stores reference outputs as fixtures under tests/golden/ which the restatement
is checked against wherever the tests run.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# ----------------------------------------------------------------------------------------------------------------
# schedules (train_diffusion_superres.py:116-169)
# ----------------------------------------------------------------------------------------------------------------
def noise_schedule(kind: str, noise_steps: int, beta_start: float = 1e-4, beta_end: float = 0.02):
    """Returns (alpha, alpha_hat, beta) as fp32 CPU tensors, built with the same torch ops as the reference ctor."""
    if kind == "linear":
        beta = torch.linspace(beta_start, beta_end, noise_steps)          # :164
        alpha = 1.0 - beta                                                # :118
        alpha_hat = torch.cumprod(alpha, dim=0)                           # :119
    elif kind == "cosine":
        f_t = torch.cos((((torch.arange(noise_steps) / noise_steps) + 0.008) / (1 + 0.008)) * torch.pi / 2) ** 2  # :167
        alpha_hat = f_t / f_t[0]                                          # :168
        # from_alpha_hat_to_beta, :144-148: 0-d tensor arithmetic, then torch.tensor(list)
        rev = []
        for t in range(len(alpha_hat) - 1, 0, -1):
            rev.append(1 - (alpha_hat[t] / alpha_hat[t - 1]))
        rev.append(1 - alpha_hat[0])
        beta = torch.tensor(rev[::-1], dtype=alpha_hat.dtype)
        alpha = 1.0 - beta                                                # :126
    else:
        raise ValueError("noise_schedule must be 'linear' or 'cosine'")
    return alpha, alpha_hat, beta


def pos_encoding(t: Tensor, channels: int = 100) -> Tensor:
    """UNet_model_superres.py:328-335; t is [n, 1] float."""
    inv_freq = 1.0 / (10000 ** (torch.arange(0, channels, 2).float() / channels))
    a = torch.sin(t.repeat(1, channels // 2) * inv_freq)
    b = torch.cos(t.repeat(1, channels // 2) * inv_freq)
    return torch.cat([a, b], dim=-1)


# ----------------------------------------------------------------------------------------------------------------
# UNet building blocks
# ----------------------------------------------------------------------------------------------------------------
def _bn(sd: SD, p: str, x: Tensor) -> Tensor:
    # eval-mode nn.BatchNorm2d (model.eval() at train_diffusion_superres.py:227)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, 1e-5)


def _conv(sd: SD, p: str, x: Tensor, stride: int = 1, padding: int = 0) -> Tensor:
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def _time_mlp(sd: SD, p: str, t: Tensor) -> Tensor:
    # Linear -> SiLU -> Linear (UNet_model_superres.py:143-151), ReLU applied by the caller (:161, :199)
    h = F.linear(t, sd[p + ".0.weight"], sd[p + ".0.bias"])
    return F.linear(F.silu(h), sd[p + ".2.weight"], sd[p + ".2.bias"])


SKIP_CONV = {"superres": "conv_upsampled_lr_img", "sar": "conv_SAR_img", "generation": "conv_skip"}


def res_conv_block(sd: SD, p: str, x: Tensor, t: Tensor, x_skip: Optional[Tensor], taps: Optional[dict] = None,
                   tag: str = "", skip_name: str = "conv_upsampled_lr_img") -> Tensor:
    """ResConvBlock.forward, UNet_model_superres.py:153-172 (the skip conv attribute is named conv_SAR_img in
    UNet_model_SAR_TO_NDVI.py:126 and conv_skip in UNet_model_generation.py:127)."""
    h = F.relu(_bn(sd, p + ".batch_norm1", _conv(sd, p + ".conv1.0", x, padding=1)))
    if x_skip is not None:
        h = h + _conv(sd, p + "." + skip_name, x_skip, padding=1)
    te = F.relu(_time_mlp(sd, p + ".time_mlp", t))[(...,) + (None,) * 2]
    h = h + te
    if taps is not None:
        taps[tag + ".h"] = h
    h = _bn(sd, p + ".batch_norm2", _conv(sd, p + ".conv2.0", h, padding=1))
    shortcut = _bn(sd, p + ".shortcut_batch_norm", _conv(sd, p + ".shortcut_conv.0", x))
    out = F.relu(shortcut + h)
    if taps is not None:
        taps[tag + ".out"] = out
    return out


def attention_block(sd: SD, p: str, x: Tensor, g: Tensor, taps: Optional[dict] = None, tag: str = "") -> Tensor:
    """AttentionBlock.forward, UNet_model_superres.py:89-108."""
    g1 = _conv(sd, p + ".w_g.0", g)
    x1 = _conv(sd, p + ".w_x.0", x, stride=2)
    psi = torch.sigmoid(_conv(sd, p + ".psi.0", F.relu(g1 + x1)))
    if taps is not None:
        taps["psi" + tag] = psi
    up = F.interpolate(psi, scale_factor=2, mode="nearest").repeat_interleave(repeats=x.shape[1], dim=1)
    return _bn(sd, p + ".result.1", _conv(sd, p + ".result.0", up * x))


def up_conv_block(sd: SD, p: str, x: Tensor, t: Tensor, taps: Optional[dict] = None, tag: str = "") -> Tensor:
    """UpConvBlock.forward, UNet_model_superres.py:197-207."""
    te = F.relu(_time_mlp(sd, p + ".time_mlp", t))[(...,) + (None,) * 2]
    x = x + te
    x = F.relu(_bn(sd, p + ".batch_norm", _conv(sd, p + ".conv", x, padding=1)))
    if taps is not None:
        taps["uc" + tag] = x
    return F.conv_transpose2d(x, sd[p + ".transform.weight"], sd[p + ".transform.bias"], stride=2, padding=1,
                              output_padding=1)


def rrdb(sd: SD, p: str, x: Tensor) -> Tensor:
    """RRDB.forward / ResidualBlock.forward, UNet_model_superres.py:237-260."""
    out = x
    for i in range(3):
        q = f"{p}.blocks.{i}"
        r = out
        out = _conv(sd, q + ".conv2", F.relu(_conv(sd, q + ".conv1", out, padding=1)), padding=1) + r
    return _conv(sd, p + ".conv_out", out, padding=1) + x


def condition_features(sd: SD, family: str, cond: Tensor, magnification: int = 1) -> Optional[Tensor]:
    """The time-invariant branch: UNet_model_superres.py:345-353 / UNet_model_SAR_TO_NDVI.py:341-343."""
    if family == "superres":
        lr = rrdb(sd, "LR_encoder", cond)
        up = F.interpolate(lr, scale_factor=magnification, mode="bicubic")
        return _conv(sd, "conv_upsampled_lr_img", up, padding=1)
    if family == "sar":
        return _conv(sd, "conv_SAR_img", rrdb(sd, "SAR_encoder", cond), padding=1)
    return None


def unet_forward(sd: SD, family: str, x: Tensor, timestep: Tensor, cond: Optional[Tensor] = None,
                 magnification: int = 1, y: Optional[Tensor] = None, taps: Optional[dict] = None) -> Tensor:
    """One epsilon prediction. family in {"superres", "sar", "generation"}.

    superres:   UNet_model_superres.py:337-379      (cond = lr_img [1 or n, C, h, w])
    sar:        UNet_model_SAR_TO_NDVI.py:333-370   (cond = SAR_img [1 or n, C, H, W])
    generation: UNet_model_generation.py:296-329    (y = class labels [1 or n] or None)
    `taps`, if given, receives intermediate activations keyed like drs_debug_fetch names.
    """
    t = pos_encoding(timestep.unsqueeze(-1).type(torch.float), 100)
    if family == "generation" and y is not None:
        t = t + sd["label_emb.weight"][y]
    x = _conv(sd, "conv0", x, padding=1)
    feat = condition_features(sd, family, cond, magnification) if family != "generation" else None
    if feat is not None:
        x = x + feat
    x_skip = x.clone()
    if taps is not None:
        taps["h0"] = x
    residual = []
    for i in range(3):
        x = res_conv_block(sd, f"conv_blocks.{i}", x, t, x_skip if i == 0 else None, taps, f"b{i}", SKIP_CONV[family])
        residual.append(x)
        x = _conv(sd, f"downs.{i}", x, stride=2, padding=1)
        if taps is not None:
            taps[f"d{i}"] = x
    x = res_conv_block(sd, "bottle_neck", x, t, None, taps, "bn")
    for i in range(3):
        g = F.relu(_bn(sd, f"gating_signals.{i}.batch_norm", _conv(sd, f"gating_signals.{i}.conv", x)))
        att = attention_block(sd, f"attention_blocks.{i}", residual[-(i + 1)], g, taps, str(i))
        x = up_conv_block(sd, f"ups.{i}", x, t, taps, str(i))
        if taps is not None:
            taps[f"g{i}"] = g
            taps[f"att{i}"] = att
            taps[f"ut{i}"] = x
        x = _conv(sd, f"up_convs.{i}", torch.cat([x, att], dim=1), padding=1)
        if taps is not None and i < 2:
            taps[f"x{i}"] = x
    return _conv(sd, "output", x)


# ----------------------------------------------------------------------------------------------------------------
# sampling (train_diffusion_superres.py:224-255 and siblings)
# ----------------------------------------------------------------------------------------------------------------
def posterior_update(x: Tensor, eps: Tensor, noise: Tensor, alpha: Tensor, alpha_hat: Tensor, beta: Tensor) -> Tensor:
    """train_diffusion_superres.py:249 with the [n,1,1,1] coefficient tensors of :240-242."""
    return 1 / torch.sqrt(alpha) * (x - ((1 - alpha) / (torch.sqrt(1 - alpha_hat))) * eps) + torch.sqrt(beta) * noise


def lerp_cfg(uncond: Tensor, cond: Tensor, cfg_scale: float) -> Tensor:
    """generate_new_imgs/train_diffusion_generation.py:242."""
    return torch.lerp(uncond, cond, cfg_scale)


def sample(sd: SD, family: str, schedule: Tuple[Tensor, Tensor, Tensor], noise_steps: int, x_T: Tensor,
           noise_fn: Callable[[int], Optional[Tensor]], cond: Optional[Tensor] = None, magnification: int = 1,
           labels: Optional[Tensor] = None, cfg_scale: float = 0.0, start_step: Optional[int] = None,
           n_steps: Optional[int] = None, eps_fn: Optional[Callable] = None) -> Tensor:
    """Ancestral sampling loop with injected noise. noise_fn(i) returns the z of step i (the reference draws
    randn_like(x) for i > 1 and uses zeros for i == 1). cond is already batched ([1 or n, ...]).
    start_step / n_steps restrict the loop to a window (teacher-forced sub-trajectories in tests)."""
    alpha, alpha_hat, beta = schedule
    x = x_T.clone()
    n = x.shape[0]
    first = noise_steps - 1 if start_step is None else start_step
    last = 1 if n_steps is None else max(1, first - n_steps + 1)
    for i in range(first, last - 1, -1):
        t = (torch.ones(n) * i).long()
        if eps_fn is not None:
            eps = eps_fn(x, t)
        elif family == "generation":
            eps = unet_forward(sd, family, x, t, y=labels)
            if labels is not None and cfg_scale > 0:
                eps = lerp_cfg(unet_forward(sd, family, x, t, y=None), eps, cfg_scale)
        else:
            eps = unet_forward(sd, family, x, t, cond, magnification)
        a = alpha[t][:, None, None, None]
        ah = alpha_hat[t][:, None, None, None]
        b = beta[t][:, None, None, None]
        z = noise_fn(i) if i > 1 else torch.zeros_like(x)
        x = posterior_update(x, eps, z, a, ah, b)
    return x


# ----------------------------------------------------------------------------------------------------------------
# aggregation sampling (Aggregation_Sampling.py)
# ----------------------------------------------------------------------------------------------------------------
def patch_grid(height: int, width: int, patch_size: int, stride: int, magnification: int) -> List[Tuple[int, int, int, int]]:
    """SR-space windows (y0, y1, x0, x1) in the reference's order (Aggregation_Sampling.py:49-66)."""
    infos: List[Tuple[int, int, int, int]] = []
    for y in range(0, height + 1, stride):
        for x in range(0, width + 1, stride):
            ys = height - patch_size if y + patch_size > height else y
            xs = width - patch_size if x + patch_size > width else x
            info = (ys * magnification, (ys + patch_size) * magnification, xs * magnification,
                    (xs + patch_size) * magnification)
            if info not in infos:
                infos.append(info)
    return infos


def gaussian_weights(tile_width: int, tile_height: int) -> np.ndarray:
    """float64 [tile_height, tile_width] outer product (Aggregation_Sampling.py:131-137), before the fp32 cast."""
    var = 0.01
    midpoint = (tile_width - 1) / 2
    # `exp`, `sqrt`, `pi` are numpy's in the reference (Aggregation_Sampling.py:6)
    x_probs = [np.exp(-(x - midpoint) * (x - midpoint) / (tile_width * tile_width) / (2 * var)) /
               np.sqrt(2 * np.pi * var) for x in range(tile_width)]
    midpoint = tile_height / 2
    y_probs = [np.exp(-(y - midpoint) * (y - midpoint) / (tile_height * tile_height) / (2 * var)) /
               np.sqrt(2 * np.pi * var) for y in range(tile_height)]
    return np.outer(y_probs, x_probs)


def blend(patches: Sequence[Tensor], infos: Sequence[Tuple[int, int, int, int]], weight: Tensor, height: int,
          width: int) -> Tensor:
    """Aggregation_Sampling.py:91-110: sequential weighted overlap-add, divide, clamp. patches: [1, C, P, P] each;
    weight: [1, C, P, P] fp32."""
    C = patches[0].shape[1]
    im_res = torch.zeros([1, C, height, width])
    pixel_count = torch.zeros([1, C, height, width])
    for p, (y0, y1, x0, x1) in zip(patches, infos):
        im_res[:, :, y0:y1, x0:x1] += p * weight
        pixel_count[:, :, y0:y1, x0:x1] += weight
    assert torch.all(pixel_count != 0)
    im_res /= pixel_count
    return torch.clamp(im_res, 0, 1)
