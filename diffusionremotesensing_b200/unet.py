"""Drop-in Residual Attention UNets whose forward runs on the sm_100a CUDA path.

The three classes keep the constructor signatures, attribute names, parameter registration order (hence the default
initialisation under a given torch seed) and state_dict keys of the reference modules

    Residual_Attention_UNet_superres      UNet_model_superres.py:266-379
    Residual_Attention_UNet_SAR_TO_NDVI   UNet_model_SAR_TO_NDVI.py:263-370
    Residual_Attention_UNet_generation    generate_new_imgs/UNet_model_generation.py:226-329

so a reference ``snapshot.pt["MODEL_STATE"]`` loads unchanged. The sub-modules only own parameters; the arithmetic of
``forward`` is done by libdrs_b200.so (tcgen05 implicit-GEMM convolutions, bf16 operands, fp32 accumulate) with
eval-mode BatchNorm semantics, which is what ``Diffusion.sample`` uses (train_diffusion_superres.py:227). There is no
PyTorch fallback: without a CUDA device or the built library ``forward`` raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch
from torch import nn

from . import _native as N

DOWN_CHANNELS = (16, 32, 64, 128, 256)
UP_CHANNELS = (256, 128, 64, 32, 16)
TIME_EMB_DIM = 100


# ----------------------------------------------------------------------------------------------------------------
# parameter containers (same registration order as the reference blocks)
# ----------------------------------------------------------------------------------------------------------------
def _time_mlp(cout: int, device) -> nn.Sequential:
    return nn.Sequential(nn.Linear(TIME_EMB_DIM, cout, device=device), nn.SiLU(), nn.Linear(cout, cout, device=device))


class AttentionBlock(nn.Module):
    """Additive attention gate parameters (UNet_model_superres.py:57-87)."""

    def __init__(self, f_g, f_x, f_int, device):
        super().__init__()
        self.w_g = nn.Sequential(nn.Conv2d(f_g, f_int, 1).to(device))
        self.w_x = nn.Sequential(nn.Conv2d(f_x, f_int, 2, stride=2).to(device))
        self.psi = nn.Sequential(nn.Conv2d(f_int, 1, 1).to(device), nn.Sigmoid())
        self.relu = nn.ReLU(inplace=False)
        self.result = nn.Sequential(nn.Conv2d(f_x, f_x, 1).to(device), nn.BatchNorm2d(f_x).to(device))


class ResConvBlock(nn.Module):
    """Residual block parameters (UNet_model_superres.py:110-151). The BatchNorm modules are registered both as
    attributes and inside the Sequentials, so both key families appear in the state_dict like in the reference."""

    def __init__(self, in_ch, out_ch, time_emb_dim, device, skip_name="conv_upsampled_lr_img"):
        super().__init__()
        self.time_mlp = _time_mlp(out_ch, device)
        self.batch_norm1 = nn.BatchNorm2d(out_ch, device=device)
        self.batch_norm2 = nn.BatchNorm2d(out_ch, device=device)
        self.shortcut_batch_norm = nn.BatchNorm2d(out_ch, device=device)
        self.relu = nn.ReLU(inplace=False)
        self.conv1 = nn.Sequential(nn.Conv2d(in_ch, out_ch, 3, padding="same", device=device), self.batch_norm1,
                                   self.relu)
        # the skip conv is called conv_upsampled_lr_img / conv_SAR_img / conv_skip in the three reference files
        setattr(self, skip_name, nn.Conv2d(in_ch, out_ch, 3, padding=1))
        self.conv2 = nn.Sequential(nn.Conv2d(out_ch, out_ch, 3, padding="same", device=device), self.batch_norm2)
        self.shortcut_conv = nn.Sequential(nn.Conv2d(in_ch, out_ch, 1, padding="same", device=device),
                                           self.shortcut_batch_norm)


class UpConvBlock(nn.Module):
    """Time add + conv + transposed conv parameters (UNet_model_superres.py:174-195)."""

    def __init__(self, in_ch, out_ch, time_emb_dim, device):
        super().__init__()
        self.time_mlp = _time_mlp(out_ch, device)
        self.batch_norm = nn.BatchNorm2d(out_ch, device=device)
        self.relu = nn.ReLU(inplace=False)
        self.conv = nn.Conv2d(in_ch, out_ch, 3, padding="same", device=device)
        self.transform = nn.ConvTranspose2d(out_ch, out_ch, 3, stride=2, padding=1, output_padding=1, device=device)


class gating_signal(nn.Module):
    """1x1 conv + BN + ReLU parameters (UNet_model_superres.py:209-220)."""

    def __init__(self, in_dim, out_dim, device):
        super().__init__()
        self.conv = nn.Conv2d(in_dim, out_dim, 1, padding="same", device=device)
        self.batch_norm = nn.BatchNorm2d(out_dim, device=device)
        self.relu = nn.ReLU(inplace=False)
        self.device = device


class ResidualBlock(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size, stride, padding)


class RRDB(nn.Module):
    """Condition encoder parameters (UNet_model_superres.py:243-254)."""

    def __init__(self, in_channels, out_channels, num_blocks=3):
        super().__init__()
        self.blocks = nn.Sequential(*[ResidualBlock(in_channels, in_channels) for _ in range(num_blocks)])
        self.conv_out = nn.Conv2d(in_channels, out_channels, 3, 1, 1)


# ----------------------------------------------------------------------------------------------------------------
# native handles
# ----------------------------------------------------------------------------------------------------------------
class _NativeHandles:
    """Owns the DrsModel / DrsPlan handles of one nn.Module and re-packs when its tensors change."""

    def __init__(self):
        self.model = None
        self.stamp = None
        self.plans: Dict[Tuple[int, int, int, int, int], int] = {}
        self.buffers: Dict[Tuple[int, int, int, int, int], Dict[str, torch.Tensor]] = {}

    def release(self):
        lib = N.lib()
        for p in self.plans.values():
            lib.drs_plan_destroy(p)
        self.plans.clear()
        self.buffers.clear()
        if self.model is not None:
            lib.drs_model_destroy(self.model)
            self.model = None
        self.stamp = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    # The handles are raw pointers owned by THIS object: a copied or unpickled module (ema = copy.deepcopy(model))
    # starts with no native state and packs its own on first use, instead of sharing -- and double-freeing -- these.
    def __deepcopy__(self, memo):
        return _NativeHandles()

    def __copy__(self):
        return _NativeHandles()

    def __reduce__(self):
        return (_NativeHandles, ())


def inv_freq_table() -> torch.Tensor:
    """1 / 10000^(2j/100) evaluated with the reference's torch ops (UNet_model_superres.py:329-331)."""
    return 1.0 / (10000 ** (torch.arange(0, TIME_EMB_DIM, 2).float() / TIME_EMB_DIM))


class _NativeUNet(nn.Module):
    """Shared trunk + the bridge to the C ABI."""

    _kind = N.MODEL_SUPERRES
    _skip_name = "conv_upsampled_lr_img"

    def _build_trunk(self, device):
        dc, uc = DOWN_CHANNELS, UP_CHANNELS
        self.conv_blocks = nn.ModuleList([ResConvBlock(dc[i], dc[i + 1], TIME_EMB_DIM, device, self._skip_name)
                                          for i in range(len(dc) - 2)])
        self.downs = nn.ModuleList([nn.Conv2d(dc[i + 1], dc[i + 1], 3, stride=2, padding=1, device=device)
                                    for i in range(len(dc) - 2)])
        self.bottle_neck = ResConvBlock(dc[-2], dc[-1], TIME_EMB_DIM, device, self._skip_name)
        self.gating_signals = nn.ModuleList([gating_signal(uc[i], uc[i + 1], device) for i in range(len(uc) - 2)])
        self.attention_blocks = nn.ModuleList([AttentionBlock(uc[i + 1], uc[i + 1], uc[i + 1], device)
                                               for i in range(len(uc) - 2)])
        self.ups = nn.ModuleList([UpConvBlock(uc[i], uc[i], TIME_EMB_DIM, device) for i in range(len(uc) - 2)])
        self.up_convs = nn.ModuleList([nn.Conv2d(int(uc[i] * 3 / 2), uc[i + 1], 3, padding=1).to(device)
                                       for i in range(len(uc) - 2)])

    # -- descriptor of the family ---------------------------------------------------------------------------
    def _desc(self) -> N.DrsModelDesc:
        raise NotImplementedError

    def pos_encoding(self, t, channels, device):
        """[sin(t f_j) | cos(t f_j)] like the reference method of the same name (UNet_model_superres.py:328-335)."""
        inv_freq = 1.0 / (10000 ** (torch.arange(0, channels, 2, device=device).float() / channels))
        return torch.cat([torch.sin(t.repeat(1, channels // 2) * inv_freq),
                          torch.cos(t.repeat(1, channels // 2) * inv_freq)], dim=-1)

    # -- native plumbing ------------------------------------------------------------------------------------
    def _handles(self) -> _NativeHandles:
        h = self.__dict__.get("_native_handles")
        if h is None:
            h = _NativeHandles()
            # not a Module attribute: stays out of state_dict; deepcopy / pickle give the copy a fresh empty instance
            self.__dict__["_native_handles"] = h
        return h

    def _stamp(self):
        return tuple((t.data_ptr(), t._version) for t in self.state_dict(keep_vars=True).values())

    def native_device(self) -> torch.device:
        dev = self.conv0.weight.device
        if dev.type != "cuda":
            raise RuntimeError(
                "the drs_b200 forward runs on a CUDA device only (parameters are on %s); there is no CPU fallback" % dev)
        return dev

    def native_model(self):
        """Packs (or re-packs after a weight change) the state_dict into a DrsModel and returns the handle."""
        h = self._handles()
        stamp = self._stamp()
        if h.model is not None and h.stamp == stamp:
            return h.model
        h.release()
        dev = self.native_device()
        lib = N.lib()
        sd = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in self.state_dict().items()
              if v.dtype.is_floating_point}
        sd["pos_encoding.inv_freq"] = inv_freq_table().contiguous()
        arr = (N.DrsTensor * len(sd))()
        keep = []
        for i, (k, v) in enumerate(sd.items()):
            name = k.encode()
            keep.append((name, v))
            arr[i].name = name
            arr[i].data = C.cast(v.data_ptr(), C.POINTER(C.c_float))
            arr[i].numel = v.numel()
        desc = self._desc()
        out = C.c_void_p()
        with torch.cuda.device(dev):
            N.check(lib.drs_model_create(C.byref(desc), arr, len(sd), dev.index or 0, C.byref(out)))
        h.model = out.value
        h.stamp = stamp
        return h.model

    def native_plan(self, nb: int, nx: int, ncond: int, S: int, mag: int):
        model = self.native_model()
        h = self._handles()
        key = (nb, nx, ncond, S, mag)
        if key not in h.plans:
            out = C.c_void_p()
            with torch.cuda.device(self.native_device()):
                N.check(N.lib().drs_plan_create(model, nb, nx, ncond, S, mag, C.byref(out)))
            h.plans[key] = out.value
        return h.plans[key]

    def sampler_buffers(self, nb: int, nx: int, ncond: int, S: int, mag: int):
        """Device buffers (x, z, eps) the sampler of a plan works in. They live as long as the plan, so the CUDA
        graphs captured on the first sample() -- which hold these addresses -- are replayed by every later one."""
        self.native_plan(nb, nx, ncond, S, mag)
        h = self._handles()
        key = (nb, nx, ncond, S, mag)
        if key not in h.buffers:
            dev = self.native_device()
            d = self._desc()
            h.buffers[key] = {
                "x": torch.empty((nx, d.x_channels, S, S), device=dev, dtype=torch.float32),
                "z": torch.empty((nx, d.x_channels, S, S), device=dev, dtype=torch.float32),
                "eps": torch.empty((nb, d.out_channels, S, S), device=dev, dtype=torch.float32),
            }
        return h.buffers[key]

    def release_native(self):
        self._handles().release()

    def _forward_native(self, x, timestep, cond, mag, labels):
        if self.training:
            raise RuntimeError("drs_b200 implements the eval-mode (sampling) forward only; call model.eval() first")
        dev = self.native_device()
        if x.device != dev:
            raise RuntimeError("input is on %s but the model is on %s" % (x.device, dev))
        n, _, S, S2 = x.shape
        if S != S2:
            raise ValueError("square inputs only")
        ncond = 0 if cond is None else cond.shape[0]
        plan = self.native_plan(n, n, ncond if ncond else 1, S, mag)
        lib = N.lib()
        st = N.stream_ptr(dev)
        x32 = x.detach().to(torch.float32).contiguous()
        t32 = timestep.detach().to(dev).type(torch.float).contiguous()
        if t32.numel() != n:
            raise ValueError("timestep must have one entry per sample")
        eps = torch.empty((n, self._desc().out_channels, S, S), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            if cond is not None:
                c32 = cond.detach().to(dev, torch.float32).contiguous()
                N.check(lib.drs_cond_encode(plan, N.ptr(c32), st))
            lab = None
            if labels is not None:
                num_classes = getattr(self, "num_classes", None) or 0
                if labels.numel() and (int(labels.min()) < 0 or int(labels.max()) >= num_classes):
                    # nn.Embedding raises for these (generate_new_imgs/UNet_model_generation.py:300-301)
                    raise IndexError("class label out of range [0, %d)" % num_classes)
                lab = labels.detach().to(dev).to(torch.int32)
                if lab.numel() == 1 and n > 1:
                    lab = lab.expand(n)
                lab = lab.contiguous()
            N.check(lib.drs_time_embed(plan, N.ptr(t32), N.ptr(lab), st))
            N.check(lib.drs_unet_forward(plan, N.ptr(x32), N.ptr(eps), st))
            N.check(lib.drs_plan_check(plan, st))
        return eps

    def debug_activation(self, plan, name: str, shape) -> torch.Tensor:
        """fp32 NCHW copy of an intermediate activation of the last forward on `plan` (tests only)."""
        dev = self.native_device()
        out = torch.empty(shape, device=dev, dtype=torch.float32)
        r = N.lib().drs_debug_fetch(plan, name.encode(), N.ptr(out), out.numel(), N.stream_ptr(dev))
        if r < 0:
            N.check(int(r))
        if r != out.numel():
            raise ValueError("activation %s has %d elements, expected %d" % (name, r, out.numel()))
        torch.cuda.synchronize(dev)
        return out


# ----------------------------------------------------------------------------------------------------------------
# the three model families
# ----------------------------------------------------------------------------------------------------------------
class Residual_Attention_UNet_superres(_NativeUNet):
    _kind = N.MODEL_SUPERRES

    def __init__(self, image_channels=3, out_dim=3, device=None):
        super().__init__()
        self.image_channels = image_channels
        self.down_channels = DOWN_CHANNELS
        self.up_channels = UP_CHANNELS
        self.out_dim = out_dim
        self.time_emb_dim = TIME_EMB_DIM
        self.device = device
        self.conv0 = nn.Conv2d(image_channels, DOWN_CHANNELS[0], 3, padding=1)
        self.LR_encoder = RRDB(in_channels=image_channels, out_channels=image_channels, num_blocks=3)
        self.conv_upsampled_lr_img = nn.Conv2d(image_channels, DOWN_CHANNELS[0], 3, padding=1)
        self._build_trunk(device)
        self.output = nn.Conv2d(UP_CHANNELS[-2], out_dim, 1)

    def _desc(self):
        return N.DrsModelDesc(N.MODEL_SUPERRES, self.image_channels, self.image_channels, self.out_dim, 0)

    def forward(self, x, timestep, lr_img, magnification_factor):
        """eps = UNet(x, t | lr_img): same positional signature as UNet_model_superres.py:337."""
        mag = int(magnification_factor)
        if mag != magnification_factor or mag < 1:
            raise ValueError("magnification_factor must be a positive integer")
        if lr_img.dim() != 4 or lr_img.shape[0] not in (1, x.shape[0]):
            raise ValueError("lr_img must be [1 or n, C, h, w]")
        if lr_img.shape[2] * mag != x.shape[2] or lr_img.shape[3] * mag != x.shape[3]:
            raise ValueError("lr_img size times the magnification must equal the size of x")
        return self._forward_native(x, timestep, lr_img, mag, None)


class Residual_Attention_UNet_SAR_TO_NDVI(_NativeUNet):
    _kind = N.MODEL_SAR_TO_NDVI
    _skip_name = "conv_SAR_img"

    def __init__(self, SAR_channels=2, NDVI_channels=1, device=None):
        super().__init__()
        self.SAR_channels = SAR_channels
        self.NDVI_channels = NDVI_channels
        self.down_channels = DOWN_CHANNELS
        self.up_channels = UP_CHANNELS
        self.time_emb_dim = TIME_EMB_DIM
        self.device = device
        self.conv0 = nn.Conv2d(NDVI_channels, DOWN_CHANNELS[0], 3, padding=1)
        self.SAR_encoder = RRDB(in_channels=SAR_channels, out_channels=SAR_channels, num_blocks=3)
        self.conv_SAR_img = nn.Conv2d(SAR_channels, DOWN_CHANNELS[0], 3, padding=1)
        self._build_trunk(device)
        self.output = nn.Conv2d(UP_CHANNELS[-2], NDVI_channels, 1)

    def _desc(self):
        return N.DrsModelDesc(N.MODEL_SAR_TO_NDVI, self.NDVI_channels, self.SAR_channels, self.NDVI_channels, 0)

    def forward(self, NDVI_img, timestep, SAR_img):
        """Same positional signature as UNet_model_SAR_TO_NDVI.py:333."""
        if SAR_img.dim() != 4 or SAR_img.shape[0] not in (1, NDVI_img.shape[0]):
            raise ValueError("SAR_img must be [1 or n, C, H, W]")
        if SAR_img.shape[2:] != NDVI_img.shape[2:]:
            raise ValueError("SAR_img and NDVI_img must have the same spatial size")
        return self._forward_native(NDVI_img, timestep, SAR_img, 1, None)


class Residual_Attention_UNet_generation(_NativeUNet):
    _kind = N.MODEL_GENERATION
    _skip_name = "conv_skip"

    def __init__(self, image_channels=3, out_dim=3, num_classes=None, device=None):
        super().__init__()
        self.image_channels = image_channels
        self.down_channels = DOWN_CHANNELS
        self.up_channels = UP_CHANNELS
        self.out_dim = out_dim
        self.time_emb_dim = TIME_EMB_DIM
        self.device = device
        self.num_classes = num_classes
        self.conv0 = nn.Conv2d(image_channels, DOWN_CHANNELS[0], 3, padding=1)
        self._build_trunk(device)
        self.output = nn.Conv2d(UP_CHANNELS[-2], out_dim, 1)
        if num_classes is not None:
            self.label_emb = nn.Embedding(num_classes, TIME_EMB_DIM).to(device=device)

    def _desc(self):
        return N.DrsModelDesc(N.MODEL_GENERATION, self.image_channels, 0, self.out_dim, self.num_classes or 0)

    def forward(self, x, timestep, y=None):
        """Same positional signature as generate_new_imgs/UNet_model_generation.py:296."""
        if y is not None and self.num_classes is None:
            raise ValueError("labels given but the model was built without num_classes")
        return self._forward_native(x, timestep, None, 1, y)
