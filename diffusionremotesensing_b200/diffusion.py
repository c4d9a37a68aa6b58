"""Drop-in ``Diffusion`` classes whose ``sample()`` runs the whole reverse process on the sm_100a CUDA path.

Constructor arguments, public attributes (``model``, ``alpha``, ``alpha_hat``, ``beta``, ``noise_steps``,
``image_size``, ``magnification_factor``, ``device``), the ``sample`` signatures, the RNG consumption (x_T from the
CPU default generator, one device ``normal_`` per step with i > 1), the error behaviour and the side effect of
leaving the model in train() mode follow

    superres    train_diffusion_superres.py:78-255
    SAR->NDVI   train_diffusion_SAR_TO_NDVI.py:79-249
    generation  generate_new_imgs/train_diffusion_generation.py:81-259

Training (``train``, losses, EMA, DDP) is out of scope. Two optional keyword arguments extend
the reference signatures without changing positional use: ``x_T`` (start state) and ``noise`` (callable
``noise(i) -> [n, C, S, S]`` used instead of the device generator), which the parity tests use to inject identical
noise into this path and the oracle.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional

import torch
from torch import nn

from . import _native as N


def _schedule(noise_schedule: str, noise_steps: int, beta_start: float, beta_end: float):
    """alpha, alpha_hat, beta (fp32, CPU) with the reference's own torch ops so the tables are bit-identical
    (train_diffusion_superres.py:116-169)."""
    if noise_schedule == "linear":
        beta = torch.linspace(beta_start, beta_end, noise_steps)
        alpha = 1.0 - beta
        alpha_hat = torch.cumprod(alpha, dim=0)
    elif noise_schedule == "cosine":
        steps = torch.arange(noise_steps) / noise_steps
        f_t = torch.cos(((steps + 0.008) / (1 + 0.008)) * torch.pi / 2) ** 2
        alpha_hat = f_t / f_t[0]
        ratios = [1 - alpha_hat[0]] + [1 - (alpha_hat[t] / alpha_hat[t - 1]) for t in range(1, noise_steps)]
        beta = torch.tensor(ratios, dtype=alpha_hat.dtype)
        alpha = 1.0 - beta
    else:
        raise ValueError("noise_schedule must be 'linear' or 'cosine', got %r" % (noise_schedule,))
    return alpha, alpha_hat, beta


class _DiffusionBase:
    _default_model_name = "superres"

    def _init_common(self, noise_schedule, model, snapshot_path, noise_steps, beta_start, beta_end, device, image_size,
                     model_name, multiple_gpus, ema_smoothing):
        self.noise_steps = noise_steps
        self.beta_start = beta_start
        self.beta_end = beta_end
        self.image_size = image_size
        self.model_name = model_name
        self.device = device
        self.snapshot_path = snapshot_path
        self.multiple_gpus = multiple_gpus
        self.ema_smoothing = ema_smoothing
        self.model = model.to(self.device)
        self.epochs_run = 0
        if snapshot_path is not None and os.path.exists(snapshot_path):
            print("Loading snapshot")
            self._load_snapshot()
        self.noise_schedule = noise_schedule
        alpha, alpha_hat, beta = _schedule(noise_schedule, noise_steps, beta_start, beta_end)
        self._host_tables = (alpha, alpha_hat, beta)
        self.alpha = alpha.to(self.device)
        self.alpha_hat = alpha_hat.to(self.device)
        self.beta = beta.to(self.device)

    # -- snapshot wire format (train_diffusion_superres.py:257-308) ---------------------------------------------
    def _load_snapshot(self):
        snapshot = torch.load(self.snapshot_path, map_location="cpu")
        state = snapshot["MODEL_STATE"]
        target = self.model.module if hasattr(self.model, "module") else self.model
        state = type(state)((k.replace("module.", ""), v) for k, v in state.items())
        target.load_state_dict(state)
        target.to(self.device)
        self.epochs_run = snapshot["EPOCHS_RUN"]
        print(f"Resuming training from snapshot at Epoch {self.epochs_run}")

    def _save_snapshot(self, epoch, model):
        target = model.module if hasattr(model, "module") else model
        torch.save({"MODEL_STATE": target.state_dict(), "EPOCHS_RUN": epoch}, self.snapshot_path)
        print(f"Epoch {epoch} | Training snapshot saved at {self.snapshot_path}")

    # -- forward process ---------------------------------------------------------
    def noise_images(self, x, t):
        """x_t = sqrt(alpha_hat[t]) * x + sqrt(1 - alpha_hat[t]) * eps, eps ~ N(0, 1) drawn like the reference
        (train_diffusion_superres.py:171-190). Returns (x_t, eps). Runs the CUDA kernel when x is on a CUDA device."""
        sqrt_ah = torch.sqrt(self.alpha_hat[t])
        sqrt_1m = torch.sqrt(1 - self.alpha_hat[t])
        epsilon = torch.randn_like(x, dtype=torch.float32)
        if x.device.type != "cuda":
            raise RuntimeError("drs_b200 noise_images runs on a CUDA device only; there is no CPU fallback")
        x32 = x.to(torch.float32).contiguous()
        out = torch.empty_like(x32)
        per_sample = x32[0].numel()
        with torch.cuda.device(x.device):
            N.check(N.lib().drs_noise_images(N.ptr(x32), N.ptr(epsilon.contiguous()),
                                             N.ptr(sqrt_ah.to(x.device, torch.float32).contiguous()),
                                             N.ptr(sqrt_1m.to(x.device, torch.float32).contiguous()), N.ptr(out),
                                             x32.shape[0], per_sample, N.stream_ptr(x.device)))
        return out, epsilon

    def sample_timesteps(self, n):
        """Uniform integer timesteps in [1, noise_steps) (train_diffusion_superres.py:192-205)."""
        return torch.randint(low=1, high=self.noise_steps, size=(n,))

    # -- per-step coefficients ---------------------------------------------------------------------------------
    def _coefficients(self):
        """c1 = 1/sqrt(alpha), c2 = (1-alpha)/sqrt(1-alpha_hat), c3 = sqrt(beta) as the reference evaluates them
        inside the update expression (train_diffusion_superres.py:249), fp32 on the host."""
        alpha, alpha_hat, beta = self._host_tables
        c1 = 1 / torch.sqrt(alpha)
        c2 = (1 - alpha) / (torch.sqrt(1 - alpha_hat))
        c3 = torch.sqrt(beta)
        return c1.contiguous(), c2.contiguous(), c3.contiguous()

    def _native_device(self) -> torch.device:
        dev = torch.device(self.device)
        if dev.type != "cuda":
            raise RuntimeError("drs_b200 samples on a CUDA device only (device=%r); there is no CPU fallback"
                               % (self.device,))
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        return dev

    # -- the reverse process -----------------------------------------------------------------------------------
    def _reverse_process(self, model, x: torch.Tensor, cond: Optional[torch.Tensor], mag: int,
                         labels: Optional[torch.Tensor], cfg: bool, cfg_scale: float,
                         noise: Optional[Callable[[int], torch.Tensor]], frames: Optional[List[torch.Tensor]],
                         use_graph: bool = True, start_step: Optional[int] = None,
                         n_steps: Optional[int] = None, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """x: [nx, C, S, S] fp32 on the device (updated in place and returned); cond: [1 or nx, Cc, h, w] or None;
        labels: int32 [nx] or None; cfg: run the conditional and unconditional passes as one batch of 2 nx."""
        dev = x.device
        lib = N.lib()
        nx, _, S, _ = x.shape
        nb = 2 * nx if cfg else nx
        ncond = 1 if cond is None else cond.shape[0]
        if cfg and cond is not None:
            raise ValueError("classifier-free guidance is defined for the generation family only")
        plan = model.native_plan(nb, nx, ncond, S, mag)
        st = N.stream_ptr(dev)
        c1, c2, c3 = self._coefficients()
        lab_host = None
        if labels is not None:
            lab = labels.to("cpu", torch.int32).reshape(-1)
            if lab.numel() == 1 and nx > 1:
                lab = lab.expand(nx)
            if lab.numel() != nx:
                raise ValueError("target_class must have 1 or n entries")
            if cfg:
                lab = torch.cat([lab, torch.full((nx,), -1, dtype=torch.int32)])
            lab_host = lab.contiguous()
        # the sampler works in buffers owned by the plan (same addresses on every call -> captured graphs stay valid)
        bufs = model.sampler_buffers(nb, nx, ncond, S, mag)
        if tuple(bufs["x"].shape) != tuple(x.shape):
            raise ValueError("state has shape %s, the model expects %s" % (tuple(x.shape), tuple(bufs["x"].shape)))
        xs, zbuf, eps = bufs["x"], bufs["z"], bufs["eps"]
        xs.copy_(x)
        first = self.noise_steps - 1 if start_step is None else start_step
        last = 1 if n_steps is None else max(1, first - n_steps + 1)
        with torch.cuda.device(dev):
            if cond is not None:
                N.check(lib.drs_cond_encode(plan, N.ptr(cond), st))
            N.check(lib.drs_sampler_prepare(plan, self.noise_steps, N.ptr(c1), N.ptr(c2), N.ptr(c3), N.ptr(lab_host),
                                            float(cfg_scale), st))
            N.check(lib.drs_sampler_begin(plan, N.ptr(xs), N.ptr(zbuf), N.ptr(eps), first, st))
            for i in range(first, last - 1, -1):
                if i > 1:
                    if noise is None:
                        zbuf.normal_(generator=generator)  # same generator call sequence as torch.randn_like(x)
                    else:
                        zbuf.copy_(noise(i).to(dev, torch.float32))
                N.check(lib.drs_sampler_step(plan, 1 if use_graph else 0, st))
                if frames is not None:
                    frames.append(xs.clone())
            N.check(lib.drs_plan_check(plan, st))
        x.copy_(xs)
        return x

    def _finish(self, model, frames, generate_video):
        if generate_video:
            self.video_maker(frames, os.path.join(os.getcwd(), "models_run", self.model_name, "results",
                                                  "video_denoising.mp4"), 100)
        model.train()  # the reference leaves the model in train mode (train_diffusion_superres.py:254)

    @staticmethod
    def video_maker(frames, video_path, fps):
        """Writes the per-step frames with OpenCV (utils.py:344-390 is the reference's plotting helper; only the
        call contract -- frames, path, fps -- is kept)."""
        import cv2
        import numpy as np
        os.makedirs(os.path.dirname(video_path), exist_ok=True)
        first = frames[0]
        h, w = first.shape[-2:]
        writer = cv2.VideoWriter(video_path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (w, h))
        for f in frames:
            img = f[0].detach().clamp(0, 1).mul(255).byte().permute(1, 2, 0).cpu().numpy()
            if img.shape[2] == 1:
                img = np.repeat(img, 3, axis=2)
            writer.write(np.ascontiguousarray(img[:, :, ::-1]))
        writer.release()

    def _start_state(self, n, channels, x_T, dev, cpu_generator=None, generator=None):
        if x_T is None and generator is not None and cpu_generator is None:
            # batched aggregation sampling: drawn on the device (a host randn of a 31-patch batch costs ~40 ms)
            return torch.randn((n, channels, self.image_size, self.image_size), device=dev, generator=generator)
        if x_T is None:
            # CPU default generator, like the reference
            x = torch.randn((n, channels, self.image_size, self.image_size), generator=cpu_generator)
        else:
            x = x_T
            if tuple(x.shape) != (n, channels, self.image_size, self.image_size):
                raise ValueError("x_T must be [n, C, image_size, image_size]")
        return x.to(dev, torch.float32).clone().contiguous()


class Diffusion(_DiffusionBase):
    """Super-resolution diffusion process (train_diffusion_superres.py:78)."""

    def __init__(self, noise_schedule: str, model: nn.Module, snapshot_path: str, noise_steps=1000, beta_start=1e-4,
                 beta_end=0.02, device="cuda", magnification_factor=4, image_size=224, model_name="superres",
                 Degradation_type="BSRGAN", multiple_gpus=False, ema_smoothing=False):
        self.magnification_factor = magnification_factor
        self.Degradation_type = Degradation_type
        self._init_common(noise_schedule, model, snapshot_path, noise_steps, beta_start, beta_end, device, image_size,
                          model_name, multiple_gpus, ema_smoothing)

    def sample(self, n, model, lr_img, input_channels=3, generate_video=False, *, x_T=None, noise=None,
               use_graph=True, n_steps=None):
        """n stochastic super-resolutions of ONE low-resolution image [C, h, w] -> [n, C, S, S] fp32, unclamped.
        n_steps (keyword extension, tests / benchmarks): stop after the first n_steps reverse steps of the chain."""
        dev = self._native_device()
        lr = lr_img.to(dev).unsqueeze(0)
        return self._sample_conditioned(n, model, lr, input_channels, generate_video, x_T, noise, use_graph,
                                        n_steps=n_steps)

    def sample_batched(self, model, lr_imgs, input_channels=3, *, x_T=None, noise=None, use_graph=True,
                       generator=None, cpu_generator=None, n_steps=None):
        """One super-resolution per low-resolution image of a batch [n, C, h, w] (used by aggregation sampling; the
        reference can only express this as n sequential sample(1, ...) calls, Aggregation_Sampling.py:94-95).
        generator / cpu_generator: optional private torch.Generators for the per-step device noise and for x_T
        (default: the global ones, like the reference)."""
        dev = self._native_device()
        return self._sample_conditioned(lr_imgs.shape[0], model, lr_imgs.to(dev), input_channels, False, x_T, noise,
                                        use_graph, generator, cpu_generator, n_steps)

    def _sample_conditioned(self, n, model, lr, input_channels, generate_video, x_T, noise, use_graph,
                            generator=None, cpu_generator=None, n_steps=None):
        dev = self._native_device()
        frames = [] if generate_video else None
        model.eval()
        with torch.no_grad():
            if self.Degradation_type.lower() not in ("downblur", "bsrgan", "downblurnoise"):
                raise ValueError("The degradation type must be either BSRGAN or DownBlur")
            x = self._start_state(n, input_channels, x_T, dev, cpu_generator, generator)
            mag = int(self.magnification_factor)
            if lr.shape[-1] * mag != self.image_size or lr.shape[-2] * mag != self.image_size:
                raise ValueError("lr_img size %s times magnification %d must equal image_size %d"
                                 % (tuple(lr.shape[-2:]), mag, self.image_size))
            lr = lr.to(torch.float32).contiguous()
            x = self._reverse_process(model, x, lr, mag, None, False, 0.0, noise, frames, use_graph,
                                      n_steps=n_steps, generator=generator)
        self._finish(model, frames, generate_video)
        return x


class Diffusion_SAR_TO_NDVI(_DiffusionBase):
    """SAR -> NDVI diffusion process (train_diffusion_SAR_TO_NDVI.py:79)."""

    def __init__(self, noise_schedule: str, model: nn.Module, snapshot_path: str, noise_steps=1000, beta_start=1e-4,
                 beta_end=0.02, device="cuda", image_size=224, model_name="SAR_TO_NDVI", multiple_gpus=False,
                 ema_smoothing=False):
        self._init_common(noise_schedule, model, snapshot_path, noise_steps, beta_start, beta_end, device, image_size,
                          model_name, multiple_gpus, ema_smoothing)

    def sample(self, n, model, SAR_img, NDVI_channels=1, generate_video=False, *, x_T=None, noise=None,
               use_graph=True, n_steps=None):
        dev = self._native_device()
        sar = SAR_img.to(dev).unsqueeze(0).to(torch.float32).contiguous()
        frames = [] if generate_video else None
        model.eval()
        with torch.no_grad():
            x = self._start_state(n, NDVI_channels, x_T, dev)
            if sar.shape[-1] != self.image_size or sar.shape[-2] != self.image_size:
                raise ValueError("SAR_img must be [C, image_size, image_size]")
            x = self._reverse_process(model, x, sar, 1, None, False, 0.0, noise, frames, use_graph, n_steps=n_steps)
        self._finish(model, frames, generate_video)
        return x


class Diffusion_generation(_DiffusionBase):
    """Class-conditional generation with classifier-free guidance
    (generate_new_imgs/train_diffusion_generation.py:81)."""

    def __init__(self, noise_schedule: str, model: nn.Module, snapshot_path: str, noise_steps=1000, beta_start=1e-4,
                 beta_end=0.02, device="cuda", image_size=224, model_name="generation", multiple_gpus=False,
                 ema_smoothing=False):
        self._init_common(noise_schedule, model, snapshot_path, noise_steps, beta_start, beta_end, device, image_size,
                          model_name, multiple_gpus, ema_smoothing)

    def sample(self, n, model, target_class=None, cfg_scale=3, input_channels=3, generate_video=False, *, x_T=None,
               noise=None, use_graph=True, n_steps=None):
        dev = self._native_device()
        frames = [] if generate_video else None
        model.eval()
        with torch.no_grad():
            x = self._start_state(n, input_channels, x_T, dev)
            # With target_class None the reference evaluates model(x, t, None) twice and lerps a tensor with
            # itself, which returns it unchanged: one unconditional pass is bit-equivalent.
            cfg = target_class is not None and cfg_scale > 0
            x = self._reverse_process(model, x, None, 1, target_class, cfg, float(cfg_scale), noise, frames, use_graph,
                                      n_steps=n_steps)
        self._finish(model, frames, generate_video)
        return x
