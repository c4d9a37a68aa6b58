"""ORACLE -- test infrastructure only. Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, build container only) on seeded synthetic inputs. Re-run with:  python oracle/make_golden.py

Every fixture stores outputs of the reference's own classes:
  schedules.npz        Diffusion(...).alpha / alpha_hat / beta           (train_diffusion_superres.py:116-126)
  forward_<family>.npz Residual_Attention_UNet_*.forward in eval mode     (UNet_model_*.py forward)
  sample_<family>.npz  Diffusion.sample() with injected noise             (train_diffusion_*.py sample)
  aggregation.npz      split_aggregation_sampling: patch grids, Gaussian weights, blend with a stub sampler
                                                                            (Aggregation_Sampling.py:30-138)
Inputs and weights are NOT stored: they are regenerated from numpy PCG64 seeds by tests/common.py.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import reference_loader as RL  # noqa: E402
import common as T  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

FORWARD_CASES = {
    # family: (S, n, timesteps, cond shape, magnification, labels)
    "superres": dict(S=64, n=2, t=[37, 5], cond=(1, 3, 32, 32), mag=2, y=None),
    "sar": dict(S=32, n=2, t=[49, 1], cond=(1, 2, 32, 32), mag=1, y=None),
    "generation": dict(S=32, n=3, t=[11, 30, 2], cond=None, mag=1, y=[3, 7, 0]),
}
SAMPLE_CASES = {
    "superres": dict(S=64, n=2, steps=12, schedule="cosine"),
    "sar": dict(S=32, n=2, steps=10, schedule="linear"),
    "generation": dict(S=32, n=2, steps=8, schedule="cosine", y=[4], cfg=3),
}
GRID_CASES = [(2048, 2048, 128, 64, 2), (2048, 2048, 128, 96, 2), (1024, 1024, 128, 64, 2), (512, 512, 64, 32, 2),
              (300, 300, 128, 64, 2), (256, 256, 128, 100, 2), (200, 328, 64, 48, 4), (128, 128, 128, 128, 2),
              (129, 131, 128, 17, 1)]


def ref_model(family, seed):
    unet_cls, diff_cls = RL.load(family)
    args = {"superres": (3, 3, "cpu"), "sar": (2, 1, "cpu"), "generation": (3, 3, 10, "cpu")}[family]
    ours = T.build_model(family)
    sd = T.synthetic_state_dict(ours, seed)
    m = unet_cls(*args)
    m.load_state_dict(sd)
    m.eval()
    return m, diff_cls, sd


def forward_inputs(family, case, seed=100):
    x_ch = {"superres": 3, "sar": 1, "generation": 3}[family]
    x = T.np_randn(seed + 1, case["n"], x_ch, case["S"], case["S"])
    t = torch.tensor(case["t"], dtype=torch.long)
    cond = T.np_rand(seed + 2, *case["cond"]) if case["cond"] else None
    y = torch.tensor(case["y"], dtype=torch.long) if case["y"] else None
    return x, t, cond, y


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)

    # ---- schedules -------------------------------------------------------------------------------------------
    _, diff_cls = RL.load("superres")
    dummy = torch.nn.Linear(1, 1)
    sched = {}
    for kind, T_ in (("cosine", 50), ("cosine", 1500), ("linear", 6), ("linear", 1000)):
        d = diff_cls(kind, dummy, "/nonexistent", noise_steps=T_, device="cpu", magnification_factor=2, image_size=64,
                     Degradation_type="DownBlur")
        for name in ("alpha", "alpha_hat", "beta"):
            sched[f"{kind}{T_}_{name}"] = getattr(d, name).numpy()
    np.savez_compressed(os.path.join(OUT, "schedules.npz"), **sched)

    # ---- state_dict layout of the reference modules (keys, shapes, in registration order) -----------------------
    import json
    layout = {}
    for fam in FORWARD_CASES:
        unet_cls, _ = RL.load(fam)
        args = {"superres": (3, 3, "cpu"), "sar": (2, 1, "cpu"), "generation": (3, 3, 10, "cpu")}[fam]
        layout[fam] = [[k, list(v.shape)] for k, v in unet_cls(*args).state_dict().items()]
    with open(os.path.join(OUT, "state_dict_layout.json"), "w") as f:
        json.dump(layout, f)

    # ---- one UNet call per family ----------------------------------------------------------------------------
    for fam, case in FORWARD_CASES.items():
        m, _, _ = ref_model(fam, seed=7)
        x, t, cond, y = forward_inputs(fam, case)
        if fam == "superres":
            eps = m(x, t, cond, case["mag"])
        elif fam == "sar":
            eps = m(x, t, cond)
        else:
            eps = m(x, t, y)
            eps_uncond = m(x, t, None)
        out = {"eps": eps.numpy()}
        if fam == "generation":
            out["eps_uncond"] = eps_uncond.numpy()
        # intermediate activations through forward hooks on the reference modules
        taps = {}
        hooks = [
            m.conv_blocks[0].register_forward_hook(lambda mod, i, o: taps.__setitem__("b0.out", o)),
            m.bottle_neck.register_forward_hook(lambda mod, i, o: taps.__setitem__("bn.out", o)),
            m.attention_blocks[0].register_forward_hook(lambda mod, i, o: taps.__setitem__("att0", o)),
            m.ups[2].register_forward_hook(lambda mod, i, o: taps.__setitem__("ut2", o)),
        ]
        if fam == "superres":
            m(x, t, cond, case["mag"])
        elif fam == "sar":
            m(x, t, cond)
        else:
            m(x, t, y)
        for h in hooks:
            h.remove()
        for k, v in taps.items():
            out["tap_" + k + "_absmax"] = np.float64(v.abs().max().item())
            out["tap_" + k + "_sum"] = np.float64(v.double().sum().item())
            out["tap_" + k + "_corner"] = v[0, :8, :4, :4].numpy()
        np.savez_compressed(os.path.join(OUT, f"forward_{fam}.npz"), **out)
        print(fam, "forward eps absmax", float(eps.abs().max()))

    # ---- short trajectories through the reference Diffusion.sample() with injected noise ---------------------
    for fam, case in SAMPLE_CASES.items():
        m, diff_cls, _ = ref_model(fam, seed=7)
        S, n, steps = case["S"], case["n"], case["steps"]
        x_ch = {"superres": 3, "sar": 1, "generation": 3}[fam]
        x_T = T.np_randn(200, n, x_ch, S, S)
        noise_fn = lambda i: T.np_randn(1000 + i, n, x_ch, S, S)  # noqa: E731
        kw = dict(noise_steps=steps, device="cpu", image_size=S)
        if fam == "superres":
            d = diff_cls(case["schedule"], m, "/nonexistent", magnification_factor=2, Degradation_type="DownBlur", **kw)
        else:
            d = diff_cls(case["schedule"], m, "/nonexistent", **kw)
        with RL.injected_noise(x_T, noise_fn) as arm:
            arm(steps)
            if fam == "superres":
                x0 = d.sample(n, m, T.np_rand(201, 3, S // 2, S // 2), input_channels=3)
            elif fam == "sar":
                x0 = d.sample(n, m, T.np_rand(201, 2, S, S), NDVI_channels=1)
            else:
                x0 = d.sample(n, m, target_class=torch.tensor(case["y"]), cfg_scale=case["cfg"], input_channels=3)
        np.savez_compressed(os.path.join(OUT, f"sample_{fam}.npz"), x0=x0.numpy())
        print(fam, "sample absmax", float(x0.abs().max()))

    # ---- aggregation sampling -----------------------------------------------------------------------------------
    agg_cls = RL.load_aggregation()

    class StubDiffusion:
        """sample() = nearest x k of the patch plus a patch-dependent ramp (deterministic stand-in for the UNet)."""
        model = None

        def __init__(self, k):
            self.k = k
            self.calls = 0

        def sample(self, n, model, lr, input_channels=3, generate_video=False):
            up = torch.nn.functional.interpolate(lr.unsqueeze(0), scale_factor=self.k, mode="nearest")
            self.calls += 1
            return up * (1.0 + 0.01 * self.calls) - 0.05

    agg = {}
    for (H, W, P, s, k) in GRID_CASES:
        a = agg_cls(torch.zeros(1, 3, H, W), P, s, k, StubDiffusion(k), "cpu")
        agg[f"grid_{H}_{W}_{P}_{s}_{k}"] = np.asarray(a.patches_sr_infos, dtype=np.int32)
    a = agg_cls(torch.zeros(1, 3, 64, 64), 32, 16, 2, StubDiffusion(2), "cpu")
    agg["weight_64"] = a.weight[0, 0].numpy()
    a = agg_cls(torch.zeros(1, 3, 256, 256), 128, 64, 2, StubDiffusion(2), "cpu")
    w256 = a.weight[0, 0].numpy()
    agg["weight_256_sha256"] = np.frombuffer(bytes.fromhex(sha(w256)), dtype=np.uint8)
    agg["weight_256_probe"] = np.asarray([w256[0, 0], w256[128, 127], w256[128, 128], w256[127, 128]], dtype=np.float32)
    # blend of a small scene: LR 80 x 104, patch 32, stride 24, k = 2
    img = T.np_rand(300, 1, 3, 80, 104)
    a = agg_cls(img, 32, 24, 2, StubDiffusion(2), "cpu")
    agg["blend_small"] = a.aggregation_sampling().numpy()
    np.savez_compressed(os.path.join(OUT, "aggregation.npz"), **agg)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  ", f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
