"""GPU tests of the one-call entry points against the oracle."""
import pytest
import torch

import common as T
import diffusionremotesensing_b200 as D
from oracle import restatement as R

pytestmark = pytest.mark.gpu


def test_super_resolver_matches_oracle(cuda_device, tmp_path):
    name = "Residual_Attention_UNet_superres_magnification2_LRimgsize32_test_downblur"
    m, sd = T.default_init_model("superres", seed=4)
    (tmp_path / name / "weights").mkdir(parents=True)
    torch.save({"MODEL_STATE": sd, "EPOCHS_RUN": 3}, tmp_path / name / "weights" / "snapshot.pt")
    lr = T.np_rand(5, 3, 32, 32)
    steps = 8
    x_T = T.np_randn(6, 1, 3, 64, 64)
    noise = lambda i: T.np_randn(600 + i, 1, 3, 64, 64)  # noqa: E731
    got = D.super_resolver(lr, str(cuda_device), name, noise_steps=steps, snapshot_root=str(tmp_path), x_T=x_T,
                           noise=noise)
    with torch.no_grad():
        ref = R.sample(sd, "superres", R.noise_schedule("cosine", steps), steps, x_T, noise, cond=lr.unsqueeze(0),
                       magnification=2)
    assert got.shape == (1, 3, 64, 64) and 0.0 <= got.min().item() and got.max().item() <= 1.0
    assert T.psnr_ref_range(got, ref.clamp(0, 1)) >= 40.0
    assert D.parse_model_name(name) == (2, 64)
    with pytest.raises(IndexError):
        D.parse_model_name("no_fields_here")


def test_sar_generator_normalises_and_samples(cuda_device, tmp_path):
    name = "Residual_Attention_UNet_EMA_imgsize128_SAR_TO_NDVI"   # the one model name the reference function knows
    m, sd = T.default_init_model("sar", seed=6)
    (tmp_path / name / "weights").mkdir(parents=True)
    torch.save({"MODEL_STATE": sd, "EPOCHS_RUN": 1}, tmp_path / name / "weights" / "snapshot.pt")
    sar = T.np_rand(8, 2, 128, 128) * 1.6 - 0.8          # in (-1, 1) with negatives -> (x + 1) / 2
    x_T = T.np_randn(9, 2, 1, 128, 128)
    noise = lambda i: T.np_randn(700 + i, 2, 1, 128, 128)  # noqa: E731
    got = D.SAR_to_NDVI_generator(sar, str(cuda_device), n_generations=2, noise_steps=5, snapshot_root=str(tmp_path),
                                  x_T=x_T, noise=noise)
    with torch.no_grad():
        ref = R.sample(sd, "sar", R.noise_schedule("cosine", 5), 5, x_T, noise, cond=((sar + 1) / 2).unsqueeze(0))
    assert got.shape == (2, 1, 128, 128)
    assert T.psnr_ref_range(got, ref) >= 40.0
    with pytest.raises(ValueError):
        D.SAR_to_NDVI_generator(sar * 3, str(cuda_device), noise_steps=5)


def test_noise_images_bit_exact(cuda_device):
    d = D.Diffusion("cosine", torch.nn.Linear(1, 1), "/nonexistent", noise_steps=1500, device=str(cuda_device))
    x = T.np_rand(11, 4, 3, 32, 32).to(cuda_device)
    t = torch.tensor([1, 700, 1499, 42], device=cuda_device)
    torch.manual_seed(5)
    x_t, eps = d.noise_images(x, t)
    ah = d.alpha_hat[t]
    want = torch.sqrt(ah)[:, None, None, None] * x + torch.sqrt(1 - ah)[:, None, None, None] * eps
    assert torch.equal(x_t, want)
    ts = d.sample_timesteps(64)
    assert ts.min() >= 1 and ts.max() < 1500


def test_prepare_scene_and_generate_per_class(cuda_device):
    scene = D.prepare_scene(T.np_rand(12, 3, 60, 70))
    assert scene.shape == (1, 3, 64, 64)
    assert D.prepare_scene(T.np_rand(12, 3, 96, 96)).shape == (1, 3, 96, 96)
    m, sd = T.default_init_model("generation", seed=2)
    m.to(cuda_device)
    d = D.Diffusion_generation("linear", m, "/nonexistent", noise_steps=4, device=str(cuda_device), image_size=32)
    out = D.generate_per_class(m, d, 10)     # default RNG path; parity: test_gpu_baseline_configs.py
    assert out.shape == (10, 3, 32, 32) and 0.0 <= out.min().item() and out.max().item() <= 1.0


def test_model_copies_own_their_native_state(cuda_device):
    # ema = copy.deepcopy(model) after a forward must not share (and later double-free) the packed model / plans
    import copy
    m, _ = T.default_init_model("sar", seed=3)
    m.to(cuda_device).eval()
    x = T.np_randn(13, 2, 1, 32, 32).to(cuda_device)
    t = torch.full((2,), 5, device=cuda_device)
    sar = T.np_rand(14, 1, 2, 32, 32).to(cuda_device)
    with torch.no_grad():
        a = m(x, t, sar).clone()
        ema = copy.deepcopy(m)
        assert ema._handles() is not m._handles() and ema._handles().model is None
        b = ema(x, t, sar).clone()
        c = m(x, t, sar).clone()
    assert torch.equal(a, b) and torch.equal(a, c)
    del ema
    import gc
    gc.collect()
    with torch.no_grad():
        assert torch.equal(m(x, t, sar), a)


def test_forward_rejects_out_of_range_labels(cuda_device):
    m, _ = T.default_init_model("generation", seed=2)
    m.to(cuda_device).eval()
    x = T.np_randn(15, 2, 3, 32, 32).to(cuda_device)
    t = torch.full((2,), 5, device=cuda_device)
    with pytest.raises(IndexError):
        m(x, t, torch.tensor([3, 10], device=cuda_device))
