"""GPU parity of aggregation sampling: patch tiling and blend indexing bit-exact (BASELINE.json north_star), blend
arithmetic bit-exact fp32 given identical patches, end-to-end scene within the PSNR bound."""
import os

import numpy as np
import pytest
import torch

import common as T
import diffusionremotesensing_b200 as D
from oracle import restatement as R
from test_oracle import stub_patches

pytestmark = pytest.mark.gpu


def weight2d(P, dev):
    return torch.tensor(R.gaussian_weights(P, P)).to(torch.float32).to(dev)


@pytest.mark.parametrize("H,W,P,s,k", [(80, 104, 32, 24, 2), (64, 64, 32, 16, 2), (96, 72, 32, 20, 2),
                                       (40, 40, 16, 16, 4), (33, 47, 32, 5, 1)])
def test_blend_bit_exact_vs_oracle(cuda_device, H, W, P, s, k):
    img = T.np_rand(300 + H, 1, 3, H, W)
    infos = R.patch_grid(H, W, P, s, k)
    patches = stub_patches(img, infos, k, P)
    w = weight2d(P * k, "cpu")
    want = R.blend(patches, infos, torch.tile(w, (1, 3, 1, 1)), H * k, W * k)
    got, wsum = D.blend_patches(torch.cat(patches).to(cuda_device), infos, w.to(cuda_device), H * k, W * k)
    torch.cuda.synchronize()
    assert torch.equal(got.cpu(), want)
    assert (wsum > 0).all()


def test_blend_matches_golden_scene(cuda_device):
    g = np.load(os.path.join(T.GOLDEN, "aggregation.npz"))
    img = T.np_rand(300, 1, 3, 80, 104)
    infos = R.patch_grid(80, 104, 32, 24, 2)
    patches = torch.cat(stub_patches(img, infos, 2, 32)).to(cuda_device)
    got, _ = D.blend_patches(patches, infos, weight2d(64, cuda_device), 160, 208)
    assert torch.equal(got.cpu(), torch.from_numpy(g["blend_small"]))


def test_blend_arbitrary_window_list_and_uncovered_pixels(cuda_device):
    # windows that are not a row-major grid take the scatter path (one launch per patch, same summation order)
    P = 16
    infos = [(0, 16, 0, 16), (8, 24, 8, 24), (0, 16, 8, 24), (8, 24, 0, 16)]
    patches = [T.np_randn(400 + i, 1, 3, P, P) for i in range(4)]
    w = weight2d(P, "cpu")
    want = R.blend(patches, infos, torch.tile(w, (1, 3, 1, 1)), 24, 24)
    got, _ = D.blend_patches(torch.cat(patches).to(cuda_device), infos, w.to(cuda_device), 24, 24)
    assert torch.equal(got.cpu(), want)
    with pytest.raises(Exception):      # the reference asserts pixel_count != 0 (Aggregation_Sampling.py:108)
        D.blend_patches(torch.cat(patches[:1]).to(cuda_device), infos[:1], w.to(cuda_device), 24, 24)


def test_patchifier_and_weights_bit_exact(cuda_device):
    g = np.load(os.path.join(T.GOLDEN, "aggregation.npz"))

    class NoDiffusion:
        model = None

    a = D.split_aggregation_sampling(torch.zeros(1, 3, 300, 300), 128, 64, 2, NoDiffusion(), str(cuda_device))
    assert np.array_equal(np.asarray(a.patches_sr_infos, np.int32), g["grid_300_300_128_64_2"])
    assert len(a.patches_lr) == 16 and a.patches_lr[5].shape == (1, 3, 128, 128)
    a = D.split_aggregation_sampling(torch.zeros(1, 3, 64, 64), 32, 16, 2, NoDiffusion(), str(cuda_device))
    assert a.weight.shape == (1, 3, 64, 64)
    assert np.array_equal(a.weight[0, 0].cpu().numpy().view(np.uint32), g["weight_64"].view(np.uint32))


def test_aggregation_end_to_end(cuda_device):
    # LR 48 x 48 scene, patches 32 / stride 16 (9 patches), x2, 10 linear steps; oracle = per-patch restatement chains
    steps, P, s, k = 10, 32, 16, 2
    m = T.build_model("superres")
    sd = T.synthetic_state_dict(m, 13)
    m.to(cuda_device)
    d = D.Diffusion("linear", m, "/nonexistent", noise_steps=steps, device=str(cuda_device), magnification_factor=k,
                    image_size=P * k, Degradation_type="DownBlur")
    img = T.np_rand(90, 1, 3, 48, 48)
    agg = D.split_aggregation_sampling(img.to(cuda_device), P, s, k, d, str(cuda_device), patch_batch=4)
    x_T = lambda p: T.np_randn(7000 + p, 1, 3, P * k, P * k)                 # noqa: E731
    noise = lambda p, i: T.np_randn(8000 + 100 * p + i, 1, 3, P * k, P * k)  # noqa: E731
    got = agg.aggregation_sampling(noise=noise, x_T=x_T)
    infos = R.patch_grid(48, 48, P, s, k)
    assert infos == agg.patches_sr_infos
    sched = R.noise_schedule("linear", steps)
    ref_patches = []
    with torch.no_grad():
        for p, (y0, y1, x0, x1) in enumerate(infos):
            lr = img[:, :, y0 // k:y0 // k + P, x0 // k:x0 // k + P]
            ref_patches.append(R.sample(sd, "superres", sched, steps, x_T(p), lambda i, p=p: noise(p, i), cond=lr,
                                        magnification=k))
    w = torch.tile(torch.tensor(R.gaussian_weights(P * k, P * k)).to(torch.float32), (1, 3, 1, 1))
    want = R.blend(ref_patches, infos, w, 96, 96)
    assert got.shape == want.shape
    # the scene is clamped to [0, 1]; compare against the unclamped reference range of the patches too
    psnr = T.psnr_ref_range(got, want)
    print(f"[aggregation] PSNR {psnr:.1f} dB")
    assert psnr >= 40.0
