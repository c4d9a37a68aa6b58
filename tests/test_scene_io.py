"""CPU tests of the file boundary of aggregation sampling: PNG open, Image.BICUBIC
square-resize to the nearest size of the fixed list, ToTensor, ToPILImage().save -- against the reference's own CLI
body (Aggregation_Sampling.py:140-205) run from the imported reference with a recording stand-in for the sampler."""
import os
import sys
import types

import numpy as np
import pytest
import torch

import common as T
import diffusionremotesensing_b200 as D
from oracle import reference_loader as RL

needs_reference = pytest.mark.skipif(not RL.available(), reason="reference checkout not present on this host")


def write_png(path, w, h, seed):
    from PIL import Image
    rng = np.random.Generator(np.random.PCG64(seed))
    Image.fromarray(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)).save(path)


def test_nearest_scene_size_rule():
    # Aggregation_Sampling.py:176-185: |s - w| + |s - h| minimal, first of equals wins
    assert D.nearest_scene_size(70, 60) == 64
    assert D.nearest_scene_size(96, 97) == 128         # |128-96| + |128-97| = 63 < |64-96| + |64-97| = 65
    assert D.nearest_scene_size(96, 96) == 64          # 64 vs 64: a tie, np.argmin keeps the first
    assert D.nearest_scene_size(2000, 2100) == 2048
    assert D.nearest_scene_size(9500, 9600) == 10000


def test_load_scene_square_and_resized(tmp_path):
    from PIL import Image
    write_png(tmp_path / "sq.png", 96, 96, 1)
    a = D.load_scene(str(tmp_path / "sq.png"))
    assert a.shape == (1, 3, 96, 96) and a.dtype == torch.float32 and 0.0 <= a.min() and a.max() <= 1.0
    raw = np.asarray(Image.open(tmp_path / "sq.png"), dtype=np.float32) / 255.0
    assert torch.equal(a[0], torch.from_numpy(raw).permute(2, 0, 1))
    write_png(tmp_path / "ns.png", 70, 60, 2)
    b = D.load_scene(str(tmp_path / "ns.png"))
    assert b.shape == (1, 3, 64, 64)
    # PIL's BICUBIC (a = -0.5 on uint8) is not torch's bicubic (a = -0.75 on float): the two must differ
    t = torch.from_numpy(np.asarray(Image.open(tmp_path / "ns.png"), dtype=np.float32) / 255.0).permute(2, 0, 1)
    f = torch.nn.functional.interpolate(t.unsqueeze(0), size=(64, 64), mode="bicubic", align_corners=False)
    assert (b - f.clamp(0, 1)).abs().max() > 1e-3
    # a tensor scene goes through the same filter
    c = D.prepare_scene(t)
    assert torch.equal(b, c)


def test_save_scene_roundtrip(tmp_path):
    from PIL import Image
    img = T.np_rand(5, 1, 3, 40, 40)
    D.save_scene(img, str(tmp_path / "out.png"))
    back = np.asarray(Image.open(tmp_path / "out.png"))
    assert back.shape == (40, 40, 3)
    assert np.array_equal(back, (img[0] * 255).byte().permute(1, 2, 0).numpy())   # ToPILImage: mul(255).byte()


@needs_reference
@pytest.mark.parametrize("w,h", [(96, 96), (70, 60), (150, 100)])
def test_file_boundary_matches_reference_launch(tmp_path, w, h):
    RL.load("superres")
    RL.load_aggregation()
    ref_mod = sys.modules["Aggregation_Sampling"]
    captured = {}
    fake_out = T.np_rand(9, 1, 3, 128, 128)

    class Recorder:
        def __init__(self, img_lr, patch_size, stride, magnification_factor, diffusion_model, device):
            captured["img_lr"] = img_lr.clone()
            captured["args"] = (patch_size, stride, magnification_factor)

        def aggregation_sampling(self):
            return fake_out.clone()

    src = tmp_path / "in.png"
    write_png(src, w, h, 3)
    args = types.SimpleNamespace(
        snapshot_folder_path=str(tmp_path), snapshot_name="missing.pt", magnification_factor=2, inp_out_channels=3,
        noise_schedule="cosine", device="cpu", model_input_size=64, noise_steps=5, model_name="t",
        Degradation_type="DownBlur", patch_size=32, stride=16, destination_path=str(tmp_path / "ref_out.png"),
        img_lr_path=str(src), UNet_type="Residual Attention UNet")
    saved_path = list(sys.path)
    real = ref_mod.split_aggregation_sampling
    ref_mod.split_aggregation_sampling = Recorder
    try:
        sys.path.insert(0, RL.REFERENCE_ROOT)
        ref_mod.launch(args)
    finally:
        ref_mod.split_aggregation_sampling = real
        sys.path[:] = saved_path
    ours = D.load_scene(str(src))
    assert torch.equal(ours, captured["img_lr"]), "scene preparation differs from Aggregation_Sampling.py:170-191"
    D.save_scene(fake_out, str(tmp_path / "our_out.png"))
    from PIL import Image
    assert np.array_equal(np.asarray(Image.open(tmp_path / "our_out.png")),
                          np.asarray(Image.open(tmp_path / "ref_out.png")))
