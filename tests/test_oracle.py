"""CPU tests that pin the oracle (oracle/restatement.py):

  * against the golden fixtures generated from the unmodified reference (tests/golden, made by oracle/make_golden.py);
  * against the reference itself, imported from /root/reference when that checkout exists (build container only).

Floating-point fixtures are compared bit-exactly first and, failing that (another host CPU can pick other oneDNN
kernels), within 2e-6 of the tensor's magnitude; integer fixtures (patch grids) are always exact.
"""
import json
import os

import numpy as np
import pytest
import torch

import common as T
from oracle import make_golden as G
from oracle import reference_loader as RL
from oracle import restatement as R

needs_reference = pytest.mark.skipif(not RL.available(), reason="reference checkout not present on this host")


def assert_close_or_equal(test: torch.Tensor, ref: torch.Tensor, rel=2e-6):
    if torch.equal(test, ref):
        return
    scale = ref.abs().max().item()
    err = (test - ref).abs().max().item()
    assert err <= rel * max(scale, 1.0), f"max err {err:.3e} vs magnitude {scale:.3e}"


def golden(name):
    return np.load(os.path.join(T.GOLDEN, name))


def oracle_sd(family, seed=7):
    return T.synthetic_state_dict(T.build_model(family), seed)


@pytest.mark.parametrize("kind,steps", [("cosine", 50), ("cosine", 1500), ("linear", 6), ("linear", 1000)])
def test_schedule_matches_reference_tables(kind, steps):
    g = golden("schedules.npz")
    alpha, alpha_hat, beta = R.noise_schedule(kind, steps)
    for name, t in (("alpha", alpha), ("alpha_hat", alpha_hat), ("beta", beta)):
        assert_close_or_equal(t, torch.from_numpy(g[f"{kind}{steps}_{name}"]), rel=1e-7)


def test_schedule_known_answers():
    # SURVEY.md section 4 probes of the reference ctor
    _, ah, b = R.noise_schedule("cosine", 50)
    assert ah[0].item() == 1.0 and abs(ah[1].item() - 0.99825251) < 1e-7 and abs(ah[-1].item() - 9.71188e-4) < 1e-8
    assert b[0].item() == 0.0 and abs(b[1].item() - 1.7474890e-3) < 1e-8 and abs(b[-1].item() - 0.74975830) < 1e-6
    _, ah, b = R.noise_schedule("linear", 6)
    assert abs(b[1].item() - 4.08e-3) < 1e-8 and abs(ah[-1].item() - 0.94106179) < 1e-7


def test_pos_encoding_known_answers():
    e = R.pos_encoding(torch.tensor([[7.0]]))
    assert e.shape == (1, 100)
    assert abs(e[0, 0].item() - 0.65698659) < 1e-6 and abs(e[0, 50].item() - 0.75390226) < 1e-6


@pytest.mark.parametrize("family", T.FAMILIES)
def test_forward_matches_golden(family):
    case = G.FORWARD_CASES[family]
    sd = oracle_sd(family)
    x, t, cond, y = G.forward_inputs(family, case)
    g = golden(f"forward_{family}.npz")
    taps = {}
    with torch.no_grad():
        eps = R.unet_forward(sd, family, x, t, cond, case["mag"], y, taps)
    assert_close_or_equal(eps, torch.from_numpy(g["eps"]))
    if family == "generation":
        with torch.no_grad():
            assert_close_or_equal(R.unet_forward(sd, family, x, t, y=None), torch.from_numpy(g["eps_uncond"]))
    for name in ("b0.out", "bn.out", "att0", "ut2"):
        assert_close_or_equal(taps[name][0, :8, :4, :4], torch.from_numpy(g[f"tap_{name}_corner"]))
        assert abs(taps[name].abs().max().item() - float(g[f"tap_{name}_absmax"])) <= 2e-6 * float(g[f"tap_{name}_absmax"])


@pytest.mark.parametrize("family", T.FAMILIES)
def test_sample_matches_golden(family):
    case = G.SAMPLE_CASES[family]
    sd = oracle_sd(family)
    S, n, steps = case["S"], case["n"], case["steps"]
    x_ch = {"superres": 3, "sar": 1, "generation": 3}[family]
    x_T = T.np_randn(200, n, x_ch, S, S)
    noise_fn = lambda i: T.np_randn(1000 + i, n, x_ch, S, S)  # noqa: E731
    sched = R.noise_schedule(case["schedule"], steps)
    with torch.no_grad():
        if family == "superres":
            x0 = R.sample(sd, family, sched, steps, x_T, noise_fn, cond=T.np_rand(201, 3, S // 2, S // 2).unsqueeze(0),
                          magnification=2)
        elif family == "sar":
            x0 = R.sample(sd, family, sched, steps, x_T, noise_fn, cond=T.np_rand(201, 2, S, S).unsqueeze(0))
        else:
            x0 = R.sample(sd, family, sched, steps, x_T, noise_fn, labels=torch.tensor(case["y"]), cfg_scale=case["cfg"])
    assert_close_or_equal(x0, torch.from_numpy(golden(f"sample_{family}.npz")["x0"]), rel=2e-5)


@pytest.mark.parametrize("H,W,P,s,k", G.GRID_CASES)
def test_patch_grid_exact(H, W, P, s, k):
    want = golden("aggregation.npz")[f"grid_{H}_{W}_{P}_{s}_{k}"]
    got = np.asarray(R.patch_grid(H, W, P, s, k), dtype=np.int32)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_patch_grid_counts():
    # (LR size, patch, stride) -> number of patches
    for (size, P, s), count in {(2048, 128, 64): 961, (2048, 128, 96): 441, (2048, 128, 128): 256,
                                (1024, 128, 64): 225, (512, 64, 32): 225, (300, 128, 64): 16, (256, 128, 100): 9}.items():
        assert len(R.patch_grid(size, size, P, s, 2)) == count
    assert [g[0] for g in R.patch_grid(300, 300, 128, 64, 2)][::4] == [0, 128, 256, 344]


def test_gaussian_weights_bit_exact():
    g = golden("aggregation.npz")
    w64 = torch.tensor(R.gaussian_weights(64, 64)).to(torch.float32).numpy()
    assert np.array_equal(w64.view(np.uint32), g["weight_64"].view(np.uint32))
    w256 = torch.tensor(R.gaussian_weights(256, 256)).to(torch.float32).numpy()
    assert bytes.fromhex(G.sha(w256)) == g["weight_256_sha256"].tobytes()
    assert np.array_equal(np.asarray([w256[0, 0], w256[128, 127], w256[128, 128], w256[127, 128]], np.float32),
                          g["weight_256_probe"])
    assert np.unravel_index(np.argmax(w256), w256.shape) == (128, 127)     # x midpoint (W-1)/2, y midpoint H/2
    assert np.array_equal(w256, w256[:, ::-1]) and not np.array_equal(w256, w256[::-1, :])


def stub_patches(img, infos, k, P):
    """The stub sampler of oracle/make_golden.py, evaluated per patch in patch order."""
    out = []
    for call, (y0, y1, x0, x1) in enumerate(infos, start=1):
        lr = img[:, :, y0 // k:y0 // k + P, x0 // k:x0 // k + P]
        up = torch.nn.functional.interpolate(lr, scale_factor=k, mode="nearest")
        out.append(up * (1.0 + 0.01 * call) - 0.05)
    return out


def test_blend_matches_golden():
    g = golden("aggregation.npz")
    img = T.np_rand(300, 1, 3, 80, 104)
    infos = R.patch_grid(80, 104, 32, 24, 2)
    patches = stub_patches(img, infos, 2, 32)
    w = torch.tile(torch.tensor(R.gaussian_weights(64, 64)).to(torch.float32), (1, 3, 1, 1))
    res = R.blend(patches, infos, w, 160, 208)
    assert torch.equal(res, torch.from_numpy(g["blend_small"]))


def test_state_dict_layout_fixture_is_consistent():
    with open(os.path.join(T.GOLDEN, "state_dict_layout.json")) as f:
        layout = json.load(f)
    assert {fam: len(v) for fam, v in layout.items()} == {"superres": 299, "sar": 299, "generation": 284}


# ---- against the imported reference (build container only) -----------------------------------------------------------
@needs_reference
@pytest.mark.parametrize("family", T.FAMILIES)
def test_forward_bit_equal_to_reference(family):
    m, _, sd = G.ref_model(family, seed=11)
    case = G.FORWARD_CASES[family]
    x, t, cond, y = G.forward_inputs(family, case, seed=500)
    with torch.no_grad():
        if family == "superres":
            ref = m(x, t, cond, case["mag"])
        elif family == "sar":
            ref = m(x, t, cond)
        else:
            ref = m(x, t, y)
        got = R.unet_forward(sd, family, x, t, cond, case["mag"], y)
    assert torch.equal(ref, got)


@needs_reference
def test_batched_condition_equals_per_sample_reference():
    # the oracle for batched aggregation sampling: the reference UNet accepts lr_img[n, ...]
    m, _, sd = G.ref_model("superres", seed=11)
    x = T.np_randn(1, 3, 3, 64, 64)
    lr = T.np_rand(2, 3, 3, 32, 32)
    t = torch.tensor([20, 20, 20])
    with torch.no_grad():
        batched = R.unet_forward(sd, "superres", x, t, lr, 2)
        single = torch.cat([m(x[i:i + 1], t[i:i + 1], lr[i:i + 1], 2) for i in range(3)])
    assert (batched - single).abs().max().item() <= 2e-6 * single.abs().max().item()


@needs_reference
def test_aggregation_matches_reference_class():
    agg_cls = RL.load_aggregation()

    class Stub:
        model = None
        calls = 0

        def sample(self, n, model, lr, input_channels=3, generate_video=False):
            self.calls += 1
            return torch.nn.functional.interpolate(lr.unsqueeze(0), scale_factor=2, mode="nearest") * (1 + 0.01 * self.calls) - 0.05

    img = T.np_rand(301, 1, 3, 96, 72)
    a = agg_cls(img, 32, 20, 2, Stub(), "cpu")
    infos = R.patch_grid(96, 72, 32, 20, 2)
    assert infos == a.patches_sr_infos
    w = torch.tile(torch.tensor(R.gaussian_weights(64, 64)).to(torch.float32), (1, 3, 1, 1))
    assert torch.equal(w, a.weight)
    assert torch.equal(R.blend(stub_patches(img, infos, 2, 32), infos, w, 192, 144), a.aggregation_sampling())
