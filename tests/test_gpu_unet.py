"""GPU parity of one UNet evaluation through the C ABI against the oracle (CPU fp32) and the golden fixtures.

Tolerance (BASELINE.json north_star): per-step UNet output max|eps - eps_ref| / max|eps_ref| <= 2e-2 for the bf16
tensor-core path against the fp32 reference. Intermediate activations are held to the same bound so that a failure
names the first layer that drifts.
"""
import os

import numpy as np
import pytest
import torch

import common as T
from oracle import make_golden as G
from oracle import restatement as R

pytestmark = pytest.mark.gpu

TOL = 2e-2
LAYERS = ["h0", "b0.h", "b0.out", "d0", "b1.h", "b1.out", "d1", "b2.h", "b2.out", "d2", "bn.h", "bn.out",
          "g0", "psi0", "att0", "uc0", "ut0", "x0", "g1", "psi1", "att1", "uc1", "ut1", "x1",
          "g2", "psi2", "att2", "uc2", "ut2"]


def gpu_model(family, dev, seed=7):
    m = T.build_model(family)
    sd = T.synthetic_state_dict(m, seed)
    m.to(dev).eval()
    return m, sd


def call(m, family, x, t, cond, mag, y):
    if family == "superres":
        return m(x, t, cond, mag)
    if family == "sar":
        return m(x, t, cond)
    return m(x, t, y)


@pytest.mark.parametrize("family", T.FAMILIES)
def test_forward_matches_oracle_layer_by_layer(cuda_device, family):
    case = G.FORWARD_CASES[family]
    m, sd = gpu_model(family, cuda_device)
    x, t, cond, y = G.forward_inputs(family, case)
    taps = {}
    with torch.no_grad():
        ref = R.unet_forward(sd, family, x, t, cond, case["mag"], y, taps)
        got = call(m, family, x.to(cuda_device), t.to(cuda_device), None if cond is None else cond.to(cuda_device),
                   case["mag"], None if y is None else y.to(cuda_device))
    n, S = case["n"], case["S"]
    plan = m.native_plan(n, n, 1, S, case["mag"])
    report = []
    for name in LAYERS:
        want = taps[name]
        have = m.debug_activation(plan, name, tuple(want.shape))
        report.append((name, T.max_rel_err(have, want)))
    report.append(("eps", T.max_rel_err(got, ref)))
    text = ", ".join(f"{k}={v:.2e}" for k, v in report)
    print(f"[{family}] {text}")
    bad = [(k, v) for k, v in report if not v <= TOL]
    assert not bad, f"{family}: layers over tolerance {bad}; all: {text}"
    g = np.load(os.path.join(T.GOLDEN, f"forward_{family}.npz"))
    assert T.max_rel_err(got, torch.from_numpy(g["eps"])) <= TOL


def test_generation_unconditional_and_broadcast_label(cuda_device):
    case = G.FORWARD_CASES["generation"]
    m, sd = gpu_model("generation", cuda_device)
    x, t, _, _ = G.forward_inputs("generation", case)
    g = np.load(os.path.join(T.GOLDEN, "forward_generation.npz"))
    with torch.no_grad():
        got = m(x.to(cuda_device), t.to(cuda_device), None)
        assert T.max_rel_err(got, torch.from_numpy(g["eps_uncond"])) <= TOL
        one = torch.tensor([5])
        ref = R.unet_forward(sd, "generation", x, t, y=one)
        got = m(x.to(cuda_device), t.to(cuda_device), one.to(cuda_device))
        assert T.max_rel_err(got, ref) <= TOL


def test_superres_batched_condition_and_per_sample_timesteps(cuda_device):
    # aggregation sampling feeds a different LR patch per sample
    m, sd = gpu_model("superres", cuda_device, seed=21)
    n = 5
    x = T.np_randn(31, n, 3, 64, 64)
    lr = T.np_rand(32, n, 3, 32, 32)
    t = torch.tensor([1, 10, 100, 700, 1499])
    with torch.no_grad():
        ref = R.unet_forward(sd, "superres", x, t, lr, 2)
        got = m(x.to(cuda_device), t.to(cuda_device), lr.to(cuda_device), 2)
    assert T.max_rel_err(got, ref) <= TOL


@pytest.mark.parametrize("S,mag,n", [(16, 1, 1), (40, 2, 3), (128, 4, 2), (256, 2, 1)])
def test_superres_shapes(cuda_device, S, mag, n):
    # smallest legal size, a ragged size (not a multiple of the 16 x 8 tile), magnification 4, the cfg-2 image size
    m, sd = gpu_model("superres", cuda_device, seed=5)
    x = T.np_randn(41, n, 3, S, S)
    lr = T.np_rand(42, 1, 3, S // mag, S // mag)
    t = torch.full((n,), 123)
    with torch.no_grad():
        ref = R.unet_forward(sd, "superres", x, t, lr, mag)
        got = m(x.to(cuda_device), t.to(cuda_device), lr.to(cuda_device), mag)
    assert T.max_rel_err(got, ref) <= TOL


def test_weight_update_repacks(cuda_device):
    m, sd = gpu_model("sar", cuda_device, seed=3)
    x = T.np_randn(51, 1, 1, 32, 32).to(cuda_device)
    sar = T.np_rand(52, 1, 2, 32, 32).to(cuda_device)
    t = torch.tensor([9], device=cuda_device)
    with torch.no_grad():
        a = m(x, t, sar)
        m.load_state_dict({k: v.to(cuda_device) for k, v in T.synthetic_state_dict(T.build_model("sar"), 4).items()})
        b = m(x, t, sar)
        ref = R.unet_forward(T.synthetic_state_dict(T.build_model("sar"), 4), "sar", x.cpu(), t.cpu(), sar.cpu())
    assert not torch.equal(a, b)
    assert T.max_rel_err(b, ref) <= TOL


def test_errors_are_loud(cuda_device):
    m, _ = gpu_model("superres", cuda_device)
    x = torch.zeros(1, 3, 64, 64, device=cuda_device)
    t = torch.zeros(1, dtype=torch.long, device=cuda_device)
    with pytest.raises(ValueError):
        m(x, t, torch.zeros(1, 3, 16, 16, device=cuda_device), 2)      # lr size * mag != x size
    m.train()
    with pytest.raises(RuntimeError):
        m(x, t, torch.zeros(1, 3, 32, 32, device=cuda_device), 2)      # eval-mode path only
    m.eval()
    with pytest.raises(Exception):
        m(torch.zeros(1, 3, 20, 20, device=cuda_device), t, torch.zeros(1, 3, 10, 10, device=cuda_device), 2)  # S % 8
