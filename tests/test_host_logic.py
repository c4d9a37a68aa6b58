"""CPU tests of the host side: state_dict layout, schedules, patch grid, Gaussian weights, the C ABI's exported
symbols, loud failure without a GPU, and the world_size-2 sharding logic on the gloo backend."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import common as T
import diffusionremotesensing_b200 as D
from diffusionremotesensing_b200 import _native as N


@pytest.mark.parametrize("family", T.FAMILIES)
def test_state_dict_layout_equals_reference(family):
    with open(os.path.join(T.GOLDEN, "state_dict_layout.json")) as f:
        want = json.load(f)[family]
    got = [[k, list(v.shape)] for k, v in T.build_model(family).state_dict().items()]
    assert got == want  # same keys, same shapes, same registration order


def test_parameter_counts():
    # 4 383 058 / 4 382 238 / 4 383 022 parameters
    counts = [sum(p.numel() for p in T.build_model(f).parameters()) for f in T.FAMILIES]
    assert counts == [4383058, 4382238, 4383022]


def test_snapshot_roundtrip(tmp_path):
    m = T.build_model("sar")
    sd = T.synthetic_state_dict(m, 1)
    path = str(tmp_path / "snapshot.pt")
    # DDP-style "module." prefixes are stripped on load (train_diffusion_superres.py:300)
    torch.save({"MODEL_STATE": {"module." + k: v for k, v in sd.items()}, "EPOCHS_RUN": 17}, path)
    fresh = T.build_model("sar")
    d = D.Diffusion_SAR_TO_NDVI("linear", fresh, path, noise_steps=10, device="cpu", image_size=32)
    assert d.epochs_run == 17
    assert all(torch.equal(v, fresh.state_dict()[k]) for k, v in sd.items())


@pytest.mark.parametrize("kind,steps", [("cosine", 50), ("cosine", 1500), ("linear", 6), ("linear", 1000)])
def test_diffusion_tables_bit_equal_to_reference(kind, steps):
    g = np.load(os.path.join(T.GOLDEN, "schedules.npz"))
    d = D.Diffusion(kind, torch.nn.Linear(1, 1), "/nonexistent", noise_steps=steps, device="cpu")
    for name in ("alpha", "alpha_hat", "beta"):
        assert np.array_equal(getattr(d, name).numpy().view(np.uint32), g[f"{kind}{steps}_{name}"].view(np.uint32))
    assert d.noise_steps == steps and d.magnification_factor == 4 and d.image_size == 224


def test_sampling_without_gpu_fails_loudly():
    m = T.build_model("superres")
    d = D.Diffusion("cosine", m, "/nonexistent", noise_steps=5, device="cpu", magnification_factor=2, image_size=64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d.sample(1, m, torch.zeros(3, 32, 32))
    m.eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 64, 64), torch.zeros(1, dtype=torch.long), torch.zeros(1, 3, 32, 32), 2)


def test_product_code_does_not_import_the_oracle():
    pkg = os.path.join(T.ROOT, "diffusionremotesensing_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
                assert "/root/reference" not in text, f


def test_c_abi_exports_every_declared_symbol():
    # the drop-in boundary (drs_b200.h) and the test / instrumentation hooks (part 2 of drs_b200_diag.h) live in
    # libdrs_b200.so; the micro-benchmarks (part 1 of drs_b200_diag.h) in their own libdrs_b200_diag.so
    find = lambda text: set(re.findall(r"\b(drs_[a-z0-9_]+)\s*\(", text))  # noqa: E731
    header = open(os.path.join(T.ROOT, "include", "drs_b200.h")).read()
    diag = open(os.path.join(T.ROOT, "include", "drs_b200_diag.h")).read()
    diag_own, diag_hooks = diag.split("hooks inside libdrs_b200.so")
    boundary, hooks, micro = find(header), find(diag_hooks), find(diag_own)
    assert boundary and hooks and micro, "no declarations found"
    assert not any(n.startswith("drs_debug_") for n in boundary), "debug entry points leaked into the boundary header"
    assert boundary | hooks == set(N.SIGNATURES), ((boundary | hooks) ^ set(N.SIGNATURES))
    assert micro == set(N.DIAG_SIGNATURES), (micro ^ set(N.DIAG_SIGNATURES))
    lib = ctypes.CDLL(N.LIB_PATH)
    for name in boundary | hooks:
        assert hasattr(lib, name), name
    for name in micro:
        assert not hasattr(lib, name), name + " must not be linked into the product library"
    dl = ctypes.CDLL(N.DIAG_LIB_PATH)
    for name in micro:
        assert hasattr(dl, name), name
    assert lib.drs_version() >= 100


def test_c_abi_rejects_bad_arguments_without_gpu():
    lib = N.lib()
    out = ctypes.c_void_p()
    assert lib.drs_model_create(None, None, 0, 0, ctypes.byref(out)) == N.DRS_E_INVALID
    assert b"null" in lib.drs_last_error()
    assert lib.drs_unet_forward(None, None, None, None) == N.DRS_E_INVALID
    assert lib.drs_ddpm_update(None, None, None, 1.0, 0.0, 0.0, 4, None) == N.DRS_E_INVALID


def test_patchifier_and_weights_match_golden():
    g = np.load(os.path.join(T.GOLDEN, "aggregation.npz"))

    class NoDiffusion:
        model = None

    for key in g.files:
        if key.startswith("grid_"):
            H, W, P, s, k = (int(v) for v in key.split("_")[1:])
            a = D.split_aggregation_sampling(torch.zeros(1, 3, H, W), P, s, k, NoDiffusion(), "cpu")
            assert np.array_equal(np.asarray(a.patches_sr_infos, np.int32), g[key]), key
            assert len(a.patches_lr) == len(g[key])
            y0, _, x0, _ = a.patches_sr_infos[-1]
            assert torch.equal(a.patches_lr[-1], a.img_lr[:, :, y0 // k:y0 // k + P, x0 // k:x0 // k + P])
    a = D.split_aggregation_sampling(torch.zeros(1, 3, 64, 64), 32, 16, 2, NoDiffusion(), "cpu")
    assert np.array_equal(a.weight[0, 0].numpy().view(np.uint32), g["weight_64"].view(np.uint32))
    assert torch.equal(a.weight[0, 0], a.weight[0, 2])
    with pytest.raises(AssertionError):
        D.split_aggregation_sampling(torch.zeros(1, 3, 64, 64), 32, 48, 2, NoDiffusion(), "cpu")  # stride > patch
    with pytest.raises(ValueError):
        D.split_aggregation_sampling(torch.zeros(1, 3, 16, 64), 32, 16, 2, NoDiffusion(), "cpu")  # image < patch


def test_partition_blocks():
    assert D.partition_blocks(961, 8) == [(0, 121), (121, 241), (241, 361), (361, 481), (481, 601), (601, 721),
                                          (721, 841), (841, 961)]
    assert D.partition_blocks(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    for n in (0, 1, 7, 225, 961):
        for w in (1, 2, 3, 8):
            b = D.partition_blocks(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1


GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import diffusionremotesensing_b200 as D
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
n = 7
blocks = D.partition_blocks(n, 2)
lo, hi = blocks[rank]
# each "patch" carries its global index so the gathered order can be checked
local = torch.stack([torch.full((3, 4, 4), float(i)) for i in range(lo, hi)])
out = D.gather_blocks(local, [b - a for a, b in blocks], dst=0)
if rank == 0:
    assert out.shape == (n, 3, 4, 4), out.shape
    assert [int(out[i, 0, 0, 0]) for i in range(n)] == list(range(n))
else:
    assert out is None
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_patch_sharding_and_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), T.ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


GLOO_WORKER_EMPTY_BLOCK = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import diffusionremotesensing_b200 as D
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()

class StubDiffusion:
    # stands in for Diffusion.sample_batched on a host without a GPU: nearest x2 of the LR patch
    model = None
    calls = []
    def sample_batched(self, model, lr, input_channels=3, **kw):
        StubDiffusion.calls.append(lr.shape[0])
        return torch.nn.functional.interpolate(lr, scale_factor=2, mode="nearest")

# (scene side, patch, stride, patch_batch) -> patch counts 1 (fewer patches than ranks: rank 1's block is empty) and 9
for side, P, s, pb, n_want in ((32, 32, 16, 4, 1), (64, 32, 16, 2, 9)):
    img = torch.arange(3 * side * side, dtype=torch.float32).reshape(1, 3, side, side)
    agg = D.split_aggregation_sampling(img, P, s, 2, StubDiffusion(), "cpu", patch_batch=pb)
    n = len(agg.patches_lr)
    assert n == n_want, n
    blocks = D.partition_blocks(n, 2)
    lo, hi = blocks[rank]
    StubDiffusion.calls.clear()
    local = agg.sample_patches(range(lo, hi))
    assert local.shape == (hi - lo, 3, 2 * P, 2 * P), local.shape
    # one batch size per block (short batches are padded, the duplicate dropped)
    assert len(set(StubDiffusion.calls)) <= 1, StubDiffusion.calls
    out = D.gather_blocks(local, [b - a for a, b in blocks], dst=0)
    if rank == 0:
        assert out.shape == (n, 3, 2 * P, 2 * P)
        for i, (y0, y1, x0, x1) in enumerate(agg.patches_sr_infos):
            want = torch.nn.functional.interpolate(img[:, :, y0 // 2:y1 // 2, x0 // 2:x1 // 2], scale_factor=2)
            assert torch.equal(out[i:i + 1], want), i
    dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_sharded_patch_sampling_with_empty_and_ragged_blocks_gloo(tmp_path):
    script = tmp_path / "worker2.py"
    script.write_text(GLOO_WORKER_EMPTY_BLOCK)
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), T.ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_bench_clock_sampler_filters_samples_to_the_timed_region(tmp_path):
    """bench.py's nvidia-smi reader: samples are kept by timestamp window, throttle reasons are collected, and a region
    shorter than the sampling period falls back to every sample taken during the run."""
    import datetime
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("drs_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class FakeProc:
        def terminate(self): pass
        def wait(self, timeout=None): return 0
        def kill(self): pass

    t0 = datetime.datetime(2026, 1, 1, 12, 0, 0).timestamp()
    rows = [(t0 + 0.1, 1200, "Not Active"), (t0 + 1.1, 1965, "Not Active"), (t0 + 1.2, 1950, "Active"),
            (t0 + 2.5, 900, "Not Active")]
    path = tmp_path / "clocks.csv"
    with open(path, "w") as f:
        for ts, mhz, cap in rows:
            stamp = datetime.datetime.fromtimestamp(ts).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
            f.write(f"{stamp}, {mhz}, 1965, Not Active, Not Active, Not Active, {cap}\n")

    def sampler():
        s = bench.ClockSampler.__new__(bench.ClockSampler)
        s.proc, s.path, s.f = FakeProc(), str(path), open(path, "a")
        return s

    out = sampler().stop(t0 + 1.0, t0 + 2.0)
    assert out["samples"] == 2 and out["sm_mhz"] == 1957.5 and out["sm_max_mhz"] == 1965
    assert out["reasons"] == ["sw_power_cap"] and out["window"] == "timed region"
    # (stop() removes the file: write it again for the fallback case)
    with open(path, "w") as f:
        stamp = datetime.datetime.fromtimestamp(t0 + 5.0).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
        f.write(f"{stamp}, 1500, 1965, Not Active, Not Active, Not Active, Not Active\n")
    out = sampler().stop(t0 + 1.0, t0 + 1.001)
    assert out["samples"] == 1 and out["sm_mhz"] == 1500 and out["window"].startswith("whole run")


def test_default_patch_batching_at_the_cfg5_shape():
    """961 patches (LR 2048 x 2048, patch 128, stride 64) with the default patch_batch: 8 equal batches of 121 on one
    rank (7 padded duplicates dropped), one batch of 120 / 121 per rank on eight."""

    class Stub:
        model = None
        calls = []

        def sample_batched(self, model, lr, input_channels=3, **kw):
            Stub.calls.append(lr.shape[0])
            return torch.zeros(lr.shape[0], 3, 4, 4)

    agg = D.split_aggregation_sampling(torch.zeros(1, 3, 2048, 2048), 128, 64, 2, Stub(), "cpu")
    assert len(agg.patches_lr) == 961 and agg.patch_batch == 128
    # the stub returns 4 x 4 "patches": only the batching arithmetic is under test
    agg.patch_size = 2
    out = agg.sample_patches(range(961))
    assert out.shape[0] == 961 and Stub.calls == [121] * 8
    for lo, hi in D.partition_blocks(961, 8):
        Stub.calls.clear()
        assert agg.sample_patches(range(lo, hi)).shape[0] == hi - lo
        assert Stub.calls == [hi - lo]


def test_committed_ncu_traffic_follows_from_the_committed_capture():
    """bench.py's roofline.traffic comes from profiles/ncu_traffic.json: it must be what profiles/make_ncu_traffic.py
    derives from the committed conv-chain capture."""
    import json
    out = subprocess.run([sys.executable, os.path.join(T.ROOT, "profiles", "make_ncu_traffic.py"),
                          os.path.join("profiles", "ncu_r2_conv_chain_full.csv")], cwd=T.ROOT, capture_output=True,
                         text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    want = json.loads(out.stdout)
    have = json.load(open(os.path.join(T.ROOT, "profiles", "ncu_traffic.json")))
    assert have["launches"] == want["launches"] == 28
    for k in ("dram_bytes_per_launch", "dram_read_bytes_per_eval", "dram_write_bytes_per_eval"):
        assert abs(have[k] - want[k]) <= 1e-6 * want[k], k
