"""The alternative code paths of the convolution kernel, each forced through its environment switch in a child
process (the switches are read once per process) and checked with the same layer-level parity cases as the default
path (tests/test_gpu_conv_layers.py):

  * DRS_CG2=all             CTA-pair kernel (tcgen05.mma.cta_group::2, conv_gemm2c.cu) wherever it is applicable
  * DRS_V2_NO_TMA_STORE=1   per-thread global stores instead of the staged TMA-store epilogue
  * DRS_V2_GENERIC_EPILOGUE run-time flag epilogue instead of the compile-time variants
  * DRS_V2_NO_SOLO=1        transposed convolutions drained by one epilogue group per tile
  * DRS_DISABLE_V2=1        first-generation kernel for every layer
  * DRS_NO_NARROW=1         no 32-channel launch variants on small grids (the default test sizes otherwise use them)
  * DRS_NO_GATE_FUSION=1    attention gate as two launches (psi map through memory) instead of the fused program
  * DRS_ROW=force / DRS_ROW=0  row-streaming kernel (conv_row.cu) on every layer it can express whatever the grid
                            size / on none (the second-generation kernel for everything)
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_child(extra_env, args):
    env = dict(os.environ)
    env.update(extra_env)
    env["DRS_V2_VERBOSE"] = "1"
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-s", "-m", "gpu", "-p", "no:cacheprovider"] + args
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)


@pytest.mark.parametrize("env", [{"DRS_V2_NO_TMA_STORE": "1"}, {"DRS_V2_GENERIC_EPILOGUE": "1"}, {"DRS_V2_NO_SOLO": "1"},
                                 {"DRS_DISABLE_V2": "1"}, {"DRS_V2_NO_PDL": "1"}])
def test_layer_parity_on_alternative_paths(cuda_device, env):
    r = run_child(env, ["tests/test_gpu_conv_layers.py"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_cta_pair_kernel_layer_parity(cuda_device):
    r = run_child({"DRS_CG2": "all"}, ["tests/test_gpu_conv_layers.py"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    # the switch must really have routed launches through the CTA-pair kernel
    assert "CTA-pair kernel" in r.stderr or "CTA-pair kernel" in r.stdout, "no launch used the CTA-pair kernel"


def test_cta_pair_kernel_unet_and_sampler_parity(cuda_device):
    # the sampler tests re-allocate the time table after the plan was bound: every launch argument block, the
    # CTA-pair one included, has to follow (regression test for a stale pointer)
    r = run_child({"DRS_CG2": "all"}, ["tests/test_gpu_unet.py", "tests/test_gpu_sampler.py"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "CTA-pair kernel" in r.stderr or "CTA-pair kernel" in r.stdout


def test_unet_parity_without_narrow_variants(cuda_device):
    r = run_child({"DRS_NO_NARROW": "1"}, ["tests/test_gpu_unet.py"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_row_kernel_layer_and_unet_parity_forced(cuda_device):
    r = run_child({"DRS_ROW": "force"}, ["tests/test_gpu_conv_layers.py", "tests/test_gpu_unet.py",
                                         "tests/test_gpu_sampler.py"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "row kernel" in r.stderr or "row kernel" in r.stdout, "no launch used the row-streaming kernel"


def test_parity_without_row_kernel(cuda_device):
    r = run_child({"DRS_ROW": "0"}, ["tests/test_gpu_unet.py", "tests/test_gpu_full_size.py::test_full_resolution_eps_matches_oracle",
                                     "tests/test_gpu_baseline_configs.py::test_cfg2_batch16_trajectory_vs_oracle"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "row kernel" not in r.stderr and "row kernel" not in r.stdout


def test_unet_parity_without_gate_fusion(cuda_device):
    r = run_child({"DRS_NO_GATE_FUSION": "1"}, ["tests/test_gpu_unet.py",
                                                "tests/test_gpu_full_size.py::test_full_resolution_eps_matches_oracle"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
