"""Layer-level parity of the tcgen05 implicit-GEMM kernel through the C ABI (drs_debug_conv2d).

Reference op per kind: torch.nn.functional.conv2d / conv_transpose2d in fp32 on bf16-rounded operands
(what nn.Conv2d / nn.ConvTranspose2d compute in UNet_model_superres.py:70-85,122-141,184-185,297-299), followed by
the eval-BatchNorm affine. Tolerance: bf16 output rounding (2^-8 relative) + fp32 accumulation-order noise.
"""
import pytest
import torch
import torch.nn.functional as F

from diffusionremotesensing_b200 import _native as N

pytestmark = pytest.mark.gpu

KINDS = {
    "3x3": N.CONV_3x3, "3x3s2": N.CONV_3x3_S2, "1x1": N.CONV_1x1, "2x2s2": N.CONV_2x2_S2, "T3x3s2": N.CONV_T3x3_S2,
}


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def ref_conv(x, w, b, kind):
    if kind == "3x3":
        return F.conv2d(x, w, b, padding=1)
    if kind == "3x3s2":
        return F.conv2d(x, w, b, stride=2, padding=1)
    if kind == "1x1":
        return F.conv2d(x, w, b)
    if kind == "2x2s2":
        return F.conv2d(x, w, b, stride=2)
    return F.conv_transpose2d(x, w, b, stride=2, padding=1, output_padding=1)


def run_native(x, w, b, scale, shift, kind, relu):
    B, Cin, H, W = x.shape
    Cout = w.shape[1] if kind == "T3x3s2" else w.shape[0]
    y_ref_shape = ref_conv(x[:1].cpu(), w.cpu(), None, kind).shape
    y = torch.empty((B, Cout, y_ref_shape[2], y_ref_shape[3]), device=x.device, dtype=torch.float32)
    wh, bh = w.cpu().contiguous(), b.cpu().contiguous()
    sh = scale.cpu().contiguous() if scale is not None else None
    th = shift.cpu().contiguous() if shift is not None else None
    N.check(N.lib().drs_debug_conv2d(N.ptr(x), N.ptr(wh), N.ptr(bh), N.ptr(sh), N.ptr(th), N.ptr(y), B, Cin, Cout, H, W,
                                     KINDS[kind], int(relu), x.device.index or 0, N.stream_ptr(x.device)))
    torch.cuda.synchronize()
    return y


CASES = [
    # kind, B, Cin, Cout, H, W
    ("3x3", 2, 16, 32, 32, 32),
    ("3x3", 1, 32, 32, 64, 64),
    ("3x3", 3, 64, 64, 24, 40),      # ragged: not a multiple of the 16x8 tile
    ("3x3", 2, 128, 128, 16, 16),
    ("3x3", 1, 256, 256, 8, 8),      # 64 pixels < one 128-pixel tile
    ("3x3", 5, 128, 256, 4, 4),      # tile spans 8 images, batch 5
    ("1x1", 2, 16, 32, 32, 32),
    ("1x1", 2, 256, 128, 8, 8),
    ("3x3s2", 2, 32, 32, 32, 32),
    ("3x3s2", 1, 128, 128, 16, 16),
    ("2x2s2", 2, 64, 64, 32, 32),
    ("2x2s2", 3, 128, 128, 8, 8),
    ("T3x3s2", 2, 64, 64, 16, 16),
    ("T3x3s2", 1, 256, 256, 8, 8),
    ("T3x3s2", 2, 128, 128, 12, 20),
    # full-width rows (multiples of 128 pixels, <= 64 output channels): the shapes the row-streaming kernel takes
    # (conv_row.cu; by default only above a size threshold, with DRS_ROW=force always -- test_gpu_kernel_variants.py)
    ("3x3", 2, 16, 32, 24, 128),
    ("3x3", 1, 32, 32, 40, 256),
    ("3x3", 3, 64, 64, 17, 128),     # odd row count: ranges end inside a strip
    ("3x3", 2, 64, 32, 5, 256),
    ("3x3", 1, 32, 64, 4, 128),
    ("3x3", 4, 16, 16, 33, 256),
]


@pytest.mark.parametrize("kind,B,Cin,Cout,H,W", CASES)
def test_conv_matches_torch(cuda_device, kind, B, Cin, Cout, H, W):
    g = torch.Generator().manual_seed(1234 + B * 7 + Cin + H)
    x = bf16r(torch.randn((B, Cin, H, W), generator=g)).to(cuda_device)
    wshape = (Cin, Cout, 3, 3) if kind == "T3x3s2" else (Cout, Cin) + {"3x3": (3, 3), "3x3s2": (3, 3), "1x1": (1, 1),
                                                                      "2x2s2": (2, 2)}[kind]
    w = bf16r(torch.randn(wshape, generator=g) / (Cin * 3) ** 0.5).to(cuda_device)
    b = torch.randn((Cout,), generator=g).to(cuda_device)
    scale = (torch.rand((Cout,), generator=g) + 0.5).to(cuda_device)
    shift = torch.randn((Cout,), generator=g).to(cuda_device)
    for relu, use_affine in ((False, False), (True, True)):
        y = run_native(x, w, b, scale if use_affine else None, shift if use_affine else None, kind, relu)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        r = ref_conv(x.double(), w.double(), b.double(), kind)
        if use_affine:
            r = r * scale.double()[None, :, None, None] + shift.double()[None, :, None, None]
        if relu:
            r = r.clamp_min(0)
        err = (y.double() - r).abs().max().item()
        ref_mag = r.abs().max().item()
        assert y.shape == r.shape
        assert err <= 6e-3 * ref_mag + 1e-3, f"{kind} relu={relu}: max err {err:.3e} vs magnitude {ref_mag:.3e}"


@pytest.mark.parametrize("kind,B,Cin,Cout,H,W", [("3x3s2", 16, 64, 64, 128, 128), ("3x3", 16, 64, 64, 128, 128),
                                                 ("T3x3s2", 16, 64, 64, 64, 64), ("3x3", 8, 64, 32, 256, 256)])
def test_persistent_kernel_is_deterministic_and_correct_on_many_tiles(cuda_device, kind, B, Cin, Cout, H, W):
    """Many pixel tiles per persistent CTA (tile pairs, odd tile counts, shared-memory ring reuse): the result must be
    bit-identical from launch to launch and match torch. Regression test for a ring-phase race that corrupted the
    second tile of a pair when the next pair of the same CTA had no second tile."""
    g = torch.Generator().manual_seed(99)
    x = bf16r(torch.randn((B, Cin, H, W), generator=g)).to(cuda_device)
    wshape = (Cin, Cout, 3, 3) if kind == "T3x3s2" else (Cout, Cin, 3, 3)
    w = bf16r(torch.randn(wshape, generator=g) / (Cin * 3) ** 0.5).to(cuda_device)
    b = torch.randn((Cout,), generator=g).to(cuda_device)
    r = ref_conv(x.double(), w.double(), b.double(), kind)
    first = None
    for _ in range(25):
        y = run_native(x, w, b, None, None, kind, False)
        if first is None:
            first = y.clone()
            assert (y.double() - r).abs().max().item() <= 6e-3 * r.abs().max().item() + 1e-3
        else:
            assert torch.equal(y, first), "result changed between identical launches"
