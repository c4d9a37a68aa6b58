"""Drop-in ``split_aggregation_sampling`` (Aggregation_Sampling.py:9-138) on the CUDA path.

Same constructor, public attributes (``patches_lr``, ``patches_sr_infos``, ``weight``) and methods (``patchifier``,
``aggregation_sampling``, ``gaussian_weights``). Differences in HOW, not WHAT:

  * the reference super-resolves the patches one after another with ``sample(1, ...)``; here they run as batches
    through ``Diffusion.sample_batched`` (every patch is an independent chain, eval-mode BatchNorm has no
    cross-sample statistics);
  * with ``torch.distributed`` initialised the row-major patch list is block-partitioned over the ranks, each rank
    samples its block on its own GPU, one gather brings the finished patches to rank 0, which blends them in the
    reference's patch order and broadcasts the scene. The BLEND is bit-identical whatever the world size; the whole
    scene is bit-identical across world sizes only with injected noise (``noise=`` / ``x_T=`` hooks): without them
    every block draws its start states and per-step noise from a private DEVICE generator seeded
    ``torch.initial_seed() + 1 + first patch index of the block``, so that identically seeded ranks do not reuse one
    noise stream for different patches (the reference's sequential per-patch CPU / device RNG stream cannot be
    reproduced by concurrent chains, and a host randn of a whole patch batch would cost more than its sampling);
  * the Gaussian overlap blend, the division and the clamp are one CUDA kernel (``drs_blend``) that accumulates
    each output pixel in patch order with separately rounded fp32 multiply and add, like the reference's
    ``im_res[...] += patch_sr * weight`` sequence.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
from numpy import exp, pi, sqrt

from . import _native as N


def partition_blocks(n_items: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous block partition of range(n_items): the first n_items % world_size ranks get one extra item."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    base, extra = divmod(n_items, world_size)
    out, start = [], 0
    for r in range(world_size):
        size = base + (1 if r < extra else 0)
        out.append((start, start + size))
        start += size
    return out


GATHER_CHUNK_BYTES = 64 << 20


def gather_blocks(local: torch.Tensor, counts: Sequence[int], dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Gathers per-rank blocks [counts[r], ...] to rank `dst` in rank order (the one collective of the path).
    Point-to-point: every rank sends its block, `dst` receives it straight into its slice of the result (no padding
    to the largest block, no staging copies); empty blocks are skipped on both sides. Blocks travel in pieces of at
    most 64 MB (whole items): the first NCCL transfer of a several-hundred-MB buffer was measured at ~3 GB/s on two
    GPUs (115 ms for 378 MB) against 260-420 GB/s for 94 MB messages."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    local = local.contiguous()
    item_bytes = max(1, local[0].numel() * local.element_size()) if local.shape[0] else 1
    per_chunk = max(1, GATHER_CHUNK_BYTES // item_bytes)
    peer = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    if rank != dst:
        for i in range(0, counts[rank], per_chunk):
            dist.send(local[i:min(i + per_chunk, counts[rank])], dst=peer(dst), group=group)
        return None
    out = torch.empty((sum(counts),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    out[offs[dst]:offs[dst + 1]] = local
    reqs = []
    for r in range(world):
        if r == dst:
            continue
        for i in range(0, counts[r], per_chunk):
            reqs.append(dist.irecv(out[offs[r] + i:offs[r] + min(i + per_chunk, counts[r])], src=peer(r), group=group))
    for q in reqs:
        q.wait()
    return out


def blend_patches(patches: torch.Tensor, infos: Sequence[Tuple[int, int, int, int]], weight2d: torch.Tensor,
                  height: int, width: int, clamp: bool = True, out: Optional[torch.Tensor] = None,
                  wsum: Optional[torch.Tensor] = None, coords: Optional[torch.Tensor] = None):
    """drs_blend on device tensors: patches [n, C, P, P] fp32, weight2d [P, P] fp32 -> ([1, C, H, W], [H, W]).
    out / wsum / coords (int32 [n, 4] on the host) may be passed in to reuse buffers across calls."""
    dev = patches.device
    if dev.type != "cuda":
        raise RuntimeError("drs_blend runs on a CUDA device only; there is no CPU fallback")
    n, ch, P, _ = patches.shape
    if coords is None:
        coords = torch.tensor([list(i) for i in infos], dtype=torch.int32).contiguous()
    if out is None:
        out = torch.empty((1, ch, height, width), device=dev, dtype=torch.float32)
    if wsum is None:
        wsum = torch.empty((height, width), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        N.check(N.lib().drs_blend(N.ptr(patches.contiguous()), N.ptr(coords), n, N.ptr(weight2d.contiguous()),
                                  N.ptr(out), N.ptr(wsum), ch, height, width, P, 1 if clamp else 0,
                                  N.stream_ptr(dev)))
    return out, wsum


class split_aggregation_sampling:
    def __init__(self, img_lr, patch_size, stride, magnification_factor, diffusion_model, device, patch_batch=128):
        assert stride <= patch_size
        self.img_lr = img_lr
        self.patch_size = patch_size
        self.stride = stride
        self.magnification_factor = magnification_factor
        self.diffusion_model = diffusion_model
        self.device = device
        self.model = diffusion_model.model
        self.patch_batch = patch_batch
        batch_size, channels, height, width = img_lr.shape
        if height < patch_size or width < patch_size:
            raise ValueError("the image (%d x %d) is smaller than one patch (%d)" % (height, width, patch_size))
        self.patches_lr, self.patches_sr_infos = self.patchifier(img_lr, patch_size, stride, magnification_factor)
        self.weight = self.gaussian_weights(patch_size * magnification_factor, patch_size * magnification_factor,
                                            batch_size)

    def patchifier(self, img_to_split, patch_size, stride=None, magnification_factor=1):
        """Row-major windows with the last row / column clamped to the border and duplicates dropped
        (Aggregation_Sampling.py:30-74). Returns views of the image and the SR-space (y0, y1, x0, x1) tuples."""
        stride = patch_size if stride is None else stride
        _, _, height, width = img_to_split.shape
        k = magnification_factor
        starts_y = [min(y, height - patch_size) for y in range(0, height + 1, stride)]
        starts_x = [min(x, width - patch_size) for x in range(0, width + 1, stride)]
        patches_lr, infos, seen = [], [], set()
        for ys in starts_y:
            for xs in starts_x:
                info = (ys * k, (ys + patch_size) * k, xs * k, (xs + patch_size) * k)
                if info in seen:
                    continue
                seen.add(info)
                patches_lr.append(img_to_split[:, :, ys:ys + patch_size, xs:xs + patch_size])
                infos.append(info)
        return patches_lr, infos

    def gaussian_weights(self, tile_width, tile_height, nbatches):
        """float64 outer product of two Gaussians (x midpoint (W-1)/2, y midpoint H/2, variance 0.01 of the
        normalised coordinate), cast to fp32 and tiled to [nbatches, 3, H, W] (Aggregation_Sampling.py:118-138).
        Evaluated with numpy scalars on the host exactly like the reference; never recomputed on the device."""
        var = 0.01
        mid_x = (tile_width - 1) / 2
        mid_y = tile_height / 2
        norm = sqrt(2 * pi * var)
        x_probs = [exp(-(x - mid_x) * (x - mid_x) / (tile_width * tile_width) / (2 * var)) / norm
                   for x in range(tile_width)]
        y_probs = [exp(-(y - mid_y) * (y - mid_y) / (tile_height * tile_height) / (2 * var)) / norm
                   for y in range(tile_height)]
        weights = torch.tensor(np.outer(y_probs, x_probs)).to(torch.float32).to(self.device)
        return torch.tile(weights, (nbatches, 3, 1, 1))

    # -- sampling ------------------------------------------------------------------------------------------------
    def sample_patches(self, indices: Sequence[int], noise: Optional[Callable] = None,
                       x_T: Optional[Callable] = None, private_rng: bool = False) -> torch.Tensor:
        """SR patches [len(indices), C, P*k, P*k] for the given patch indices.
        The block is cut into ceil(n / patch_batch) batches of ONE size (the shorter ones are padded by repeating
        their last patch; the duplicate is dropped), so a single plan, time table and pair of CUDA graphs serve the
        whole block (961 patches, patch_batch 128 -> 8 batches of 121). An empty block returns [0, C, P*k, P*k].
        patch_batch (an extension: the reference samples patch by patch) defaults to 128: at 128 -> 256 a reverse step
        costs 41.5 us per patch in batches of 31, 38.4 at 61 and 37.4-38.1 from 111 up (scripts/diag_batch_sweep.py;
        ~35 MB of plan workspace per patch), and every sample() call carries ~3 ms + 0.16 ms per patch of set-up.
        noise(patch_index, step) / x_T(patch_index) inject per-patch noise (parity tests); private_rng draws from
        generators seeded by the block's first patch index instead of the global ones (sharded runs)."""
        indices = list(indices)
        k = self.magnification_factor
        channels = self.img_lr.shape[1]
        side = self.patch_size * k
        if not indices:
            return torch.empty((0, channels, side, side), device=self.device, dtype=torch.float32)
        dm = self.diffusion_model
        n = len(indices)
        n_batches = -(-n // self.patch_batch)
        size = -(-n // n_batches)
        gen = None
        if private_rng and noise is None:
            # start states and per-step noise from one device generator seeded by the block's first patch index
            gen = torch.Generator(device=self.device).manual_seed(torch.initial_seed() + 1 + indices[0])
        outs = []
        for b0 in range(0, n, size):
            idx = indices[b0:b0 + size]
            real = len(idx)
            idx = idx + [idx[-1]] * (size - real)
            lr = torch.cat([self.patches_lr[i] for i in idx], dim=0).to(self.device)
            xt = None if x_T is None else torch.cat([x_T(i) for i in idx], dim=0)
            nz = None if noise is None else (lambda step, idx=idx: torch.cat([noise(i, step) for i in idx], dim=0))
            sr = dm.sample_batched(self.model, lr, input_channels=lr.shape[1], x_T=xt, noise=nz, generator=gen)
            outs.append(sr[:real])
        return torch.cat(outs, dim=0)

    def aggregation_sampling(self, noise: Optional[Callable] = None, x_T: Optional[Callable] = None, group=None):
        batch_size, channels, height, width = self.img_lr.shape
        if batch_size != 1:
            raise ValueError("aggregation sampling handles one scene at a time (batch size 1), like the reference")
        k = self.magnification_factor
        n = len(self.patches_lr)
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if distributed:
            rank, world = dist.get_rank(group), dist.get_world_size(group)
            blocks = partition_blocks(n, world)
            lo, hi = blocks[rank]
            local = self.sample_patches(range(lo, hi), noise, x_T, private_rng=True)
            patches = gather_blocks(local, [b - a for a, b in blocks], dst=0, group=group)
        else:
            rank = 0
            patches = self.sample_patches(range(n), noise, x_T, private_rng=True)
        H, W = height * k, width * k
        if rank == 0:
            im_res, _ = blend_patches(patches, self.patches_sr_infos, self.weight[0, 0], H, W, clamp=True)
        else:
            im_res = torch.empty((1, channels, H, W), device=self.device, dtype=torch.float32)
        if distributed:
            dist.broadcast(im_res, src=0, group=group)
        return im_res
