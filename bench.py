#!/usr/bin/env python
"""Benchmark of the reverse-diffusion sampling hot path (BASELINE.json metric, config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (N = 1): super-resolution x2, LR 128 -> 256, batch n = 16 stochastic samples of one LR image, 1500-step cosine
schedule, synthetic Sentinel-2-shaped RGB input, random-init weights (reference default init under manual_seed(0) +
randomised BatchNorm statistics). A "step" is one reverse step of the whole batch: one UNet
evaluation + the DDPM posterior update with freshly drawn noise. With N > 1 every rank runs the same per-GPU batch on
its own GPU (weak scaling, no data-path collective).

  value      image-steps per second over all GPUs with x_t, the condition features and the time tables resident in HBM
  e2e        the same metric through the public API, Diffusion.sample(n, model, lr_img) with noise_steps = K + 1,
             from pinned host tensors to a host result: H2D of lr_img and x_T, condition encode, time-table
             preparation, K graph-replayed steps, D2H of the samples -- all inside the timed region
  roofline   tensor-core bound: algorithmic conv FLOPs of one UNet evaluation / CUDA-event time of the chain of
             tcgen05 implicit-GEMM launches of one forward (drs_plan_time_forward: events around the chain on the
             launching stream). Both fractions are printed -- frac_sustained (MEASURED_PEAKS.json
             bf16_tflops_sustained: a 4 s back-to-back cuBLAS run that pulls the clock down) and frac_burst
             (bf16_tflops: best single GEMM at full clock); `frac` is the one that matches this run's own clock record
             (SM clock at its maximum and no power cap -> burst). `traffic` = DRAM bytes per launch from the committed
             ncu capture (profiles/ncu_traffic.json).
             roofline.hbm_kernels: the HBM-bound kernels of the path -- conv0, the posterior update (both at cfg 2) and
             the aggregation blend (cfg-5 shape, 961 patches -> 3 x 4096 x 4096) -- algorithmic bytes / CUDA-event
             time on a flushed L2 / MEASURED_PEAKS.json hbm_gbs
  aggregation  cfg 5 as a STRONG-scaling leg: the 961 overlapping 128 -> 256 patches of an LR 2048 x 2048 scene, block
             partitioned over the N ranks, sampled for 50 reverse steps per patch, gathered to rank 0 (NCCL) and blended
             there; seconds = max over ranks of the whole thing (gather and blend inside the timed region)
  cpu_baseline  the reference algorithm (oracle port: the same torch CPU ops the reference's modules call) timed on
             this box's host cores on a bounded sample of the same workload

--impl reference prints the CPU arm alone (rank 0 only under torchrun).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "unet_denoise_image_steps_per_sec"
UNIT = "image-steps/s"
NOISE_STEPS = 1500
BATCH = 16
S = 256
MAG = 2
WORKLOAD = "superres_x2_LR128to256_batch16_cosine1500"


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops"))),
                "tflops_burst": float(p.get("bf16_tflops", p.get("bf16_tflops_sustained"))),
                "hbm_gbs": float(p["hbm_gbs"]), "source": "MEASURED_PEAKS.json"}
    return {"tflops": 1400.0, "tflops_burst": 1650.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons. The sampler is started before the warm-up (nvidia-smi needs a few hundred
    milliseconds to produce its first line) and the samples are filtered to the timed region by their timestamps."""

    def __init__(self, index):
        self.proc = None
        self.path = os.path.join("/tmp", f"drs_clocks_{os.getpid()}.csv")
        q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.f = open(self.path, "w", buffering=1)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_first(self, timeout=3.0):
        """Blocks until nvidia-smi has produced its first line (it needs a few hundred milliseconds to start)."""
        t0 = time.time()
        while self.proc is not None and time.time() - t0 < timeout:
            try:
                if os.path.getsize(self.path) > 0:
                    return
            except OSError:
                pass
            time.sleep(0.02)

    def stop(self, t_begin=None, t_end=None):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[1]), float(parts[2]),
                             {n for n, v in zip(names, parts[3:7]) if v.lower().startswith("active")}))
            except ValueError:
                continue
        window = "timed region"
        sel = [r for r in rows if t_begin is not None and t_begin <= r[0] <= t_end]
        if not sel:
            # region shorter than the sampling period: fall back to every sample taken under load (warm-up included)
            sel, window = rows, "whole run (timed region shorter than the sampling period)"
        if sel:
            reasons = set()
            for r in sel:
                reasons |= r[3]
            out = {"sm_mhz": statistics.median([r[1] for r in sel]), "sm_max_mhz": max(r[2] for r in sel),
                   "reasons": sorted(reasons), "samples": len(sel), "window": window}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def synthetic_model_and_inputs():
    import torch
    from diffusionremotesensing_b200 import synthetic as T
    model, sd = T.default_init_model("superres", seed=0, bn_seed=1)
    lr = T.np_rand(2, 3, S // MAG, S // MAG)       # Sentinel-2-shaped RGB in [0, 1)
    return model, sd, lr


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_arm(steps, warmup, budget_s, n_images=None):
    """Times `steps` reverse steps of the oracle port on n_images of the 16-image batch (bounded sample)."""
    import torch
    from diffusionremotesensing_b200 import synthetic as T
    from oracle import restatement as R
    torch.set_grad_enabled(False)
    # all host cores: torchrun exports OMP_NUM_THREADS=1 to its workers, which must not reach the CPU arm
    host_cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(1, host_cores))
    cores = torch.get_num_threads()
    _, sd, lr = synthetic_model_and_inputs()
    sched = R.noise_schedule("cosine", NOISE_STEPS)
    cond = lr.unsqueeze(0)
    x1 = T.np_randn(3, 1, 3, S, S)
    t0 = time.perf_counter()
    R.unet_forward(sd, "superres", x1, torch.tensor([NOISE_STEPS - 1]), cond, MAG)
    t_img = time.perf_counter() - t0          # includes first-call overhead: an upper bound
    if n_images is None:
        n_images = int(max(1, min(BATCH, budget_s / max(1e-3, (steps + warmup) * t_img))))
    x = T.np_randn(3, n_images, 3, S, S)
    alpha, alpha_hat, beta = sched

    def one_step(x, i):
        t = (torch.ones(n_images) * i).long()
        eps = R.unet_forward(sd, "superres", x, t, cond, MAG)
        z = torch.randn_like(x)
        return R.posterior_update(x, eps, z, alpha[t][:, None, None, None], alpha_hat[t][:, None, None, None],
                                  beta[t][:, None, None, None])

    i = NOISE_STEPS - 1
    for _ in range(warmup):
        x = one_step(x, i)
        i -= 1
    t0 = time.perf_counter()
    for _ in range(steps):
        x = one_step(x, i)
        i -= 1
    dt = time.perf_counter() - t0
    value = n_images * steps / dt
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} reverse steps (after {warmup} warm-up) of {n_images} of the {BATCH} images of {WORKLOAD}, "
                      f"oracle/restatement.py (the torch CPU ops the reference modules call), fp32, {cores} threads",
            "ms_per_step": 1e3 * dt / steps, "n_images": n_images}


def run_reference(args):
    rank, world, _ = rank_world()
    if rank != 0:
        return
    cb = cpu_reference_arm(args.steps, args.warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_step_sample": cb["n_images"], "noise_steps": NOISE_STEPS},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
SCENE_ARG = [2048]


def aggregation_leg(model, dev, k_steps, rank, world, roofline):
    """cfg 5 (Aggregation_Sampling.py:94-110): the 961 overlapping 128 -> 256 patches of an LR 2048 x 2048 scene.
    The row-major patch list is block-partitioned over the ranks; every rank samples its block in equal batches
    (k_steps reverse steps of a (k_steps + 1)-step cosine schedule), one gather brings the SR patches to rank 0, which
    blends them. Timed from the first patch to the blended scene, max over ranks: a fixed amount of work whatever N."""
    import torch
    import torch.distributed as dist
    from diffusionremotesensing_b200 import synthetic as T
    import diffusionremotesensing_b200 as D
    from diffusionremotesensing_b200.aggregation import blend_patches, gather_blocks, partition_blocks
    side, P, stride, k = SCENE_ARG[0], 128, 64, MAG
    scene = T.np_rand(7, 1, 3, side, side).to(dev)
    d = D.Diffusion("cosine", model, "/nonexistent", noise_steps=k_steps + 1, device=str(dev),
                    magnification_factor=k, image_size=P * k, Degradation_type="DownBlur")
    agg = D.split_aggregation_sampling(scene, P, stride, k, d, str(dev), patch_batch=128)
    n = len(agg.patches_lr)
    blocks = partition_blocks(n, world)
    counts = [b - a for a, b in blocks]
    lo, hi = blocks[rank]
    n_batches = max(1, -(-(hi - lo) // agg.patch_batch))
    size = -(-(hi - lo) // n_batches)
    # warm-up: one batch of this rank's batch size builds its plan, time table and CUDA graphs; one small gather opens
    # the NCCL point-to-point connections to rank 0 (hundreds of milliseconds on first use)
    warm = agg.sample_patches(range(lo, min(hi, lo + size)), private_rng=True)
    if world > 1:
        gather_blocks(warm[:1], [1] * world, dst=0)
    H = W = side * k
    if rank == 0:
        # ... and one blend of the scene geometry builds the window tables and loads the kernel (one-time set-up, like
        # graph capture; the buffers go back to torch's caching allocator)
        dummy = torch.zeros((n, 3, P * k, P * k), device=dev)
        blend_patches(dummy, agg.patches_sr_infos, agg.weight[0, 0].contiguous(), H, W, clamp=True)
        del dummy
    del warm
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    local = agg.sample_patches(range(lo, hi), private_rng=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if world > 1:
        dist.barrier()      # so that gather_ms is the transfer, not the wait for the slowest rank (that is in `seconds`)
        torch.cuda.synchronize()
    t1b = time.perf_counter()
    patches = gather_blocks(local, counts, dst=0) if world > 1 else local
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    blend_ms = None
    if rank == 0:
        w2d = agg.weight[0, 0].contiguous()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, wsum = blend_patches(patches, agg.patches_sr_infos, w2d, H, W, clamp=True)
        e1.record()
        torch.cuda.synchronize()
        blend_ms = e0.elapsed_time(e1)
    t3 = time.perf_counter()
    if rank == 0:
        assert torch.isfinite(out).all()
    tt = torch.tensor([t3 - t0, t1 - t0], device=dev, dtype=torch.float64)
    phases = torch.tensor([t1 - t0, t1b - t1, t2 - t1b, t3 - t2], device=dev, dtype=torch.float64)
    per_rank = [phases.clone() for _ in range(world)]
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_gather(per_rank, phases)
    total, sample_s = float(tt[0]), float(tt[1])
    res = {"workload": f"cfg 5: LR {side}x{side} scene -> {n} patches {P}->{P * k}, stride {stride}, "
                       f"{k_steps} reverse steps per patch", "scaling": "strong", "patches": n, "steps": k_steps,
           "n_gpus": world, "patches_per_rank": counts, "patch_batch": size, "seconds": total,
           "sample_seconds_max_rank": sample_s, "gather_ms": (t2 - t1b) * 1e3, "blend_ms": blend_ms,
           "image_steps_per_sec": n * k_steps / total,
           "per_rank_ms": {"columns": ["sample", "wait_for_slowest_rank", "gather", "blend"],
                           "rows": [[round(float(v) * 1e3, 2) for v in r] for r in per_rank]},
           "note": "seconds = first patch to blended scene on rank 0, max over ranks; gather_ms / blend_ms are rank 0's"}
    if rank == 0 and roofline is not None:
        # blend kernel on a flushed L2, cache-hit call (no table upload): algorithmic bytes = patches in + scene and
        # weight-sum map out (Aggregation_Sampling.py:96-110 materialises pixel_count too)
        from diffusionremotesensing_b200 import _native as N
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
        coords = torch.tensor([list(i) for i in agg.patches_sr_infos], dtype=torch.int32).contiguous()
        st = N.stream_ptr(dev)
        times = []
        for i in range(5):
            # the flush READS 512 MB (no dirty lines left behind) and keeps the GPU busy while the host enqueues the
            # blend, so the events bracket the kernel alone
            N.check(N.lib().drs_debug_l2_flush(N.ptr(flush), flush.numel(), st))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            blend_patches(patches, agg.patches_sr_infos, w2d, H, W, clamp=True, out=out, wsum=wsum, coords=coords)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms_b = min(times[1:])
        by = float(patches.numel() * 4 + 3 * H * W * 4 + H * W * 4)
        hbm = roofline["hbm_kernels"]["peak_gbs"]
        roofline["hbm_kernels"]["blend_gather4_kernel"] = {
            "bytes": by, "ms": ms_b, "gbs": by / (ms_b * 1e6), "frac": by / (ms_b * 1e6) / hbm,
            "what": f"{n} SR patches [3,{P * k},{P * k}] fp32 in, scene [3,{H},{W}] + weight-sum map out "
                    "(cfg-5 shape), best of 4 after one warm call, L2 flushed (512 MB read) before each"}
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    from diffusionremotesensing_b200 import synthetic as T
    import diffusionremotesensing_b200 as D
    from diffusionremotesensing_b200 import _native as N
    import ctypes as C

    rank, world, local = rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    # stdout carries exactly one JSON line: anything libraries print on file descriptor 1 meanwhile (NCCL's version
    # banner, for one) is sent to stderr, and the line itself is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = N.lib()

    model, sd, lr = synthetic_model_and_inputs()
    model.to(dev).eval()
    K, W = args.steps, args.warmup
    diffusion = D.Diffusion("cosine", model, "/nonexistent", noise_steps=NOISE_STEPS, device=str(dev),
                            magnification_factor=MAG, image_size=S, Degradation_type="DownBlur")

    # ---- device-resident leg: K graph-replayed steps -------------------------------------------------------
    st = N.stream_ptr(dev)
    plan = model.native_plan(BATCH, BATCH, 1, S, MAG)
    lr_dev = lr.unsqueeze(0).to(dev).contiguous()
    c1, c2, c3 = diffusion._coefficients()
    x = T.np_randn(3 + rank, BATCH, 3, S, S).to(dev).contiguous()
    z = torch.empty_like(x)
    eps = torch.empty_like(x)
    N.check(lib.drs_cond_encode(plan, N.ptr(lr_dev), st))
    N.check(lib.drs_sampler_prepare(plan, NOISE_STEPS, N.ptr(c1), N.ptr(c2), N.ptr(c3), None, 0.0, st))
    N.check(lib.drs_sampler_begin(plan, N.ptr(x), N.ptr(z), N.ptr(eps), NOISE_STEPS - 1, st))
    left = [NOISE_STEPS - 2]   # noisy steps before the chain reaches its last (noise-free) step

    def step():
        if left[0] <= 0:
            # more steps requested than one 1499-evaluation chain holds: start the next chain on the same state
            N.check(lib.drs_sampler_begin(plan, N.ptr(x), N.ptr(z), N.ptr(eps), NOISE_STEPS - 1, st))
            left[0] = NOISE_STEPS - 2
        left[0] -= 1
        z.normal_()
        N.check(lib.drs_sampler_step(plan, 1, st))

    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.wait_first()
    for _ in range(W):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    N.check(lib.drs_plan_check(plan, st))
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    clock_info = clocks.stop(t_begin, t_end) if clocks else None
    launches_per_step = lib.drs_sampler_launches_per_step(plan)
    value = world * BATCH * K / (ms * 1e-3)

    # ---- roofline of the tensor-core launches (rank 0, live CUDA events inside the library) ----------------
    roofline, layer_rows = None, []
    if rank == 0:
        n_l = lib.drs_plan_launch_count(plan)
        ms_out = torch.zeros(n_l, dtype=torch.float32)
        N.check(lib.drs_plan_profile(plan, N.ptr(x), N.ptr(eps), max(3, min(K, 10)), N.ptr(ms_out), st))
        N.check(lib.drs_plan_check(plan, st))
        tot_f = tot_ms = 0.0
        for i in range(n_l):
            name = C.create_string_buffer(64)
            fl, by, ctas, smem = C.c_double(), C.c_double(), C.c_int(), C.c_int()
            N.check(lib.drs_plan_launch_info(plan, i, name, 64, C.byref(fl), C.byref(by), C.byref(ctas), C.byref(smem)))
            row = {"launch": name.value.decode(), "ms": float(ms_out[i]), "gflop": fl.value / 1e9, "mbytes": by.value / 1e6,
                   "ctas": ctas.value, "smem": smem.value}
            row["tflops"] = row["gflop"] / max(row["ms"], 1e-9)
            row["gbs"] = row["mbytes"] / max(row["ms"], 1e-9)
            layer_rows.append(row)
            if i > 0:
                tot_f += fl.value
                tot_ms += float(ms_out[i])
        peaks = measured_peaks()
        # duration of the tensor-core launches as the sampler runs them: CUDA events around the whole chain of one
        # forward on the launching stream (no event between launches, gate branch on its side stream)
        ms2 = torch.zeros(2, dtype=torch.float32)
        N.check(lib.drs_plan_time_forward(plan, N.ptr(x), N.ptr(eps), max(5, min(K, 20)), N.ptr(ms2), st))
        N.check(lib.drs_plan_check(plan, st))
        chain_ms = float(ms2[1])
        n_conv = n_l - 1
        achieved = tot_f / (chain_ms * 1e-3) / 1e12
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        # which measured peak matches THIS run: the sustained figure belongs to a power-limited clock; a run whose SM
        # clock sat at its maximum with no power cap is compared with the burst figure
        at_max_clock = bool(clock_info and clock_info.get("sm_mhz") and clock_info.get("sm_max_mhz") and
                            clock_info["sm_mhz"] >= 0.97 * clock_info["sm_max_mhz"] and
                            "sw_power_cap" not in clock_info.get("reasons", []))
        frac_s, frac_b = achieved / peaks["tflops"], achieved / peaks["tflops_burst"]
        # HBM-bound kernels: conv0 and the posterior update at cfg 2, each launch on a flushed L2
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
        ms_h = torch.zeros(2, dtype=torch.float32)
        N.check(lib.drs_sampler_begin(plan, N.ptr(x), N.ptr(z), N.ptr(eps), NOISE_STEPS - 1, st))
        N.check(lib.drs_sampler_time_hbm_kernels(plan, N.ptr(flush), flush.numel(), 20, N.ptr(ms_h), st))
        left[0] = NOISE_STEPS - 2
        hbm = peaks["hbm_gbs"]
        name0 = C.create_string_buffer(64)
        by0 = C.c_double()
        N.check(lib.drs_plan_launch_info(plan, 0, name0, 64, None, C.byref(by0), None, None))
        upd_bytes = 16.0 * x.numel()     # read x, eps, z; write x (fp32)
        hbm_kernels = {
            "conv0_kernel": {"bytes": by0.value, "ms": float(ms_h[0]), "gbs": by0.value / (float(ms_h[0]) * 1e6),
                             "frac": by0.value / (float(ms_h[0]) * 1e6) / hbm,
                             "what": "x fp32 NCHW + condition feature in, h0 bf16 NHWC out (cfg 2)"},
            "ddpm_update_kernel": {"bytes": upd_bytes, "ms": float(ms_h[1]), "gbs": upd_bytes / (float(ms_h[1]) * 1e6),
                                   "frac": upd_bytes / (float(ms_h[1]) * 1e6) / hbm,
                                   "what": "16 B per element: x, eps, z in, x out (cfg 2), bookkeeping tail included"},
            "peak_gbs": hbm, "timing": "CUDA events around single launches, L2 flushed (512 MB read) before each, mean of 20"}
        del flush
        roofline = {"bound": "tensor", "achieved": achieved, "unit": "TFLOP/s",
                    "peak": peaks["tflops_burst"] if at_max_clock else peaks["tflops"],
                    "frac": frac_b if at_max_clock else frac_s,
                    "peak_kind": ("burst (bf16_tflops): SM clock at max, no power cap during the timed region"
                                  if at_max_clock else "sustained (bf16_tflops_sustained)"),
                    "frac_sustained": frac_s, "frac_burst": frac_b,
                    "peak_sustained": peaks["tflops"], "peak_burst": peaks["tflops_burst"],
                    "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peaks["source"],
                    "kernel": f"conv_gemm kernels (the {n_conv} tcgen05 implicit-GEMM launches of one UNet evaluation)",
                    "launches_per_eval": n_conv,
                    "algorithmic_gflop_per_eval": tot_f / 1e9, "algorithmic_gflop_per_launch": tot_f / 1e9 / n_conv,
                    "kernel_ms_per_eval": chain_ms, "kernel_ms_per_launch": chain_ms / n_conv,
                    "conv0_ms": float(ms2[0]),
                    "kernel_share_of_step": chain_ms / (ms / K),
                    "per_launch_event_sum_ms": tot_ms,
                    "hbm_kernels": hbm_kernels}
        if args.layers:
            with open(args.layers, "w") as f:
                json.dump(layer_rows, f, indent=1)

    # ---- end-to-end leg through the public API with host buffers ---------------------------------------------
    e2e = None
    d_e2e = D.Diffusion("cosine", model, "/nonexistent", noise_steps=K + 1, device=str(dev), magnification_factor=MAG,
                        image_size=S, Degradation_type="DownBlur")
    lr_host = lr.clone().pin_memory()
    xT_host = T.np_randn(5 + rank, BATCH, 3, S, S).pin_memory()
    out_host = torch.empty((BATCH, 3, S, S), dtype=torch.float32).pin_memory()
    d_e2e.sample(BATCH, model, lr_host, input_channels=3, x_T=xT_host)          # warm-up (plans, graphs, tables)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = d_e2e.sample(BATCH, model, lr_host, input_channels=3, x_T=xT_host)
    out_host.copy_(res, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = t.item()
    e2e = {"value": world * BATCH * K / dt, "unit": UNIT,
           "h2d_bytes_per_step": (lr_host.numel() + xT_host.numel()) * 4 / K,
           "d2h_bytes_per_step": out_host.numel() * 4 / K,
           "call": f"Diffusion.sample(n={BATCH}, model, lr_img) with noise_steps={K + 1} ({K} UNet evaluations), pinned host in/out",
           "seconds": dt}

    # ---- cfg 5: aggregation sampling of an LR 2048 x 2048 scene, patch list sharded over the ranks (strong scaling) ----
    aggregation = None
    if not args.no_aggregation:
        aggregation = aggregation_leg(model, dev, 50, rank, world, roofline)

    # ---- CPU baseline beside it (rank 0, N = 1 only) -----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cb = cpu_reference_arm(steps=3, warmup=1, budget_s=25.0)
        cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "image_size": S, "noise_steps": NOISE_STEPS,
                           "l2": "activations of one step (560 MB bf16) exceed the 126 MB L2; no explicit flush",
                           "parallelism": f"batch-sharded x{world}, no per-step collective"},
                "steps_per_sec": K / (ms * 1e-3), "sr_images_per_sec": value / (NOISE_STEPS - 1),
                "clocks": clock_info, "e2e": e2e, "gpu_launches": K * launches_per_step,
                "launches_per_step": launches_per_step, "roofline": roofline, "aggregation": aggregation,
                "cpu_baseline": cpu}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layers", default=None, help="write the per-launch table (JSON) to this path")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-aggregation", action="store_true", help="skip the cfg-5 aggregation (strong-scaling) leg")
    ap.add_argument("--scene", type=int, default=2048, help="LR scene side of the aggregation leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    SCENE_ARG[0] = args.scene
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
