#!/usr/bin/env python
"""Headline benchmark: lip-synced frames/s (256^2, 20 DDIM steps, CFG 1.5) + UNet step ms (BASELINE.json).

A "step" is ONE 16-frame segment through the whole hot path: 20 x {13-channel concat + CFG duplicate, UNet forward,
CFG combine + DDIM update}, VAE decode, paste-back (BASELINE.json configs[1], the single-GPU configuration the metric
is quoted on).  With N > 1 every rank processes its own K segments (segments of a clip are independent: weak
scaling) and the decoded frames are gathered to rank 0 with NCCL inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W]         # this framework (CUDA kernels behind the C-ABI)
    python bench.py --impl reference ...                         # CPU reference arm: oracle port on the host cores
    python bench.py --clip-segments 8                            # BASELINE configs[2]: ONE 8-segment clip sharded over
    python bench.py --clip-segments 94 --ddim-steps 50 --guidance 2.0   # configs[3]   the ranks (strong scaling)

Clip mode goes through the product's sharded entry LipsyncPipeline.run_clip (shard_segments -> run_segments ->
gather_frames): the clip's total work is fixed, a step is one pass over the whole clip.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FRAMES, HEIGHT, WIDTH, DDIM_STEPS, GUIDANCE = 16, 256, 256, 20, 1.5
METRIC = "lip-synced frames/s (256x256, 20 DDIM steps, CFG 1.5)"
WORKLOAD = "configs[1]: one 16-frame 256x256 segment per step: 20x(UNet3D fwd, CFG batch 2, 13x16x32x32 in, " \
           "Whisper embeds 16x50x384) + DDIM + VAE decode + paste-back, random-init stage2 weights"


def make_config(world: int, steps: int, spb: int = 1, clip: int = 0, ddim: int = DDIM_STEPS, g: float = GUIDANCE) -> dict:
    wl = WORKLOAD
    if clip:
        which = {(8, 20): "configs[2]: 5 s clip (125 frames -> 8 segments), 20 steps",
                 (94, 50): "configs[3]: 60 s clip (1500 frames -> 94 segments), 50 DDIM steps, CFG 2.0"}.get(
                     (clip, ddim), f"clip of {clip} segments, {ddim} DDIM steps")
        wl = (f"{which}: the clip's {clip} 16-frame 256x256 segments sharded over the ranks "
              "(LipsyncPipeline.run_clip), decoded frames gathered on rank 0 as fp16 over NCCL")
    return {"workload": wl, "segments_per_gpu": steps if not clip else None, "clip_segments": clip or None,
            "segments_per_batch": spb, "frames_per_segment": FRAMES, "ddim_steps": ddim,
            "guidance_scale": g, "parallelism": f"segments x{world}",
            "operands": "fp16 tensor-core operands, fp32 accumulate (bf16 cannot meet rel-L2 1e-2, DESIGN.md)",
            "l2": "no explicit flush: 2.5 GB of fp16 weights stream through the 126 MB L2 every UNet forward"}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return (p["bf16_tflops_sustained"], p["hbm_gbs"], "measured (MEASURED_PEAKS.json, sustained cuBLAS bf16)",
                p["bf16_tflops"])
    except Exception:
        return 1400.0, 6650.0, "fallback (B200_PROFILING.md)", 1590.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,power.limit")

    def __init__(self, index: int):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.index), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, lim, reasons = [], [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
            try:  # board power: under sw_power_cap the SM clock is whatever fits the limit for the current kernel mix
                pw.append(float(c[3]))
                lim.append(float(c[9]))
            except (ValueError, IndexError):
                pass
        os.unlink(self.f.name)
        sm.sort()
        pw.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w": pw[len(pw) // 2] if pw else None,
                "power_limit_w": max(lim) if lim else None}


def _cfg_batch(seg):
    x = torch.cat([seg["latents"]] * 2)
    x = torch.cat([x, torch.cat([seg["mask_latents"]] * 2), torch.cat([seg["masked_image_latents"]] * 2),
                   torch.cat([seg["ref_latents"]] * 2)], dim=1)
    a = seg["audio_embeds"][None]
    return x, torch.cat([torch.zeros_like(a), a])


def cpu_sample(threads: int, ddim: int = DDIM_STEPS):
    """bounded CPU sample for the `cpu_baseline` object of the default run: ONE CFG-batched fp32 UNet forward of the
    oracle port (oracle/unet_ref.py) and the oracle VAE decode of 2 of the 16 frames; the segment time is assembled as
    ddim * t_fwd + 8 * t_vae2 (the --impl reference arm times a WHOLE segment instead)"""
    from latentsync_b200 import synthetic as syn
    from latentsync_b200.spec import STAGE2_UNET_CONFIG
    from oracle import pipeline_ref as P
    from oracle.unet_ref import unet_forward

    torch.set_num_threads(threads)
    sd = syn.unet_state_dict(STAGE2_UNET_CONFIG, seed=0)
    vsd = syn.vae_decoder_state_dict(seed=0)
    seg = syn.segment_inputs(11, 0, FRAMES, HEIGHT, WIDTH)
    x, a = _cfg_batch(seg)
    t0 = time.perf_counter()
    unet_forward(sd, STAGE2_UNET_CONFIG, x, 951, a)
    t_fwd = time.perf_counter() - t0
    z = (seg["latents"] / 0.18215)[0].permute(1, 0, 2, 3)[:2].contiguous()
    t0 = time.perf_counter()
    P.vae_decode(vsd, z)
    t_vae2 = time.perf_counter() - t0
    t_seg = ddim * t_fwd + (FRAMES / 2) * t_vae2
    return {"value": FRAMES / t_seg, "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": f"1 CFG-batched fp32 UNet forward of the oracle port ({t_fwd:.1f} s) + oracle VAE decode of 2 of "
                      f"the 16 frames ({t_vae2:.1f} s); segment = {ddim} x forward + 8 x that decode = {t_seg:.0f} s"}


def run_reference(args):
    """--impl reference: the reference's CPU path for the same workload.  The reference is Python and cannot travel to
    the GPU box (no diffusers / decord there, and /root/reference is absent), so this times the oracle port, which
    oracle/make_golden.py pins to the reference's own modules at rel-L2 2e-6, on ALL host cores.  One step = ONE WHOLE
    segment, measured: 20 x (CFG concat, fp32 UNet forward, CFG combine, DDIM update) + VAE decode + paste-back
    (oracle/pipeline_ref.py, the restatement of lipsync_pipeline.py:500-575).  A segment is ~2-3 minutes of CPU work,
    so exactly one is timed whatever --steps asks for, after one untimed warm-up forward; `steps` / `warmup` in the
    line are what was actually run.  If one forward is so slow that a segment would not end within ~5 minutes, the loop
    is cut after `max_steps` forwards and the remainder is scaled (the line says so)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from latentsync_b200 import synthetic as syn
    from latentsync_b200.spec import STAGE2_UNET_CONFIG
    from oracle import pipeline_ref as P
    from oracle.unet_ref import unet_forward

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    ddim, g = args.ddim_steps, args.guidance
    sd = syn.unet_state_dict(STAGE2_UNET_CONFIG, seed=0)
    vsd = syn.vae_decoder_state_dict(seed=0)
    seg = syn.segment_inputs(100, 0, FRAMES, HEIGHT, WIDTH)
    x, a = _cfg_batch(seg)
    t0 = time.perf_counter()
    unet_forward(sd, STAGE2_UNET_CONFIG, x, 951, a)  # warm-up (thread pool, oneDNN primitives, page faults)
    t_warm = time.perf_counter() - t0
    budget = float(os.environ.get("LS_BENCH_REF_BUDGET_S", "280"))  # seconds allowed for the timed loop
    max_steps = None
    if t_warm * ddim > budget:
        max_steps = max(1, int(budget / t_warm))
    t0 = time.perf_counter()
    lat = P.denoise_segment(lambda s_, t_, a_: unet_forward(sd, STAGE2_UNET_CONFIG, s_, t_, a_), seg, steps=ddim,
                            guidance=g, max_steps=max_steps)
    t_loop = time.perf_counter() - t0
    t1 = time.perf_counter()
    frames = P.decode_and_paste(vsd, lat, seg)
    t_dec = time.perf_counter() - t1
    assert frames.shape == (FRAMES, 3, HEIGHT, WIDTH)
    done = ddim if max_steps is None else max_steps
    t_seg = t_loop * (ddim / done) + t_dec
    fps = FRAMES / t_seg
    sample = (f"ONE whole segment on {threads} host threads, fp32 oracle port: {done} of {ddim} CFG-batched UNet forwards "
              f"timed ({t_loop:.1f} s{'' if max_steps is None else ', scaled to ' + str(ddim)}) + VAE decode + paste-back "
              f"({t_dec:.1f} s); 1 untimed warm-up forward ({t_warm:.1f} s)")
    emit(({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": 1, "warmup": 1, "ms_per_step": t_seg * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args.gpus, 1, ddim=ddim, g=g), "unet_step_ms": t_loop / done * 1e3,
        "vae_decode_paste_ms": t_dec * 1e3,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def gpu_eager_baseline(dev, ddim: int = DDIM_STEPS, g: float = GUIDANCE):
    """BASELINE.md §1/§3's GPU bar: the reference's network as stock PyTorch eager on THIS B200, fp16 (the reference's
    own GPU dtype, scripts/inference.py), cuDNN convolutions, cuBLAS linears, F.scaled_dot_product_attention - the
    oracle port of the reference modules moved to the GPU with `.half()`, none of this repo's kernels.  Informational:
    one forward (median of 3 after a warm-up) and one whole segment (20-step CFG loop + VAE decode + paste)."""
    from latentsync_b200 import synthetic as syn
    from latentsync_b200.spec import STAGE2_UNET_CONFIG
    from oracle import pipeline_ref as P
    from oracle import unet_ref as U

    sd = {k: v.to(dev, torch.float16) for k, v in syn.unet_state_dict(STAGE2_UNET_CONFIG, seed=0).items()}
    vsd = {k: v.to(dev, torch.float16) for k, v in syn.vae_decoder_state_dict(seed=0).items()}
    seg = {k: v.to(dev, torch.float16) for k, v in syn.segment_inputs(100, 0, FRAMES, HEIGHT, WIDTH).items()}
    x, a = _cfg_batch(seg)
    old = U.USE_SDPA
    U.USE_SDPA = True
    try:
        fn = lambda s_, t_, a_: U.unet_forward(sd, STAGE2_UNET_CONFIG, s_, t_, a_)  # noqa: E731
        fn(x, 951, a)
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(x, 951, a)
            e1.record()
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1))
        fwd_ms = sorted(ts)[1]
        P.decode_and_paste(vsd, seg["latents"], seg)  # warm-up of the decoder's cuDNN algorithm choices (untimed)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lat = P.denoise_segment(fn, seg, steps=ddim, guidance=g)
        frames = P.decode_and_paste(vsd, lat, seg)
        e1.record()
        torch.cuda.synchronize(dev)
        seg_ms = e0.elapsed_time(e1)
        ok = bool(torch.isfinite(frames.float()).all().item())
    finally:
        U.USE_SDPA = old
    del sd, vsd
    torch.cuda.empty_cache()
    return {"value": FRAMES / (seg_ms * 1e-3), "unit": "frames/s", "unet_step_ms": fwd_ms, "ms_per_step": seg_ms,
            "finite": ok, "dtype": "f16",
            "what": "oracle port of the reference modules as PyTorch eager on this GPU (.half().cuda(): cuDNN conv, "
                    "cuBLAS linear, F.scaled_dot_product_attention); one segment = 20-step CFG loop + VAE decode + paste; "
                    "device-resident inputs, CUDA events"}


_RESULT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints "NCCL version ..." to stdout at
    communicator creation: a 4-GPU run's first stdout line was that, not JSON), so fd 1 is pointed at stderr for the
    whole run and the result line goes to the saved descriptor."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj) -> None:
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, line)


def _traffic_record(n_gemm: int):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the GEMM launches of ONE UNet forward, from the ncu
    launch list of the SAME plan that is timed here (tools/profile_unet.py -> tools/join_launches.py write
    profiles/unet_gemm_traffic.json).  Returned only when that capture has the timed plan's GEMM launch count."""
    try:
        with open(os.path.join(ROOT, "profiles", "unet_gemm_traffic.json")) as f:
            rec = json.load(f)
        if int(rec["gemm_launches"]) == int(n_gemm):
            return float(rec["dram_bytes"]), rec.get("source", "profiles/unet_gemm_traffic.json")
    except Exception:
        pass
    return None, "no ncu capture of this plan (launch count differs or file absent)"


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the informational legs (from_pixels, restore, gpu_eager_baseline, cpu_baseline)")
    ap.add_argument("--segments-per-batch", type=int, default=1,
                    help="advance this many segments of the clip together as one UNet batch (throughput mode; the "
                         "default 1 is BASELINE.json configs[1], one segment at a time like the reference's loop)")
    ap.add_argument("--clip-segments", type=int, default=0,
                    help="strong-scaling mode: ONE clip of this many segments sharded over the ranks through "
                         "LipsyncPipeline.run_clip (8 = BASELINE configs[2], 94 with --ddim-steps 50 --guidance 2.0 = "
                         "configs[3]); a step is one pass over the whole clip")
    ap.add_argument("--ddim-steps", type=int, default=DDIM_STEPS)
    ap.add_argument("--guidance", type=float, default=GUIDANCE)
    ap.add_argument("--profile-kernels", action="store_true", help="print the per-kernel-kind time table to stderr")
    args = ap.parse_args()
    clip = max(0, args.clip_segments)
    if args.steps is None:
        args.steps = 1 if clip else 5
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    ddim, guidance = args.ddim_steps, args.guidance

    import torch.distributed as dist

    from latentsync_b200 import _lib
    from latentsync_b200 import synthetic as syn
    from latentsync_b200.pipeline import LipsyncPipeline, shard_segments
    from latentsync_b200.scheduler import DDIMScheduler
    from latentsync_b200.spec import STAGE2_UNET_CONFIG
    from latentsync_b200.unet import UNet3DConditionModel
    from latentsync_b200.vae import AutoencoderKLDecoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"

    # ---- model with synthetic stage2 weights, plans captured into CUDA graphs
    cfg = STAGE2_UNET_CONFIG
    unet = UNet3DConditionModel.from_config(cfg)
    unet.load_state_dict(syn.unet_state_dict(cfg, seed=0))
    unet = unet.to(dev).eval()
    # decoder + encoder halves: the encoder is only used by the informational `from_pixels` leg below
    vae = AutoencoderKLDecoder({**syn.vae_decoder_state_dict(seed=0), **syn.vae_encoder_state_dict(seed=0)}, device=dev)
    pipe = LipsyncPipeline(vae, None, unet, DDIMScheduler()).to(dev)
    h, w = HEIGHT // 8, WIDTH // 8
    do_cfg = guidance > 1.0
    short = do_cfg and pipe.cfg_null_audio_shortcut
    # the plan the pipeline runs (null-audio shortcut of the CFG batch, engine.UNetEngine.plan); FLOP counts stay those of
    # the full plan (algorithmic work)
    nb = 2 if do_cfg else 1
    uplan = unet.plan(nb, FRAMES, h, w, 50, uncond_zero=short, same_sample=short and pipe.cfg_shared_prefix)
    uplan_full = unet.plan(nb, FRAMES, h, w, 50, capture=False)
    vplan = vae.plan(FRAMES, h, w)
    spb = max(1, args.segments_per_batch)
    frame_bytes16 = FRAMES * 3 * HEIGHT * WIDTH * 2  # one segment's decoded frames as fp16 (the gathered / D2H payload)

    def pin(seg):
        return {k: v.pin_memory() for k, v in seg.items()}

    if clip:
        # ---------------- strong scaling: ONE clip, `clip` segments, contiguous shards (run_clip)
        mine = list(shard_segments(clip, rank, world))
        nuniq = min(len(mine), 4)  # distinct synthetic segments per rank (inputs repeat beyond that: timing only)
        host = {i: pin(syn.segment_inputs(100, i, FRAMES, HEIGHT, WIDTH)) for i in mine[:nuniq]}
        host_of = {i: host[mine[k % nuniq]] for k, i in enumerate(mine)} if mine else {}
        resident_u = {i: {k: v.to(dev) for k, v in s_.items()} for i, s_ in host.items()}
        resident_of = {i: resident_u[mine[k % nuniq]] for k, i in enumerate(mine)} if mine else {}
        h2d = sum(v.numel() * v.element_size() for v in next(iter(host.values())).values()) * clip if mine else 0
        clip_host = torch.empty(clip * FRAMES, 3, HEIGHT, WIDTH, dtype=torch.float16).pin_memory() if rank == 0 else None
        d2h = clip * frame_bytes16
        warm_n = min(clip, world)  # warm-up clip: one segment per rank (plans captured, NCCL connections up)

        def run_pass(e2e: bool, nseg: int):
            """one pass over a clip of `nseg` segments (the real clip, or the `world`-segment warm-up clip whose
            segment r is rank r's first segment)"""
            full = nseg == clip

            def src(i):
                seg = (host_of if e2e else resident_of)[i if full else mine[0]]
                return {k: v.to(dev, non_blocking=True) for k, v in seg.items()} if e2e else seg

            out = pipe.run_clip(src, num_segments=nseg, num_inference_steps=ddim, guidance_scale=guidance,
                                segments_per_batch=spb)
            if e2e and rank == 0 and out is not None:
                clip_host[: out.shape[0]].copy_(out, non_blocking=True)

        def timed(e2e: bool):
            for _ in range(args.warmup):
                run_pass(e2e, warm_n)
            if world > 1:  # first use of the full-shard message size (NCCL, caching allocator): paid here, untimed
                pipe.gather_frames(torch.empty(len(mine) * FRAMES, 3, HEIGHT, WIDTH, dtype=torch.float16, device=dev),
                                   dst=0)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                run_pass(e2e, clip)
            b.record()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ms = torch.tensor([a.elapsed_time(b)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return ms.item()

        frames_total = FRAMES * clip * args.steps
        launches_per_step = (len(mine) * (ddim * (uplan.launches - len(uplan.hoisted) + 3) + 5 + vplan.launches + 2)) if rank == 0 else 0
        scaling = "strong"
    else:
        # ---------------- configs[1] (default): K segments per rank, weak scaling
        nseg = args.steps + args.warmup
        host = [pin(syn.segment_inputs(100 + rank, s_, FRAMES, HEIGHT, WIDTH)) for s_ in range(min(nseg, 4))]
        resident = [{k: v.to(dev) for k, v in s_.items()} for s_ in host]
        h2d = sum(v.numel() * v.element_size() for v in host[0].values()) * world
        out_host = torch.empty(world * args.steps * FRAMES, 3, HEIGHT, WIDTH, dtype=torch.float16).pin_memory() \
            if rank == 0 else None
        d2h = frame_bytes16 * world

        def run(seg_list, e2e: bool):
            """`len(seg_list)` steps (= segments) on this rank, advanced `spb` at a time; the ranks' decoded frames are
            gathered on rank 0 as fp16 (one gather of the whole block) and, end to end, copied to pinned host memory"""
            if e2e:
                seg_list = [{k: v.to(dev, non_blocking=True) for k, v in s_.items()} for s_ in seg_list]
            frames = torch.cat([f.to(torch.float16)
                                for f in pipe.run_segments(seg_list, ddim, guidance, segments_per_batch=spb)])
            if world > 1:
                frames = pipe.gather_frames(frames, dst=0)
            if e2e and rank == 0:
                out_host[: frames.shape[0]].copy_(frames, non_blocking=True)

        def timed(e2e: bool):
            segs = host if e2e else resident
            run([segs[i % len(segs)] for i in range(args.warmup)], e2e)
            if world > 1:
                # the timed pass gathers K segments per rank where the warm-up gathered W: first use of a message size
                # costs NCCL / the caching allocator ~100 ms once (tools/gather_probe.py), so it is paid here, untimed
                pipe.gather_frames(torch.empty(args.steps * FRAMES, 3, HEIGHT, WIDTH, dtype=torch.float16, device=dev),
                                   dst=0)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            run([segs[(args.warmup + i) % len(segs)] for i in range(args.steps)], e2e)
            b.record()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ms = torch.tensor([a.elapsed_time(b)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return ms.item()

        frames_total = FRAMES * args.steps * world
        # per step: the graph's launches + concat13 + tproj row copy + cfg_ddim; per segment: audio K/V GEMM + 4 time-path launches
        launches_per_step = ddim * (uplan.launches - len(uplan.hoisted) + 3) + 5 + vplan.launches + 2
        scaling = "weak"

    sampler = ClockSampler(local)
    if rank == 0 and os.environ.get("LS_BENCH_NO_SAMPLER", "0") != "1":
        sampler.start()
    _lib.reset_launch_count()
    ms_total = timed(e2e=False)
    counted = _lib.launch_count()  # C-ABI launches issued by THIS process during warm-up + timed region (graphs: at capture)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(e2e=True)
    value = frames_total / (ms_total * 1e-3)
    e2e_value = frames_total / (ms_e2e * 1e-3)

    # ---- UNet step alone (graph replay) and the per-kernel table (eager, one event pair per launch)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        uplan.replay()
    a.record()
    for _ in range(DDIM_STEPS):
        uplan.replay()
    b.record()
    torch.cuda.synchronize()
    unet_ms = a.elapsed_time(b) / DDIM_STEPS
    peak_tf, peak_bw, peak_src, peak_burst = measured_peaks()
    roof = None
    if rank == 0:
        uplan.run_timed()
        table = uplan.run_timed()
        vtable = vplan.run_timed()
        n_gemm_all, gemm_ms_eager = table["gemm"]
        # executed GEMM FLOPs / launches of the per-step graph (the once-per-segment audio K/V GEMM is not in it)
        gemm_flops = uplan.flops("gemm", in_graph_only=True)
        n_gemm = uplan.count("gemm", in_graph_only=True)
        # dominant kernel (gemm_tc_kernel / gemm_tc_pair_kernel): the plan's GEMM launches replayed as their own CUDA graph,
        # timed with CUDA events on the launching stream.
        #   sustained: that graph back to back for >= 1 s (the power state of a seconds-long step) -> / sustained cuBLAS peak
        #   burst    : 5 replays after the idle gap of the table runs above                         -> / burst cuBLAS peak
        gemm_ms_burst = uplan.time_kind_in_graph("gemm", reps=5)
        gemm_ms = uplan.time_kind_in_graph("gemm", reps=max(20, int(1000.0 / max(gemm_ms_burst, 1e-3))))
        kind_ms = {k: uplan.time_kind_in_graph(k, reps=20) for k in ("attention", "groupnorm", "layernorm")}
        # in-step estimate: the whole captured forward minus the same forward without its GEMM launches
        nongemm_ms = uplan.time_kind_in_graph(None, reps=20, exclude="gemm")
        gemm_ms_instep = max(unet_ms - nongemm_ms, 1e-6)
        gemm_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12
        seg_flops = ddim * uplan_full.flops(algorithmic=True) + vplan.flops(algorithmic=True)
        traffic, traffic_src = _traffic_record(n_gemm_all)  # the capture is one eager pass over ALL the plan's launches
        if args.profile_kernels:
            tot = sum(ms for _, ms in table.values())
            for k, (n, ms) in sorted(table.items(), key=lambda kv: -kv[1][1]):
                print(f"unet  {k:20s} {n:4d} launches {ms:8.3f} ms {100 * ms / tot:5.1f} %", file=sys.stderr)
            tot = sum(ms for _, ms in vtable.values())
            for k, (n, ms) in sorted(vtable.items(), key=lambda kv: -kv[1][1]):
                print(f"vae   {k:20s} {n:4d} launches {ms:8.3f} ms {100 * ms / tot:5.1f} %", file=sys.stderr)
            for name, pl in (("unet", uplan), ("vae", vplan)):
                for kind in ("gemm", "attention", "groupnorm", "layernorm"):
                    for d, n, ms, tf in pl.shape_table(kind):
                        print(f"{name} {kind} x{n:3d} {ms:8.3f} ms {tf:7.1f} TF/s  {d}", file=sys.stderr)
        roof = {"bound": "tensor", "achieved": gemm_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": gemm_tf / peak_tf, "traffic": traffic,
                "traffic_unit": "DRAM bytes of the GEMM launches of one UNet forward (ncu dram__bytes_read+write.sum)",
                "traffic_source": traffic_src,
                "algorithmic_bytes_per_unet_forward": uplan.bytes("gemm", in_graph_only=True),
                "peak_source": peak_src,
                "kernel": "gemm_tc_kernel / gemm_tc_pair_kernel (tcgen05 GEMM / implicit-GEMM conv, cta_group::1 / ::2)",
                "launches_per_unet_forward": n_gemm,
                "flops_per_unet_forward": gemm_flops,
                "flops_note": "EXECUTED FLOPs of the timed plan's GEMM launches (the sub-pixel phases of the upsample + 3x3 "
                              "convolutions execute 4/9 of the reference's multiply-adds; whole_step counts the reference's)",
                "avg_launch_us": 1e3 * gemm_ms / n_gemm,
                "how": "CUDA events around a CUDA graph holding only the plan's GEMM launches, in plan order, replayed "
                       "back to back for >= 1 s (sustained power state); peak = sustained cuBLAS bf16",
                "share_of_unet_forward": gemm_ms / unet_ms,
                "isolated_burst": {"achieved": gemm_flops / (gemm_ms_burst * 1e-3) / 1e12, "peak": peak_burst,
                                   "frac": gemm_flops / (gemm_ms_burst * 1e-3) / 1e12 / peak_burst,
                                   "how": "the same graph, 5 replays after an idle gap; peak = burst cuBLAS bf16"},
                "in_step_estimate": {"achieved": gemm_flops / (gemm_ms_instep * 1e-3) / 1e12, "peak": peak_tf,
                                     "frac": gemm_flops / (gemm_ms_instep * 1e-3) / 1e12 / peak_tf,
                                     "gemm_ms": gemm_ms_instep, "non_gemm_graph_ms": nongemm_ms,
                                     "how": "captured UNet forward (graph replay) minus the same forward captured "
                                            "without its GEMM launches; launch gaps land on the GEMM side"},
                "achieved_eager_events": uplan.flops("gemm") / (gemm_ms_eager * 1e-3) / 1e12,
                "share_of_unet_forward_eager_events": gemm_ms_eager / sum(ms for _, ms in table.values()),
                "other_kinds_ms_in_graph": kind_ms,
                "whole_step": {"flops_per_segment": seg_flops,
                               "achieved": seg_flops * (frames_total / FRAMES) / (ms_total * 1e-3) / 1e12 / world,
                               "frac": seg_flops * (frames_total / FRAMES) / (ms_total * 1e-3) / 1e12 / world / peak_tf,
                               "unit": "TFLOP/s per GPU (algorithmic FLOPs of the full plan)"}}

    # ---- informational: one segment from PIXELS: pinned host frames -> H2D -> VAE encode of the
    # masked and reference frames (2 x 16 images) + nearest mask resize -> 20-step loop -> decode + paste -> D2H
    extras = rank == 0 and world == 1 and not clip and not args.no_extras
    from_pixels = None
    if extras:
        seg0 = host[0]
        px_out_host = torch.empty(FRAMES, 3, HEIGHT, WIDTH, dtype=torch.float32).pin_memory()
        px_host = {"ref": seg0["ref_pixel_values"], "masks": seg0["masks"]}
        gen = torch.Generator(device=dev).manual_seed(1234)

        def pixels_step():
            ref_px = px_host["ref"].to(dev, non_blocking=True)
            masks = px_host["masks"].to(dev, non_blocking=True)
            masked_px = ref_px * masks
            mask_lat, masked_lat = pipe.prepare_mask_latents(masks, masked_px, HEIGHT, WIDTH, torch.float32, dev, gen,
                                                             False)
            ref_lat = pipe.prepare_image_latents(ref_px, dev, torch.float32, gen, False)
            seg = {"latents": resident[0]["latents"], "audio_embeds": resident[0]["audio_embeds"],
                   "mask_latents": mask_lat, "masked_image_latents": masked_lat, "ref_latents": ref_lat,
                   "ref_pixel_values": ref_px, "masks": masks}
            px_out_host.copy_(pipe.run_segments([seg], DDIM_STEPS, GUIDANCE)[0], non_blocking=True)

        for _ in range(2):
            pixels_step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            pixels_step()
        b.record()
        torch.cuda.synchronize()
        ms_px = a.elapsed_time(b) / args.steps
        eplan = vae.encode_plan(FRAMES, HEIGHT, WIDTH)
        a.record()
        for _ in range(4):
            eplan.replay()
        b.record()
        torch.cuda.synchronize()
        from_pixels = {"value": FRAMES / (ms_px * 1e-3), "unit": "frames/s", "ms_per_step": ms_px,
                       "vae_encode_ms_per_16_frames": a.elapsed_time(b) / 4, "encode_flops_per_16_frames": eplan.flops(),
                       "what": "encode(masked) + encode(ref) + 20-step loop + decode + paste, pixels from pinned host "
                               "memory, frames back to host"}

    # ---- informational: inverse-affine paste-back of 16 faces into 1080p frames, device-resident
    # and from / to pinned host memory; the reference's per-frame OpenCV path timed on the host cores beside it
    restore = None
    if extras:
        import numpy as np
        from latentsync_b200.restore import FaceRestorer
        from oracle import restore_ref as RR  # cpu_baseline of this leg only (the reference's OpenCV call sequence)

        cases = [syn.restore_case(500 + i, 1080, 1920, (0.45, 0.6), (500.0, 700.0)) for i in range(FRAMES)]
        frames_h = torch.from_numpy(np.stack([c[0] for c in cases])).pin_memory()
        faces_h = torch.from_numpy(np.stack([c[1] for c in cases])).pin_memory()
        mats = [c[2] for c in cases]
        fr_out_h = torch.empty_like(frames_h).pin_memory()
        restorer = FaceRestorer(dev)
        frames_d, faces_d = frames_h.to(dev), faces_h.to(dev)
        out_d = torch.empty_like(frames_d)

        def timed_ms(fn, n=5):
            fn(); fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n

        ms_dev = timed_ms(lambda: restorer.restore_imgs(frames_d, faces_d, mats, out=out_d))
        ms_host = timed_ms(lambda: fr_out_h.copy_(restorer.restore_imgs(frames_h, faces_h, mats), non_blocking=True))
        _, rois, _, _ = restorer.plan(mats, 210, 280, 1920, 1080)
        roi_px = int(((rois[:, 2] - rois[:, 0]) * (rois[:, 3] - rois[:, 1])).sum())
        # algorithmic bytes: frame read + write (u8 x 3), the face once, per ROI pixel the e2 / soft-mask planes once each
        # way (2 x 2 x 4 B) and the ROI frame bytes once more each way for the blend
        alg_bytes = 2 * frames_d.numel() + faces_d.numel() + roi_px * (16 + 6)
        t0 = time.perf_counter()
        ncpu = 4
        for i in range(ncpu):
            RR.restore_img_cv2(*cases[i])
        t_cpu = (time.perf_counter() - t0) / ncpu
        restore = {"value": FRAMES / (ms_dev * 1e-3), "unit": "frames/s", "ms_per_16_frames": ms_dev,
                   "e2e": {"value": FRAMES / (ms_host * 1e-3), "unit": "frames/s",
                           "h2d_bytes_per_step": frames_h.numel() + faces_h.numel(),
                           "d2h_bytes_per_step": fr_out_h.numel()},
                   "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms_dev * 1e-3) / 1e9, "peak": peak_bw,
                                "unit": "GB/s", "frac": alg_bytes / (ms_dev * 1e-3) / 1e9 / peak_bw,
                                "traffic": None, "algorithmic_bytes": alg_bytes, "roi_pixels": roi_px},
                   "cpu_baseline": {"value": 1.0 / t_cpu, "unit": "frames/s", "cores": os.cpu_count() or 1,
                                    "kind": "reference",
                                    "sample": f"{ncpu} frames through the reference's OpenCV call sequence "
                                              "(oracle.restore_ref.restore_img_cv2 = affine_transform.py:85-115), "
                                              "cv2's own thread pool"},
                   "what": "AlignRestore.restore_img for 16 x 1080p frames, ~400 px faces, byte-exact vs OpenCV "
                           "(tests/test_restore_gpu.py); 7 launches + 1 D2D copy per call"}

    # ---- informational: the reference network as stock PyTorch eager on this GPU (BASELINE.md's GPU bar)
    eager = None
    if extras:
        try:
            eager = gpu_eager_baseline(dev)
        except Exception as e:  # an informational leg must never take the headline down
            eager = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    cpu = None
    if rank == 0 and world == 1 and not clip and not args.no_cpu_baseline and not args.no_extras:
        cpu = cpu_sample(os.cpu_count() or 1)

    if rank == 0:
        emit(({
            "metric": METRIC if (ddim, guidance) == (DDIM_STEPS, GUIDANCE) else
            f"lip-synced frames/s (256x256, {ddim} DDIM steps, CFG {guidance})",
            "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": make_config(world, args.steps, spb, clip, ddim, guidance),
            "unet_step_ms": unet_ms,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "what": "pinned-host segment inputs -> H2D -> run_segments / run_clip -> fp16 frames gathered on "
                            "rank 0 -> D2H into pinned host memory, all inside the timed region"},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_note": "kernels launched on rank 0 inside the timed region: per segment ddim x (UNet graph "
                                 "launches + concat13 + time-row copy + cfg_ddim) + audio K/V GEMM + 4 time-path "
                                 "launches (once per segment: they do not depend on the latents) + VAE plan + 2",
            "c_abi_calls_counted": counted,
            "clocks": clocks,
            "roofline": roof,
            "from_pixels": from_pixels,
            "restore": restore,
            "gpu_eager_baseline": eager,
            "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
