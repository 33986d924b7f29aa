#!/usr/bin/env python
"""Headline benchmark: lip-synced frames/s (256^2, 20 DDIM steps, CFG 1.5) + UNet step ms (BASELINE.json).

A "step" is ONE 16-frame segment through the whole hot path: 20 x {13-channel concat + CFG duplicate, UNet forward,
CFG combine + DDIM update}, VAE decode, paste-back (BASELINE.json configs[1], the single-GPU configuration the metric
is quoted on).  With N > 1 every rank processes its own K segments (segments of a clip are independent: weak
scaling) and the decoded frames are gathered to rank 0 with NCCL inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W]         # this framework (CUDA kernels behind the C-ABI)
    python bench.py --impl reference ...                         # CPU reference arm: oracle port on the host cores
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


# DRAM traffic of the dominant kernel (gemm_tc_kernel): sum of dram__bytes_read.sum + dram__bytes_write.sum over the 341
# GEMM launches of ONE CFG-batched UNet forward, from the ncu capture profiles/r1b_launches_unet.csv (joined table:
# profiles/r1_launch_table.txt).  Same unit of work as `achieved` (FLOPs of those 341 launches / their summed duration).
GEMM_DRAM_BYTES_PER_UNET_FORWARD = 9.4858e9


def make_config(world: int, steps: int, spb: int = 1) -> dict:
    return {"workload": WORKLOAD, "segments_per_gpu": steps, "segments_per_batch": spb, "frames_per_segment": FRAMES, "ddim_steps": DDIM_STEPS,
            "guidance_scale": GUIDANCE, "parallelism": f"segments x{world}",
            "operands": "fp16 tensor-core operands, fp32 accumulate (bf16 cannot meet rel-L2 1e-2, DESIGN.md)",
            "l2": "no explicit flush: 2.5 GB of fp16 weights stream through the 126 MB L2 every UNet forward"}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["bf16_tflops_sustained"], p["hbm_gbs"], "measured (MEASURED_PEAKS.json, sustained cuBLAS bf16)"
    except Exception:
        return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

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
        sm, mx, reasons = [], [], set()
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
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_forward_sample(threads: int):
    """the oracle port (oracle/unet_ref.py, fp32) on the host cores: ONE CFG-batched UNet forward of the 20 a segment
    needs; frames/s extrapolated as 16 / (20 * t) (VAE decode, 5.8 % of the FLOPs, left out => flatters the CPU)"""
    from latentsync_b200 import synthetic as syn
    from latentsync_b200.spec import STAGE2_UNET_CONFIG
    from oracle.unet_ref import unet_forward

    torch.set_num_threads(threads)
    sd = syn.unet_state_dict(STAGE2_UNET_CONFIG, seed=0)
    seg = syn.segment_inputs(11, 0, FRAMES, HEIGHT, WIDTH)
    x = torch.cat([seg["latents"]] * 2)
    x = torch.cat([x, torch.cat([seg["mask_latents"]] * 2), torch.cat([seg["masked_image_latents"]] * 2),
                   torch.cat([seg["ref_latents"]] * 2)], dim=1)
    a = seg["audio_embeds"][None]
    a = torch.cat([torch.zeros_like(a), a])

    def one():
        t0 = time.perf_counter()
        unet_forward(sd, STAGE2_UNET_CONFIG, x, 951, a)
        return time.perf_counter() - t0

    return one


def run_reference(args):
    """--impl reference: the reference's CPU path for the same workload.  The reference is Python and cannot travel to
    the GPU box (no diffusers / decord there, and /root/reference is absent), so this times the oracle port, which
    oracle/make_golden.py pins to the reference's own modules at rel-L2 2e-6."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    one = cpu_forward_sample(threads)
    budget = 400.0
    t_first = one()
    warm_done = 1
    times = []
    for _ in range(max(args.warmup - 1, 0)):
        if t_first * (warm_done + 1 + args.steps) > budget:
            break
        one()
        warm_done += 1
    for _ in range(args.steps):
        times.append(one())
        if sum(times) + t_first * warm_done > budget and len(times) >= 1:
            break
    t = sum(times) / len(times)
    fps = FRAMES / (DDIM_STEPS * t)
    sample = (f"each step = 1 CFG-batched UNet forward (fp32, oracle port) of the {DDIM_STEPS} per segment; "
              f"frames/s = 16 / (20 * t_fwd), VAE decode excluded; {len(times)} timed + {warm_done} warm-up samples")
    emit(({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": warm_done, "ms_per_step": DDIM_STEPS * t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args.gpus, len(times)), "unet_step_ms": t * 1e3,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


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


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--segments-per-batch", type=int, default=1,
                    help="advance this many segments of the clip together as one UNet batch (throughput mode; the "
                         "default 1 is BASELINE.json configs[1], one segment at a time like the reference's loop)")
    ap.add_argument("--profile-kernels", action="store_true", help="print the per-kernel-kind time table to stderr")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist

    from latentsync_b200 import synthetic as syn
    from latentsync_b200.pipeline import LipsyncPipeline
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
    # the plan the pipeline runs (null-audio shortcut of the CFG batch, engine.UNetEngine.plan); FLOP counts stay those of
    # the full plan (algorithmic work)
    uplan = unet.plan(2, FRAMES, h, w, 50, uncond_zero=pipe.cfg_null_audio_shortcut, same_sample=pipe.cfg_null_audio_shortcut and pipe.cfg_shared_prefix)
    uplan_full = unet.plan(2, FRAMES, h, w, 50, capture=False)
    vplan = vae.plan(FRAMES, h, w)

    # ---- synthetic segments: device-resident copies for `value`, pinned host copies for `e2e`
    nseg = args.steps + args.warmup
    host = [{k: v.pin_memory() for k, v in syn.segment_inputs(100 + rank, s, FRAMES, HEIGHT, WIDTH).items()}
            for s in range(min(nseg, 4))]
    resident = [{k: v.to(dev) for k, v in s.items()} for s in host]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    out_host = torch.empty(FRAMES, 3, HEIGHT, WIDTH, dtype=torch.float32).pin_memory()
    d2h = out_host.numel() * out_host.element_size()
    gather_list = [torch.empty(FRAMES, 3, HEIGHT, WIDTH, device=dev) for _ in range(world)] if rank == 0 else None

    spb = max(1, args.segments_per_batch)

    def run(seg_list, e2e: bool):
        """`len(seg_list)` steps (= segments), advanced `spb` at a time"""
        for frames in pipe.run_segments(seg_list, DDIM_STEPS, GUIDANCE, segments_per_batch=spb):
            if world > 1:
                dist.gather(frames, gather_list, dst=0)
            if e2e:
                out_host.copy_(frames, non_blocking=True)

    def timed(segs, e2e: bool):
        run([segs[i % len(segs)] for i in range(args.warmup)], e2e)
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

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(resident, e2e=False)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(host, e2e=True)
    frames_total = FRAMES * args.steps * world
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
    uplan.run_timed()
    table = uplan.run_timed()
    vtable = vplan.run_timed()
    peak_tf, peak_bw, peak_src = measured_peaks()
    n_gemm, gemm_ms_eager = table["gemm"]
    # dominant kernel: the 341 GEMM launches of one UNet forward replayed as their own CUDA graph (kernel time without
    # the ~3 us host gap that the per-launch event pairs of run_timed add to each of them; that figure is kept as
    # `achieved_eager_events`)
    gemm_ms = uplan.time_kind_in_graph("gemm")
    kind_ms = {k: uplan.time_kind_in_graph(k) for k in ("attention", "groupnorm", "layernorm")}
    gemm_tf = uplan.flops("gemm") / (gemm_ms * 1e-3) / 1e12  # executed GEMM FLOPs of the plan that was timed
    seg_flops = DDIM_STEPS * uplan_full.flops() + vplan.flops()
    launches_per_step = DDIM_STEPS * (uplan.launches + 2) + vplan.launches + 2 + (1 if world > 1 else 0)
    if args.profile_kernels and rank == 0:
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

    # ---- informational: one segment from PIXELS: pinned host frames -> H2D -> VAE encode of the
    # masked and reference frames (2 x 16 images) + nearest mask resize -> 20-step loop -> decode + paste -> D2H
    from_pixels = None
    if rank == 0 and world == 1:
        seg0 = host[0]
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
            out_host.copy_(pipe.run_segments([seg], DDIM_STEPS, GUIDANCE)[0], non_blocking=True)

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
    if rank == 0 and world == 1:
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

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        t = cpu_forward_sample(threads)()
        cpu = {"value": FRAMES / (DDIM_STEPS * t), "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "1 CFG-batched fp32 UNet forward (oracle port of the reference modules) of the 20 per segment, "
                         f"{t:.1f} s; frames/s = 16 / (20 * t), VAE decode excluded"}

    if rank == 0:
        emit(({
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": make_config(world, args.steps, spb),
            "unet_step_ms": unet_ms,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": gemm_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": gemm_tf / peak_tf, "traffic": GEMM_DRAM_BYTES_PER_UNET_FORWARD,
                         "traffic_unit": "bytes per UNet forward (the 341 GEMM launches of the full plan, ncu profiles/r1c_launches_unet.csv)",
                         "algorithmic_bytes_per_unet_forward": uplan.bytes("gemm"),
                         "peak_source": peak_src,
                         "kernel": "gemm_tc_kernel (tcgen05 GEMM / implicit-GEMM conv)",
                         "launches_per_unet_forward": n_gemm,
                         "flops_per_unet_forward": uplan.flops("gemm"),
                         "avg_launch_us": 1e3 * gemm_ms / n_gemm,
                         "how": "CUDA events around a CUDA graph holding only these launches, in plan order",
                         "share_of_unet_forward": gemm_ms / unet_ms,
                         "achieved_eager_events": uplan.flops("gemm") / (gemm_ms_eager * 1e-3) / 1e12,
                         "share_of_unet_forward_eager_events": gemm_ms_eager / sum(ms for _, ms in table.values()),
                         "other_kinds_ms_in_graph": kind_ms,
                         "whole_step": {"flops_per_segment": seg_flops,
                                        "achieved": seg_flops * args.steps / (ms_total * 1e-3) / 1e12 / 1.0,
                                        "frac": seg_flops * args.steps / (ms_total * 1e-3) / 1e12 / peak_tf}},
            "from_pixels": from_pixels,
            "restore": restore,
            "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
