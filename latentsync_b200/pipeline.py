"""Drop-in for the reference's `LipsyncPipeline` (latentsync/pipelines/lipsync_pipeline.py:46-604).

Only the hot span is re-implemented: the segment loop (:500-575) - CFG duplicate + 13-channel concat,
UNet forward, CFG combine, DDIM update (x num_inference_steps), VAE decode and paste-back - as `denoise_segment`,
`decode_latents`, `paste_surrounding_pixels_back` and `run_segments`.  `__call__` keeps the reference signature and
delegates the untouched pre/post stages (video decode, face alignment, Whisper features, affine restore, ffmpeg mux)
to the reference's own `latentsync.*` utilities, which must be importable for that entry point.

Multi-GPU (new capability): segments are independent, so `run_segments` takes a contiguous shard of
the clip per rank and `gather_frames` collects decoded frames on rank 0 with one NCCL gather over NVLink.
"""
from __future__ import annotations

import math
import os
from typing import Callable, Dict, List, Optional, Sequence, Union

import torch

from . import _lib as L


def shard_segments(num_segments: int, rank: int, world_size: int) -> range:
    """contiguous block partition of segment indices (sizes differ by at most one)"""
    base, rem = divmod(num_segments, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


class LipsyncPipeline:
    cfg_null_audio_shortcut = True  # see UNetEngine.plan(uncond_zero=...); False keeps the full-batch cross-attention
    cfg_shared_prefix = True        # see UNetEngine.plan(same_sample=...): needs the shortcut above; not bitwise (GroupNorm
                                    # partial sums are chunked differently), same tolerance against the reference
    # The whole denoising loop of a segment (lipsync_pipeline.py:537-568: per step concat13, time-embedding row, UNet, CFG +
    # DDIM) as ONE CUDA graph per (plan, schedule, guidance): the same launches in the same order - bit-identical results -
    # without the four small eager launches between the per-step graph replays, which cost ~0.2 ms per step because they
    # keep the next replay from being queued behind the running one (tools/graph_alt.py).  Paths that look at every step
    # (`trace`, `teacher_latents`, `callback`) keep the per-step loop.  LS_LOOP_GRAPH=0 / loop_graph = False disables.
    loop_graph = os.environ.get("LS_LOOP_GRAPH", "1") != "0"

    def __init__(self, vae, audio_encoder, denoising_unet, scheduler):
        self.vae = vae
        self.audio_encoder = audio_encoder
        self.denoising_unet = denoising_unet
        self.scheduler = scheduler
        boc = getattr(vae.config, "block_out_channels", (128, 256, 512, 512))
        self.vae_scale_factor = 2 ** (len(boc) - 1)
        self.device = torch.device("cpu")
        self.image_processor = None
        self._progress_bar_config = {}
        self._loop_graphs: Dict[tuple, "_LoopGraph"] = {}

    def to(self, device):
        self.device = torch.device(device)
        self.denoising_unet = self.denoising_unet.to(device)
        if hasattr(self.vae, "to"):
            self.vae = self.vae.to(device)
        return self

    @property
    def _execution_device(self):
        return self.device

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config.update(kwargs)

    def check_inputs(self, height, width, callback_steps):
        """lipsync_pipeline.py:168-180"""
        assert height == width, "Height and width must be equal"
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if (callback_steps is None) or (not isinstance(callback_steps, int) or callback_steps <= 0):
            raise ValueError(f"`callback_steps` has to be a positive integer but is {callback_steps} of type"
                             f" {type(callback_steps)}.")

    def prepare_latents(self, batch_size, num_frames, num_channels_latents, height, width, dtype, device, generator):
        """lipsync_pipeline.py:182-196: ONE (b,4,1,h,w) draw repeated over all frames of the clip"""
        shape = (batch_size, num_channels_latents, 1, height // self.vae_scale_factor, width // self.vae_scale_factor)
        device = torch.device(device)
        rand_device = "cpu" if device.type == "mps" else device
        latents = torch.randn(shape, generator=generator, device=rand_device, dtype=dtype).to(device)
        latents = latents.repeat(1, 1, num_frames, 1, 1)
        return latents * self.scheduler.init_noise_sigma

    # ------------------------------------------------------------------- pixel-space pre / post
    @torch.no_grad()
    def prepare_masks_and_masked_images(self, images, mask_image: torch.Tensor):
        """ImageProcessor.prepare_masks_and_masked_images for mask="fix_mask", affine_transform=False and faces that are
        already at the working resolution (image_processor.py:145-165; the reference's per-frame Python loop):
        uint8 (f,H,W,3) / (f,3,H,W) frames -> (pixel_values, masked_pixel_values, masks), fp32 on the device.
        `mask_image`: (3,H,W) or (1,H,W) as returned by load_fixed_mask (1 = keep)."""
        dev = self.device
        if not isinstance(images, torch.Tensor):
            images = torch.from_numpy(images)
        img = images.to(dev).contiguous()
        if img.dtype != torch.uint8:
            raise TypeError("expected uint8 frames")
        hwc = img.shape[-1] == 3 and img.shape[1] != 3
        n = img.shape[0]
        H, W = (img.shape[1], img.shape[2]) if hwc else (img.shape[2], img.shape[3])
        m = mask_image.to(dev, torch.float32).contiguous()
        if tuple(m.shape[-2:]) != (H, W):
            raise ValueError(f"frames are {H}x{W} but the mask is {tuple(m.shape[-2:])}: resize the faces first")
        pixel = torch.empty(n, 3, H, W, dtype=torch.float32, device=dev)
        masked = torch.empty_like(pixel)
        L.preprocess_u8(img, m, pixel, masked)
        return pixel, masked, m[0:1].unsqueeze(0).expand(n, 1, H, W).contiguous()

    @torch.no_grad()
    def faces_to_uint8(self, faces: torch.Tensor, height: int, width: int) -> torch.Tensor:
        """front half of restore_video for faces that share one box size (lipsync_pipeline.py:350-355): anti-aliased
        bilinear resize to (height, width), [-1, 1] -> uint8, "c h w -> h w c".  Returns (f, height, width, 3) uint8 on
        the device (one D2H copy for the whole clip instead of one per frame)."""
        x = faces.to(self.device, torch.float32).contiguous()
        out = torch.empty(x.shape[0], height, width, 3, dtype=torch.uint8, device=x.device)
        L.resize_aa_u8(x, height, width, out)
        return out

    # ------------------------------------------------------------------- VAE encode of the conditioning frames
    def _encode_scaled(self, images: torch.Tensor, device, dtype, generator) -> torch.Tensor:
        """(vae.encode(x).latent_dist.sample(generator) - shift_factor) * scaling_factor -> (f, 4, h, w) fp32"""
        sf, sh = self.vae.config.scaling_factor, self.vae.config.shift_factor
        dist = self.vae.encode(images.to(device=device, dtype=dtype)).latent_dist
        if hasattr(dist, "sample_scaled"):
            # the draw diffusers' randn_tensor makes: on the generator's device when one is given, in `dtype`
            rdev = generator.device if generator is not None else torch.device(device)
            noise = torch.randn((dist.n, dist.c, dist.h, dist.w), generator=generator, device=rdev, dtype=dtype)
            return dist.sample_scaled(noise.to(device), float(sh), float(sf))
        return ((dist.sample(generator=generator) - sh) * sf).float()

    @torch.no_grad()
    def prepare_mask_latents(self, mask, masked_image, height, width, dtype, device, generator,
                             do_classifier_free_guidance):
        """lipsync_pipeline.py:284-311: nearest-resize the pixel mask to the latent grid, VAE-encode the masked frames
        (sampled, shifted, scaled), "f c h w -> 1 c f h w", duplicate for CFG."""
        mask = torch.nn.functional.interpolate(mask, size=(height // self.vae_scale_factor,
                                                           width // self.vae_scale_factor))
        masked_image_latents = self._encode_scaled(masked_image, device, dtype, generator).to(dtype)
        mask = mask.to(device=device, dtype=dtype).permute(1, 0, 2, 3).unsqueeze(0)
        masked_image_latents = masked_image_latents.permute(1, 0, 2, 3).unsqueeze(0)
        if do_classifier_free_guidance:
            mask = torch.cat([mask] * 2)
            masked_image_latents = torch.cat([masked_image_latents] * 2)
        return mask, masked_image_latents

    @torch.no_grad()
    def prepare_image_latents(self, images, device, dtype, generator, do_classifier_free_guidance):
        """lipsync_pipeline.py:313-320"""
        image_latents = self._encode_scaled(images, device, dtype, generator).to(dtype)
        image_latents = image_latents.permute(1, 0, 2, 3).unsqueeze(0)
        return torch.cat([image_latents] * 2) if do_classifier_free_guidance else image_latents

    # ------------------------------------------------------------------------------------------------ hot loop
    @torch.no_grad()
    def denoise_segment(self, latents: torch.Tensor, audio_embeds: Optional[torch.Tensor], mask_latents: torch.Tensor,
                        masked_image_latents: torch.Tensor, ref_latents: torch.Tensor, num_inference_steps: int = 20,
                        guidance_scale: float = 1.5, callback: Optional[Callable] = None, callback_steps: int = 1,
                        trace: Optional[Dict[str, list]] = None, teacher_latents: Optional[Sequence] = None):
        """Denoising loop of one segment (lipsync_pipeline.py:537-568).

        latents (1,4,f,h,w); audio_embeds (f,S,D) = the conditional half only (the uncond half is zeros, :503-507);
        mask_latents (1|2,1,f,h,w), masked_image_latents / ref_latents (1|2,4,f,h,w) - if the CFG-duplicated (2,...)
        form of prepare_mask_latents (:308-311) is passed, the first half is used (both halves are identical).
        Returns the final latents (1,4,f,h,w) fp32.  `trace`, if a dict, receives per-step "noise_pred" (guided) and
        "latents" tensors; `teacher_latents[j]`, if given, replaces the loop state before step j (per-step parity)."""
        unet, sch = self.denoising_unet, self.scheduler
        dev = unet.device
        do_cfg = guidance_scale > 1.0
        nb = 2 if do_cfg else 1
        lat = latents.to(dev, torch.float32).contiguous().clone()
        assert lat.shape[0] == 1 and lat.shape[1] == 4, "one segment per call (batch_size is 1, lipsync_pipeline.py:392)"
        _, _, F, h, w = lat.shape
        mask = mask_latents[:1].to(dev, torch.float32).contiguous()
        masked = masked_image_latents[:1].to(dev, torch.float32).contiguous()
        ref = ref_latents[:1].to(dev, torch.float32).contiguous()
        S = 0
        if unet.add_audio_layer:
            assert audio_embeds is not None
            S = audio_embeds.shape[-2]
        # the unconditional half of the CFG batch is built right here as zeros: the plan may use that (engine.plan)
        # ... and both halves get the same latents / mask / reference channels and timestep from ls_concat13 below
        short = do_cfg and S > 0 and self.cfg_null_audio_shortcut
        plan = unet.plan(nb, F, h, w, S, uncond_zero=short, same_sample=short and self.cfg_shared_prefix)
        if S:
            a = audio_embeds.to(dev, torch.float16).reshape(F * S, -1)
            buf = plan.audio_in.tensor()
            buf.zero_()  # uncond half: null audio embeds (:505-507)
            buf[(nb - 1) * F * S:, : a.shape[1]].copy_(a)
        sch.set_timesteps(num_inference_steps)
        timesteps = sch._host_timesteps
        lib = L.lib()
        eps_dbg = torch.empty_like(lat) if trace is not None else None
        t_view = plan.t_in.tensor().view(-1)
        # what does not depend on the latents runs ONCE per segment instead of once per step: the audio K/V projection of
        # all 16 cross-attention layers (a function of the audio embeddings) and the time-embedding path of all timesteps
        # (a function of t: one batched pass, row j = step j)
        if S:
            plan.run_hoisted(plan.kv_ops)
        if self.loop_graph and trace is None and teacher_latents is None and callback is None:
            lg = self._loop_graph(plan, timesteps, float(guidance_scale), nb, F, h, w)
            lg.lat.copy_(lat)
            lg.mask.copy_(mask)
            lg.masked.copy_(masked)
            lg.ref.copy_(ref)
            lg.graph.replay()
            return lg.lat.clone()
        ttab = unet.engine().time_table(timesteps)
        tproj = plan.tproj.tensor()
        for j, t in enumerate(timesteps):
            if teacher_latents is not None:
                lat.copy_(teacher_latents[j].to(dev, torch.float32))
            st = torch.cuda.current_stream().cuda_stream
            # cat([latents]*2) ; cat([x, mask, masked, ref], dim=1)  -> channels-last fp16 UNet input (:542-549)
            L._check(lib.ls_concat13(lat.data_ptr(), mask.data_ptr(), masked.data_ptr(), ref.data_ptr(), nb, F, h * w,
                                     plan.x_in.ptr, st), "ls_concat13")
            t_view.fill_(float(t))
            tproj.copy_(ttab[j].expand_as(tproj))
            plan.replay()  # noise_pred (:552-554)
            a_t, a_p = sch.step_coefficients(t)
            # CFG combine (:557-559) + DDIM step (:562) in one pass, latents updated in place
            L._check(lib.ls_cfg_ddim_step(plan.eps_out.ptr, plan.eps_out.cols, nb, F, h * w, float(guidance_scale), a_t,
                                          a_p, lat.data_ptr(), eps_dbg.data_ptr() if eps_dbg is not None else None, st),
                     "ls_cfg_ddim_step")
            if trace is not None:
                trace.setdefault("noise_pred", []).append(eps_dbg.clone())
                trace.setdefault("latents", []).append(lat.clone())
            if callback is not None and j % callback_steps == 0:
                callback(j, t, lat)
        return lat

    def _loop_graph(self, plan, timesteps, guidance_scale: float, nb: int, F: int, h: int, w: int) -> "_LoopGraph":
        """the captured loop for this (plan, schedule, guidance); built on first use (one eager pass, then the capture)"""
        key = (id(plan), tuple(float(t) for t in timesteps), guidance_scale)
        lg = self._loop_graphs.get(key)
        if lg is None:
            lg = self._loop_graphs[key] = _LoopGraph(self, plan, timesteps, guidance_scale, nb, F, h, w)
        return lg

    @torch.no_grad()
    def decode_latents(self, latents: torch.Tensor) -> torch.Tensor:
        """lipsync_pipeline.py:145-149: z / scaling_factor + shift -> "(b f) c h w" -> vae.decode(...).sample"""
        sf, sh = self.vae.config.scaling_factor, self.vae.config.shift_factor
        b, c, f, h, w = latents.shape
        if hasattr(self.vae, "plan") and sh == 0.0 and b == 1 and latents.dtype == torch.float32 and latents.is_cuda:
            plan = self.vae.plan(f, h, w)
            st = torch.cuda.current_stream().cuda_stream
            lib = L.lib()
            lat = latents.contiguous()
            L._check(lib.ls_ncfhw_to_cl(lat.data_ptr(), 1, c, f, h * w, plan.z_in.cols, 1.0 / sf, plan.z_in.ptr, st),
                     "ls_ncfhw_to_cl")
            plan.replay()
            H, W = plan.out_h, plan.out_w
            out = torch.empty(f, 3, H, W, dtype=torch.float32, device=latents.device)
            L._check(lib.ls_cl_to_ncfhw(plan.dec_out.ptr, plan.ld_out, f, 3, 1, H * W, out.data_ptr(), st),
                     "ls_cl_to_ncfhw")
            return out
        z = latents / sf + sh
        z = z.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
        vdt = getattr(self.vae, "dtype", None)
        if isinstance(vdt, torch.dtype) and not hasattr(self.vae, "plan"):
            z = z.to(vdt)
        return self.vae.decode(z).sample

    @torch.no_grad()
    def decode_and_paste(self, latents: torch.Tensor, ref_pixel_values: torch.Tensor, masks: torch.Tensor):
        """decode_latents + paste_surrounding_pixels_back(decoded, ref, 1 - masks) (:571-574) without materialising the
        decoded frames in NCHW: out = decoded * (1 - m) + ref * m, m = `masks` (1 = keep the original pixel)."""
        sf = self.vae.config.scaling_factor
        sh = float(getattr(self.vae.config, "shift_factor", 0.0) or 0.0)
        _, c, f, h, w = latents.shape
        if not hasattr(self.vae, "plan") or sh != 0.0:
            # a foreign VAE (e.g. diffusers' AutoencoderKL, as scripts/inference.py builds it) or a non-zero shift:
            # the reference's two calls, lipsync_pipeline.py:571-574
            decoded = self.decode_latents(latents)
            return self.paste_surrounding_pixels_back(decoded, ref_pixel_values, 1 - masks, latents.device,
                                                      torch.float32)
        plan = self.vae.plan(f, h, w)
        st = torch.cuda.current_stream().cuda_stream
        lib = L.lib()
        lat = latents.to(torch.float32).contiguous()
        dev = lat.device
        L._check(lib.ls_ncfhw_to_cl(lat.data_ptr(), 1, c, f, h * w, plan.z_in.cols, 1.0 / sf, plan.z_in.ptr, st),
                 "ls_ncfhw_to_cl")
        plan.replay()
        H, W = plan.out_h, plan.out_w
        ref = ref_pixel_values.to(dev, torch.float32).contiguous()
        m = masks.to(dev, torch.float32).contiguous()
        out = torch.empty(f, 3, H, W, dtype=torch.float32, device=dev)
        L._check(lib.ls_paste_back(plan.dec_out.ptr, plan.ld_out, ref.data_ptr(), m.data_ptr(), f, H * W,
                                   out.data_ptr(), st), "ls_paste_back")
        return out

    @staticmethod
    def paste_surrounding_pixels_back(decoded_latents, pixel_values, masks, device, weight_dtype):
        """lipsync_pipeline.py:328-333 (called with `1 - masks` at :572-574): decoded * masks + pixel * (1 - masks)"""
        d = decoded_latents.to(device, torch.float32).contiguous()
        n, c, H, W = d.shape
        ref = pixel_values.to(device, torch.float32).contiguous()
        keep = (1 - masks.to(device, torch.float32)).contiguous()  # kernel takes m with out = d*(1-m) + ref*m
        out = torch.empty_like(d)
        d_cl = d.permute(0, 2, 3, 1).reshape(n * H * W, c).contiguous()
        L.paste_back(d_cl, c, ref, keep, n, H * W, out)
        return out.to(weight_dtype)

    @torch.no_grad()
    def denoise_segments(self, segments: Sequence[Dict[str, torch.Tensor]], num_inference_steps: int = 20,
                         guidance_scale: float = 1.5) -> List[torch.Tensor]:
        """Denoising loops of SEVERAL independent segments advanced together as one UNet batch (batch order
        [seg0 uncond, seg0 cond, seg1 uncond, ...]).  Same arithmetic per segment as `denoise_segment` - segments never
        interact (GroupNorm statistics are per batch element, attention per frame / per pixel) - but the small 8x8 / 4x4
        levels fill more SMs and every launch is amortised over more rows.  Returns the final latents per segment."""
        if len(segments) == 1:
            g = segments[0]
            return [self.denoise_segment(g["latents"], g.get("audio_embeds"), g["mask_latents"],
                                         g["masked_image_latents"], g["ref_latents"], num_inference_steps,
                                         guidance_scale)]
        unet, sch = self.denoising_unet, self.scheduler
        dev = unet.device
        do_cfg = guidance_scale > 1.0
        nb = 2 if do_cfg else 1
        n = len(segments)
        lats = [g["latents"].to(dev, torch.float32).contiguous().clone() for g in segments]
        _, _, F, h, w = lats[0].shape
        masks = [g["mask_latents"][:1].to(dev, torch.float32).contiguous() for g in segments]
        maskeds = [g["masked_image_latents"][:1].to(dev, torch.float32).contiguous() for g in segments]
        refs = [g["ref_latents"][:1].to(dev, torch.float32).contiguous() for g in segments]
        S = segments[0]["audio_embeds"].shape[-2] if unet.add_audio_layer else 0
        plan = unet.plan(nb * n, F, h, w, S)
        rows_seg = nb * F * h * w
        if S:
            buf = plan.audio_in.tensor()
            buf.zero_()
            for i, g in enumerate(segments):
                a = g["audio_embeds"].to(dev, torch.float16).reshape(F * S, -1)
                r0 = (i * nb + (nb - 1)) * F * S
                buf[r0:r0 + F * S, : a.shape[1]].copy_(a)
        sch.set_timesteps(num_inference_steps)
        lib = L.lib()
        t_view = plan.t_in.tensor().view(-1)
        x_stride = rows_seg * plan.x_in.cols * 2  # bytes per segment in the fp16 UNet input
        e_stride = rows_seg * plan.eps_out.cols * 4
        if S:
            plan.run_hoisted(plan.kv_ops)
        ttab = unet.engine().time_table(sch._host_timesteps)
        tproj = plan.tproj.tensor()
        for j, t in enumerate(sch._host_timesteps):
            st = torch.cuda.current_stream().cuda_stream
            for i in range(n):
                L._check(lib.ls_concat13(lats[i].data_ptr(), masks[i].data_ptr(), maskeds[i].data_ptr(),
                                         refs[i].data_ptr(), nb, F, h * w, plan.x_in.ptr + i * x_stride, st),
                         "ls_concat13")
            t_view.fill_(float(t))
            tproj.copy_(ttab[j].expand_as(tproj))
            plan.replay()
            a_t, a_p = sch.step_coefficients(t)
            for i in range(n):
                L._check(lib.ls_cfg_ddim_step(plan.eps_out.ptr + i * e_stride, plan.eps_out.cols, nb, F, h * w,
                                              float(guidance_scale), a_t, a_p, lats[i].data_ptr(), None, st),
                         "ls_cfg_ddim_step")
        return lats

    @torch.no_grad()
    def run_segments(self, segments: Sequence[Dict[str, torch.Tensor]], num_inference_steps: int = 20,
                     guidance_scale: float = 1.5, segments_per_batch: int = 1) -> List[torch.Tensor]:
        """HOT LOOP 1 (lipsync_pipeline.py:500-575) over already-prepared segment inputs (keys as produced by
        synthetic.segment_inputs / the reference's prepare_* helpers).  Returns one (f,3,H,W) fp32 tensor each.
        `segments_per_batch` > 1 advances that many segments of the clip together (see denoise_segments)."""
        frames = []
        for i0 in range(0, len(segments), max(1, segments_per_batch)):
            group = segments[i0:i0 + max(1, segments_per_batch)]
            for seg, lat in zip(group, self.denoise_segments(group, num_inference_steps, guidance_scale)):
                frames.append(self.decode_and_paste(lat, seg["ref_pixel_values"], seg["masks"]))
        return frames

    @torch.no_grad()
    def run_clip(self, segments, num_segments: Optional[int] = None, num_inference_steps: int = 20,
                 guidance_scale: float = 1.5, segments_per_batch: int = 1, out_dtype: torch.dtype = torch.float16,
                 gather: bool = True, dst: int = 0) -> Optional[torch.Tensor]:
        """The segment loop of ONE clip (lipsync_pipeline.py:500-575), sharded over the ranks of the default process
        group: rank r denoises + decodes the contiguous block `shard_segments(n, r, world)` of the
        clip's segments - they are independent, every rank holds a full weight replica, there is no data-path
        collective - and the decoded frames are gathered on `dst` (NCCL point-to-point over NVLink) in `out_dtype`
        (the reference keeps decoded frames in weight_dtype = fp16, :571-574; 6.3 MB per segment).

        `segments`: the clip's prepared segment inputs (sequence indexed by segment), or a callable `i -> segment` so
        that a rank only materialises its own shard (`num_segments` is then required).  Without an initialised process
        group this is the single-GPU loop.  Returns (n_frames, 3, H, W) on `dst`, None on the other ranks
        (`gather=False`: the local shard's frames on every rank)."""
        import torch.distributed as dist

        world, rank = 1, 0
        if dist.is_available() and dist.is_initialized():
            world, rank = dist.get_world_size(), dist.get_rank()
        n = num_segments if num_segments is not None else len(segments)
        mine = shard_segments(n, rank, world)
        get = segments if callable(segments) else segments.__getitem__
        local = [get(i) for i in mine]
        frames = self.run_segments(local, num_inference_steps, guidance_scale, segments_per_batch)
        if frames:
            out = torch.cat([f.to(out_dtype) for f in frames])
        else:  # more ranks than segments: an empty shard still takes part in the gather
            out = torch.empty((0, 3, 0, 0), dtype=out_dtype, device=self.device)
        if not gather or world == 1:
            return out
        return self.gather_frames(out, dst=dst)

    @staticmethod
    def gather_frames(frames: torch.Tensor, seg_counts: Optional[Sequence[int]] = None,
                      dst: int = 0) -> Optional[torch.Tensor]:
        """Gather of the per-rank decoded frames to `dst` over the default process group (NCCL on the GPUs: point-to-
        point sends over NVLink; gloo in the CPU tests).  Ranks hold contiguous blocks of the clip in rank order;
        frames: (n_local_frames, 3, H, W) in any dtype (fp16 / uint8 keep the payload at 6.3 / 3.1 MB per segment).
        The ranks first agree on every rank's frame count and on the frame geometry (a rank may hold an empty shard, and
        the clip's last segment may be shorter), so `seg_counts` is only kept for callers of the first version.
        Returns the whole clip on `dst`, None elsewhere."""
        import torch.distributed as dist

        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
            return frames
        world, rank = dist.get_world_size(), dist.get_rank()
        n_local = int(frames.shape[0])
        meta = torch.tensor([n_local] + (list(frames.shape[1:]) if n_local else [0, 0, 0]), dtype=torch.int64,
                            device=frames.device)
        metas = [torch.empty_like(meta) for _ in range(world)]
        dist.all_gather(metas, meta)
        counts = [int(m[0].item()) for m in metas]
        shape = tuple(int(v) for v in torch.stack(metas)[:, 1:].max(dim=0).values.tolist())
        if rank == dst:
            outs = [torch.empty((counts[r],) + shape, dtype=frames.dtype, device=frames.device) for r in range(world)]
            reqs = []
            for r in range(world):
                if r == dst:
                    if counts[r]:
                        outs[r].copy_(frames)
                elif counts[r] > 0:
                    reqs.append(dist.irecv(outs[r], src=r))
            for q in reqs:
                q.wait()
            return torch.cat(outs, dim=0)
        if n_local > 0:
            dist.send(frames.contiguous(), dst=dst)
        return None

    # ------------------------------------------------------------------------------------------- full entry point
    @torch.no_grad()
    def __call__(
        self,
        video_path: str,
        audio_path: str,
        video_out_path: str,
        video_mask_path: str = None,
        num_frames: int = 16,
        video_fps: int = 25,
        audio_sample_rate: int = 16000,
        height: Optional[int] = None,
        width: Optional[int] = None,
        num_inference_steps: int = 20,
        guidance_scale: float = 1.5,
        weight_dtype: Optional[torch.dtype] = torch.float16,
        eta: float = 0.0,
        mask: str = "fix_mask",
        mask_image_path: str = "latentsync/utils/mask.png",
        generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
        callback: Optional[Callable[[int, int, torch.FloatTensor], None]] = None,
        callback_steps: Optional[int] = 1,
        data_path: Optional[str] = None,
        start_from_backwards: Optional[bool] = False,
        force_video_length: Optional[bool] = False,
        use_darken: Optional[bool] = False,
        brightness_factor: Optional[float] = 1.0,
        **kwargs,
    ):
        """Same signature and side effect (an mp4 at `video_out_path`) as lipsync_pipeline.py:361-604.  Stages outside
        the hot path run the reference's own code, imported from its `latentsync` package."""
        try:  # untouched stages
            import numpy as np
            import soundfile as sf
            from latentsync.pipelines.affine_transform_video import affine_transform_video
            from latentsync.utils.image_processor import ImageProcessor, load_fixed_mask
            from latentsync.utils.repeat import (pad_whisper_chunks, pad_whisper_chunks_end,
                                                 pad_whisper_chunks_to_target, repeat_to_length, truncate_to_length)
            from latentsync.utils.util import read_audio, read_video, write_video
        except ImportError as e:
            raise ImportError("LipsyncPipeline.__call__ needs the reference's untouched pre/post-processing package "
                              "`latentsync` (face alignment, Whisper front end, ffmpeg I/O) on sys.path; the "
                              f"accelerated span is available as run_segments(): {e}") from e
        import os
        import shutil
        import subprocess

        if eta != 0.0:
            raise NotImplementedError("eta > 0 is not supported (the reference always passes eta=0.0)")
        is_train = self.denoising_unet.training
        self.denoising_unet.eval()
        device = self._execution_device
        mask_image = load_fixed_mask(height, mask_image_path)
        self.image_processor = ImageProcessor(height, mask=mask, device="cuda", mask_image=mask_image)
        if data_path:
            loaded = torch.load(data_path)
            faces, boxes, affine_matrices = loaded["faces"], loaded["boxes"], loaded["affine_matrices"]
            original_video_frames = read_video(video_path, use_decord=False)
        else:
            faces, original_video_frames, boxes, affine_matrices = affine_transform_video(self.image_processor,
                                                                                          video_path)
        height = height or self.denoising_unet.config.sample_size * self.vae_scale_factor
        width = width or self.denoising_unet.config.sample_size * self.vae_scale_factor
        self.check_inputs(height, width, callback_steps)
        self.video_fps = video_fps
        do_cfg = guidance_scale > 1.0
        self.scheduler.set_timesteps(num_inference_steps, device=device)
        audio_samples = read_audio(audio_path)
        whisper_feature = self.audio_encoder.audio2feat(audio_path)
        whisper_chunks = self.audio_encoder.feature2chunks(feature_array=whisper_feature, fps=video_fps)
        padding_duration = 0
        if not force_video_length:
            if start_from_backwards:
                whisper_chunks, audio_samples, padding_duration, _ = pad_whisper_chunks(
                    whisper_chunks, whisper_chunks[0].shape, audio_samples, audio_sample_rate, self.video_fps)
            else:
                whisper_chunks, audio_samples, padding_duration = pad_whisper_chunks_end(
                    whisper_chunks, whisper_chunks[0].shape, audio_samples, audio_sample_rate, self.video_fps)
            if len(whisper_chunks) > len(faces):
                n = len(whisper_chunks)
                faces, boxes = repeat_to_length(faces, n), repeat_to_length(boxes, n)
                original_video_frames = repeat_to_length(original_video_frames, n)
                affine_matrices = repeat_to_length(affine_matrices, n)
        else:
            whisper_chunks, audio_samples, padding_duration = pad_whisper_chunks_to_target(
                whisper_chunks, whisper_chunks[0].shape, audio_samples, audio_sample_rate, len(faces), fps=self.video_fps)
        if len(faces) != len(whisper_chunks) and start_from_backwards:
            n = len(whisper_chunks)
            faces, boxes = truncate_to_length(faces, n), truncate_to_length(boxes, n)
            original_video_frames = truncate_to_length(original_video_frames, n)
            affine_matrices = truncate_to_length(affine_matrices, n)

        # the reference draws the shared initial noise in weight_dtype (fp16) - :489-498 - and so do we: the loop then
        # carries it in fp32, but it starts from exactly the reference's (fp16-rounded) values
        all_latents = self.prepare_latents(1, len(whisper_chunks), self.vae.config.latent_channels, height, width,
                                           weight_dtype, device, generator).float()
        # `segments_per_batch` (extra keyword, swallowed by the reference's **kwargs): advance that many consecutive
        # segments of the clip as ONE UNet batch (denoise_segments: same arithmetic per segment, +8-13 % frames/s at 2-4).
        # The per-segment preparation - and with it the order of the generator's draws - stays the reference's.
        spb = max(1, int(kwargs.get("segments_per_batch", 1)))
        if callback is not None:
            spb = 1  # the callback contract is per step of ONE segment
        # Multi-GPU: when a torch.distributed process group is up, every rank calls __call__ with the
        # same arguments; the segments of the clip are sharded in contiguous blocks (shard_segments), frames are
        # gathered on rank 0, which alone restores / writes the video.  Ranks skip the segments they do not own but
        # still make the generator draws those segments would have made, so the clip is the single-GPU clip.
        import torch.distributed as dist

        world, rank = 1, 0
        if dist.is_available() and dist.is_initialized():
            world, rank = dist.get_world_size(), dist.get_rank()
        num_inferences = math.ceil(len(whisper_chunks) / num_frames)
        mine = shard_segments(num_inferences, rank, world)
        synced = []
        pending = []  # prepared segments waiting for their batch

        def flush():
            if not pending:
                return
            if len(pending) == 1:
                g = pending[0]
                lats = [self.denoise_segment(g["latents"], g["audio_embeds"], g["mask_latents"], g["masked_image_latents"],
                                             g["ref_latents"], num_inference_steps, guidance_scale if do_cfg else 1.0,
                                             callback, callback_steps)]
            else:
                lats = self.denoise_segments(pending, num_inference_steps, guidance_scale if do_cfg else 1.0)
            for g, lat in zip(pending, lats):
                synced.append(self.decode_and_paste(lat, g["ref_px"], g["masks"]).to(weight_dtype))
            pending.clear()

        lat_hw = (height // self.vae_scale_factor, width // self.vae_scale_factor)
        for i in range(num_inferences):
            inference_faces = faces[i * num_frames:(i + 1) * num_frames]
            if i not in mine:
                # the two encode draws of a segment this rank does not own (prepare_mask_latents, prepare_image_latents)
                if generator is not None:
                    shape = (len(inference_faces), self.vae.config.latent_channels) + lat_hw
                    for _ in range(2):
                        torch.randn(shape, generator=generator, device=generator.device, dtype=weight_dtype)
                continue
            audio_embeds = None
            if self.denoising_unet.add_audio_layer:
                audio_embeds = torch.stack(whisper_chunks[i * num_frames:(i + 1) * num_frames]).to(device)
            latents = all_latents[:, :, i * num_frames:(i + 1) * num_frames]
            if mask == "fix_mask" and tuple(inference_faces.shape[-2:]) == (height, width):
                ref_px, masked_px, masks = self.prepare_masks_and_masked_images(inference_faces, mask_image)
            else:  # other mask modes need landmarks; other sizes need torchvision's uint8 resize
                ref_px, masked_px, masks = self.image_processor.prepare_masks_and_masked_images(
                    inference_faces, affine_transform=False)
            # VAE encode of the masked / reference frames (:525-535), fp32 latents for the loop
            mask_lat, masked_lat = self.prepare_mask_latents(masks, masked_px, height, width, weight_dtype, device,
                                                             generator, False)
            ref_lat = self.prepare_image_latents(ref_px, device, weight_dtype, generator, False)
            seg = {"latents": latents, "audio_embeds": audio_embeds, "mask_latents": mask_lat,
                   "masked_image_latents": masked_lat, "ref_latents": ref_lat, "ref_px": ref_px, "masks": masks}
            # only full segments of one shape share a batch (the last one of a clip may be shorter)
            if pending and (latents.shape != pending[0]["latents"].shape or audio_embeds is None):
                flush()
            pending.append(seg)
            if len(pending) >= spb:
                flush()
        flush()
        if world > 1:
            local = (torch.cat(synced) if synced else
                     torch.empty((0, 3, height, width), dtype=weight_dtype, device=device))
            gathered = self.gather_frames(local, dst=0)
            if rank != 0:  # rank 0 alone restores the frames and writes the video
                if is_train:
                    self.denoising_unet.train()
                return None
            synced = [gathered]
        self.image_processor_restore = self.image_processor.restorer
        frames = self._restore_video(torch.cat(synced), original_video_frames, boxes, affine_matrices)
        remain = int(frames.shape[0] / video_fps * audio_sample_rate)
        audio_samples = audio_samples[:remain].cpu().numpy()
        if is_train:
            self.denoising_unet.train()
        temp_dir = "temp"
        if os.path.exists(temp_dir):
            shutil.rmtree(temp_dir)
        os.makedirs(temp_dir, exist_ok=True)
        write_video(os.path.join(temp_dir, "video.mp4"), frames, fps=25, use_darken=use_darken,
                    brightness_factor=brightness_factor)
        sf.write(os.path.join(temp_dir, "audio.wav"), audio_samples, audio_sample_rate)
        v, a = os.path.join(temp_dir, "video.mp4"), os.path.join(temp_dir, "audio.wav")
        if start_from_backwards or force_video_length:
            command = f"ffmpeg -y -loglevel error -nostdin -i {v} -i {a} -c:v libx264 -c:a aac -q:v 0 -q:a 0 {video_out_path}"
        else:
            command = (f"ffmpeg -y -loglevel error -nostdin -i {v} -i {a} -c:v libx264 -c:a aac -q:v 0 -q:a 0 -t "
                       f"$(ffprobe -v error -show_entries format=duration -of default=noprint_wrappers=1:nokey=1 {v} | "
                       f"awk '{{print $1-{padding_duration}}}') {video_out_path}")
        subprocess.run(command, shell=True)

    def _restore_video(self, faces, video_frames, boxes, affine_matrices, frames_per_call: int = 32):
        """lipsync_pipeline.py:343-358 on the GPU: anti-aliased resize -> uint8 (faces_to_uint8, ls_resize_aa_u8), then
        the inverse-affine paste-back of AlignRestore.restore_img (affine_transform.py:85-115) for `frames_per_call`
        frames per launch (restore.FaceRestorer, ls_restore_faces: the same bytes as the reference's per-frame cv2
        path).  Returns the restored frames as one uint8 array (n, H, W, 3) like restore_video."""
        import numpy as np

        from .restore import FaceRestorer

        n = len(faces)
        video_frames = np.asarray(video_frames[:n])
        if getattr(self, "_face_restorer", None) is None:
            self._face_restorer = FaceRestorer(self.device)
        restorer = self._face_restorer
        sizes = [(int(b[3] - b[1]), int(b[2] - b[0])) for b in boxes[:n]]
        out = np.empty_like(video_frames)
        for hw in sorted(set(sizes)):  # frames that share a box size go together (usually all of them)
            idx = [i for i, s in enumerate(sizes) if s == hw]
            for c0 in range(0, len(idx), frames_per_call):
                chunk = idx[c0:c0 + frames_per_call]
                u8 = self.faces_to_uint8(torch.stack([faces[i] for i in chunk]), hw[0], hw[1])
                mats = [affine_matrices[i] for i in chunk]
                res = restorer.restore_imgs(video_frames[chunk], u8, mats)
                restorer.check_status()
                out[chunk] = res.cpu().numpy()
        return out



class _LoopGraph:
    """One CUDA graph holding every launch of a segment's denoising loop (LipsyncPipeline.loop_graph): static fp32 state
    (`lat` is updated in place by ls_cfg_ddim_step, `mask` / `masked` / `ref` are the conditioning channels of ls_concat13),
    the time-embedding table of the schedule, and the graph.  The audio K/V projection stays outside (once per segment,
    before the replay)."""

    def __init__(self, pipe: "LipsyncPipeline", plan, timesteps, guidance_scale: float, nb: int, F: int, h: int, w: int):
        unet, sch = pipe.denoising_unet, pipe.scheduler
        dev = unet.device
        self.plan = plan  # keeps id(plan) unique for the life of the entry
        self.lat = torch.zeros(1, 4, F, h, w, dtype=torch.float32, device=dev)
        self.mask = torch.zeros(1, 1, F, h, w, dtype=torch.float32, device=dev)
        self.masked = torch.zeros(1, 4, F, h, w, dtype=torch.float32, device=dev)
        self.ref = torch.zeros(1, 4, F, h, w, dtype=torch.float32, device=dev)
        self.ttab = unet.engine().time_table(timesteps)
        coeffs = [sch.step_coefficients(t) for t in timesteps]
        lib = L.lib()
        tproj = plan.tproj.tensor()

        def loop():
            st = torch.cuda.current_stream().cuda_stream
            for j in range(len(timesteps)):
                L._check(lib.ls_concat13(self.lat.data_ptr(), self.mask.data_ptr(), self.masked.data_ptr(),
                                         self.ref.data_ptr(), nb, F, h * w, plan.x_in.ptr, st), "ls_concat13")
                tproj.copy_(self.ttab[j].expand_as(tproj))
                plan.run(hoisted=False)
                a_t, a_p = coeffs[j]
                L._check(lib.ls_cfg_ddim_step(plan.eps_out.ptr, plan.eps_out.cols, nb, F, h * w, guidance_scale, a_t, a_p,
                                              self.lat.data_ptr(), None, st), "ls_cfg_ddim_step")

        if plan.graph is None:
            plan.capture()  # warms every launch of the plan up (function attributes, scratch) before our own capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            loop()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            loop()
