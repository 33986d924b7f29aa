"""Parity of the CUDA hot path (through the drop-in classes and the C-ABI) against
  (a) golden vectors produced by the REFERENCE's own modules (tests/golden/*.pt, oracle/make_golden.py), and
  (b) the CPU oracle port (oracle/unet_ref.py, oracle/pipeline_ref.py) run on the same seeded inputs.

Tolerances are BASELINE.json's: per-step predicted noise rel-L2 <= 1e-2 against the reference fp32 path, final frames
PSNR >= 40 dB.  Operands are fp16 with fp32 accumulation (bf16 operands cannot meet 1e-2).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-2
WEIGHT_SEED, INPUT_SEED = 0, 11


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


def cfg_batch(seg, lat=None):
    lat = seg["latents"] if lat is None else lat
    x = torch.cat([lat] * 2)
    x = torch.cat([x, torch.cat([seg["mask_latents"]] * 2), torch.cat([seg["masked_image_latents"]] * 2),
                   torch.cat([seg["ref_latents"]] * 2)], dim=1)
    a = seg["audio_embeds"][None]
    return x, torch.cat([torch.zeros_like(a), a])


_models = {}


def get_unet(name):
    from latentsync_b200 import synthetic as syn
    from latentsync_b200.spec import STAGE2_UNET_CONFIG, TINY_UNET_CONFIG
    from latentsync_b200.unet import UNet3DConditionModel

    if name not in _models:
        cfg = TINY_UNET_CONFIG if name == "tiny" else STAGE2_UNET_CONFIG
        sd = syn.unet_state_dict(cfg, seed=WEIGHT_SEED)
        m = UNet3DConditionModel.from_config(cfg)
        m.load_state_dict(sd, strict=True)
        m = m.to("cuda").eval()
        _models[name] = (m, sd, cfg)
    return _models[name]


def get_pipe(name):
    from latentsync_b200 import synthetic as syn
    from latentsync_b200.pipeline import LipsyncPipeline
    from latentsync_b200.scheduler import DDIMScheduler
    from latentsync_b200.vae import AutoencoderKLDecoder

    key = "pipe_" + name
    if key not in _models:
        unet, _, _ = get_unet(name)
        vsd = syn.vae_decoder_state_dict(seed=WEIGHT_SEED)
        vae = AutoencoderKLDecoder(vsd, device="cuda")
        _models[key] = (LipsyncPipeline(vae, None, unet, DDIMScheduler()).to("cuda"), vsd)
    return _models[key]


def test_unet_tiny_forward_vs_reference_golden_and_port(monkeypatch):
    """one CFG-batched forward, quarter-width config: vs the reference's output and, block by block, vs the port"""
    from latentsync_b200 import synthetic as syn
    from oracle.unet_ref import unet_forward

    monkeypatch.setenv("LS_DEBUG_TAPS", "1")
    unet, sd, cfg = get_unet("tiny")
    gold = torch.load(os.path.join(GOLDEN, "unet_tiny.pt"))
    seg = syn.segment_inputs(INPUT_SEED, 0, 16, 128, 128)
    x, a = cfg_batch(seg)
    y = unet(x.cuda(), 951, encoder_hidden_states=a.cuda()).sample
    taps = {}
    yp = unet_forward(sd, cfg, x, 951, a, taps=taps)
    plan = unet.plan(2, 16, 16, 16, 50)
    report = []
    for name in taps:
        if name in plan.taps:
            report.append((name, rel_l2(plan.tap_tensor(name), taps[name])))
    print("per-block rel-L2 vs port:", ", ".join(f"{n}={e:.2e}" for n, e in report))
    assert torch.isfinite(y).all()
    e_port, e_gold = rel_l2(y, yp), rel_l2(y, gold["noise_pred"])
    print(f"tiny forward: vs port {e_port:.3e}, vs reference golden {e_gold:.3e}")
    assert report and max(e for _, e in report) < TOL
    assert e_gold < TOL and e_port < TOL


def test_unet_stage1_variant_without_motion_modules_vs_reference_golden():
    """configs/unet/stage1.yaml: use_motion_module = false - the same drop-in class without temporal layers, against a
    golden vector produced by the reference's own modules (oracle/make_golden.py tiny_stage1)"""
    from latentsync_b200 import synthetic as syn
    from latentsync_b200.spec import TINY_UNET_CONFIG
    from latentsync_b200.unet import UNet3DConditionModel

    cfg = dict(TINY_UNET_CONFIG)
    cfg["use_motion_module"] = False
    sd = syn.unet_state_dict(cfg, seed=WEIGHT_SEED)
    assert not any("motion_modules" in k for k in sd)
    unet = UNet3DConditionModel.from_config(cfg)
    unet.load_state_dict(sd, strict=True)
    unet = unet.to("cuda").eval()
    gold = torch.load(os.path.join(GOLDEN, "unet_tiny_stage1.pt"))
    seg = syn.segment_inputs(INPUT_SEED, 1, 16, 128, 128)
    x, a = cfg_batch(seg)
    y = unet(x.cuda(), gold["t"], encoder_hidden_states=a.cuda()).sample
    e = rel_l2(y, gold["noise_pred"])
    print(f"tiny stage1 variant (no motion modules): rel-L2 vs reference golden {e:.3e}")
    assert e < TOL


def test_unet_forward_signature_variants():
    """timestep as int / 0-d tensor / (B,) tensor, 4-D vs 3-D encoder_hidden_states, fp16 sample, return_dict=False"""
    from latentsync_b200 import synthetic as syn

    unet, _, _ = get_unet("tiny")
    seg = syn.segment_inputs(INPUT_SEED, 0, 16, 128, 128)
    x, a = cfg_batch(seg)
    x, a = x.cuda(), a.cuda()
    y0 = unet(x, 951, encoder_hidden_states=a).sample
    y1 = unet(x, torch.tensor(951, device="cuda"), encoder_hidden_states=a.reshape(32, 50, 384)).sample
    y2 = unet(x, torch.tensor([951, 951]), encoder_hidden_states=a, return_dict=False)[0]
    assert torch.equal(y0, y1) and torch.equal(y0, y2)  # the whole forward is deterministic (no fp atomics)
    y3 = unet(x.half(), 951.0, encoder_hidden_states=a.half()).sample
    assert y3.dtype == torch.float16 and rel_l2(y3, y0) < 5e-3
    with pytest.raises(ValueError):
        unet(x[:, :12], 951, encoder_hidden_states=a)
    with pytest.raises(ValueError):
        unet(x, 951)


def test_loop_tiny_vs_reference_golden():
    """4 DDIM steps with CFG 1.5 at the tiny config, free-running: guided noise and latents after every step"""
    from latentsync_b200 import synthetic as syn

    pipe, _ = get_pipe("tiny")
    gold = torch.load(os.path.join(GOLDEN, "unet_tiny.pt"))
    seg = syn.segment_inputs(INPUT_SEED, 0, 16, 128, 128)
    trace = {}
    pipe.denoise_segment(seg["latents"], seg["audio_embeds"], seg["mask_latents"], seg["masked_image_latents"],
                         seg["ref_latents"], num_inference_steps=4, guidance_scale=1.5, trace=trace)
    for j in range(4):
        e_n = rel_l2(trace["noise_pred"][j], gold["loop4_noise_pred"][j])
        e_l = rel_l2(trace["latents"][j], gold["loop4_latents"][j])
        print(f"step {j}: noise {e_n:.3e} latents {e_l:.3e}")
        assert e_n < TOL and e_l < TOL


def test_loop_50_steps_cfg2_vs_oracle():
    """BASELINE.json configs[3] schedule: 50 DDIM steps ('leading' spacing: t = 981, 961, ..., 1), guidance 2.0, free
    running at the quarter-width config against the oracle port of the reference modules + the restated scheduler"""
    from latentsync_b200 import synthetic as syn
    from oracle import pipeline_ref as P
    from oracle.unet_ref import unet_forward

    pipe, _ = get_pipe("tiny")
    unet, sd, cfg = get_unet("tiny")
    seg = syn.segment_inputs(INPUT_SEED, 3, 16, 128, 128)
    trace, want = {}, {}
    pipe.denoise_segment(seg["latents"], seg["audio_embeds"], seg["mask_latents"], seg["masked_image_latents"],
                         seg["ref_latents"], num_inference_steps=50, guidance_scale=2.0, trace=trace)
    assert pipe.scheduler._host_timesteps[:2] == [981, 961] and pipe.scheduler._host_timesteps[-1] == 1
    P.denoise_segment(lambda x, t, a: unet_forward(sd, cfg, x, t, a), seg, steps=50, guidance=2.0, trace=want)
    errs = [rel_l2(trace["noise_pred"][j], want["noise_pred"][j]) for j in range(50)]
    e_lat = rel_l2(trace["latents"][-1], want["latents"][-1])
    print(f"50 steps, g=2.0: guided-noise rel-L2 max {max(errs):.3e} (step {errs.index(max(errs))}), final latents {e_lat:.3e}")
    assert max(errs) < TOL and e_lat < TOL


@pytest.mark.parametrize("frames,guidance", [(16, 1.0), (8, 1.5)])
def test_loop_without_cfg_and_short_segment(frames, guidance):
    """guidance_scale <= 1 disables classifier-free guidance (batch 1, lipsync_pipeline.py:446,542) and a segment may
    hold fewer than 16 frames (num_frames argument, :371); 3 free-running steps vs the oracle loop"""
    from latentsync_b200 import synthetic as syn
    from oracle import pipeline_ref as P
    from oracle.unet_ref import unet_forward

    pipe, _ = get_pipe("tiny")
    unet, sd, cfg = get_unet("tiny")
    seg = syn.segment_inputs(INPUT_SEED, 5, frames, 128, 128)
    trace, want = {}, {}
    lat = pipe.denoise_segment(seg["latents"], seg["audio_embeds"], seg["mask_latents"], seg["masked_image_latents"],
                               seg["ref_latents"], num_inference_steps=3, guidance_scale=guidance, trace=trace)
    want_lat = P.denoise_segment(lambda x, t, a: unet_forward(sd, cfg, x, t, a), seg, steps=3, guidance=guidance,
                                 trace=want)
    errs = [rel_l2(trace["noise_pred"][j], want["noise_pred"][j]) for j in range(3)]
    print(f"frames={frames} guidance={guidance}: noise rel-L2 {max(errs):.3e}, latents {rel_l2(lat, want_lat):.3e}")
    assert lat.shape == (1, 4, frames, 16, 16) and max(errs) < TOL and rel_l2(lat, want_lat) < TOL


def test_unet_512px_shape_batched_segments_vs_oracle():
    """BASELINE.json configs[4] geometry: 13 x 16 x 64 x 64 UNet input (512 x 512 pixels), temporal layers on, several
    segments in one launch (CFG batch 2 x 2 segments here) - quarter-width weights so that the CPU oracle stays in
    seconds.  Exercises the tcgen05 attention at S = 4096 / 1024 / 256 and the cluster GroupNorm at those sizes."""
    from latentsync_b200 import synthetic as syn
    from oracle.unet_ref import unet_forward

    unet, sd, cfg = get_unet("tiny")
    segs = [syn.segment_inputs(INPUT_SEED, s, 16, 512, 512) for s in range(2)]
    xs, as_ = zip(*[cfg_batch(g) for g in segs])
    x, a = torch.cat(xs), torch.cat(as_)  # [seg0 uncond, seg0 cond, seg1 uncond, seg1 cond]
    y = unet(x.cuda(), 501, encoder_hidden_states=a.cuda()).sample
    assert y.shape == (4, 4, 16, 64, 64) and torch.isfinite(y).all()
    want = unet_forward(sd, cfg, x[:2], 501, a[:2])
    e = rel_l2(y[:2], want)
    print(f"512px shape, 2 segments batched: segment 0 rel-L2 vs oracle {e:.3e}")
    assert e < TOL
    # segments never interact: segment 1 alone gives bit-identical rows
    y1 = unet(x[2:].cuda(), 501, encoder_hidden_states=a[2:].cuda()).sample
    assert rel_l2(y[2:], y1) < 2e-3


@pytest.mark.parametrize("nimg,h", [(2, 16), (16, 32)])
def test_vae_decode_and_paste_vs_oracle(nimg, h):
    """AutoencoderKL.decode restatement (oracle/pipeline_ref.py) vs the CUDA plan; PSNR in [-1,1] frame space"""
    from latentsync_b200 import synthetic as syn
    from oracle import pipeline_ref as P

    pipe, vsd = get_pipe("tiny")
    seg = syn.segment_inputs(INPUT_SEED, 0, nimg, 8 * h, 8 * h)
    lat = syn.approx_normal(5, "vae.lat", (1, 4, nimg, h, h)) * 0.18215 * 3.0
    want = P.decode_and_paste(vsd, lat, seg)
    got = pipe.decode_and_paste(lat.cuda(), seg["ref_pixel_values"], seg["masks"])
    dec = pipe.decode_latents(lat.cuda())
    want_dec = P.vae_decode(vsd, (lat / 0.18215)[0].permute(1, 0, 2, 3))
    e = rel_l2(dec, want_dec)
    p = P.psnr(got.cpu(), want)
    print(f"vae decode {nimg}x{h}: rel-L2 {e:.3e}, pasted PSNR {p:.1f} dB, |dec| {want_dec.abs().mean():.3f}")
    assert e < TOL and p >= 40.0
    # paste-back must reproduce the reference pixels exactly where the mask keeps them
    keep = seg["masks"].bool().expand_as(want)
    assert torch.equal(got.cpu()[keep], seg["ref_pixel_values"][keep])


@pytest.mark.parametrize("nimg,H", [(2, 128), (16, 256)])
def test_vae_encode_vs_oracle(nimg, H):
    """AutoencoderKL.encode restatement (oracle/pipeline_ref.py) vs the CUDA encoder plan: moments, sample with a given
    noise, and prepare_mask_latents / prepare_image_latents (lipsync_pipeline.py:284-320) through the drop-in classes"""
    from latentsync_b200 import synthetic as syn
    from latentsync_b200.pipeline import LipsyncPipeline
    from latentsync_b200.scheduler import DDIMScheduler
    from latentsync_b200.vae import AutoencoderKL
    from oracle import pipeline_ref as P

    pipe0, vsd = get_pipe("tiny")
    esd = syn.vae_encoder_state_dict(seed=0)
    vae = AutoencoderKL({**vsd, **esd}, device="cuda")
    pipe = LipsyncPipeline(vae, None, pipe0.denoising_unet, DDIMScheduler()).to("cuda")
    seg = syn.segment_inputs(INPUT_SEED, 0, nimg, H, H)
    ref_px = seg["ref_pixel_values"]
    masked_px = ref_px * seg["masks"]
    want_m = P.vae_encode_moments(esd, masked_px)
    dist = vae.encode(masked_px.to("cuda")).latent_dist
    got_m = torch.cat([dist.mean, dist._nchw(4)], dim=1)
    e = rel_l2(got_m, want_m)
    print(f"vae encode {nimg}x{H}: moments rel-L2 {e:.3e}, |mean| {want_m[:, :4].abs().mean():.3f}")
    assert e < TOL
    noise = syn.approx_normal(9, "enc.noise", (nimg, 4, H // 8, H // 8))
    z = dist.sample_scaled(noise.to("cuda"), 0.0, 0.18215)
    assert rel_l2(z, P.gaussian_sample(want_m, noise) * 0.18215) < TOL
    assert rel_l2(dist.mode(), want_m[:, :4]) < TOL
    # drop-in helpers: same generator => same draw as the reference's randn_tensor (device draw, weight dtype)
    g = torch.Generator(device="cuda").manual_seed(77)
    nz1 = torch.randn((nimg, 4, H // 8, H // 8), generator=g, device="cuda", dtype=torch.float32)
    nz2 = torch.randn((nimg, 4, H // 8, H // 8), generator=g, device="cuda", dtype=torch.float32)
    g = torch.Generator(device="cuda").manual_seed(77)
    m_lat, masked_lat = pipe.prepare_mask_latents(seg["masks"], masked_px, H, H, torch.float32, "cuda", g, True)
    ref_lat = pipe.prepare_image_latents(ref_px, "cuda", torch.float32, g, True)
    want_mask, want_masked = P.prepare_mask_latents(esd, seg["masks"], masked_px, nz1.cpu(), H, H)
    want_ref = P.prepare_image_latents(esd, ref_px, nz2.cpu())
    assert m_lat.shape == (2, 1, nimg, H // 8, H // 8) and masked_lat.shape == (2, 4, nimg, H // 8, H // 8)
    assert torch.equal(m_lat[0], m_lat[1]) and torch.equal(masked_lat[0], masked_lat[1])
    assert torch.equal(m_lat[:1].cpu(), want_mask)
    assert rel_l2(masked_lat[:1], want_masked) < TOL and rel_l2(ref_lat[:1], want_ref) < TOL


def test_pixel_pre_and_post_processing_vs_oracle():
    """prepare_masks_and_masked_images (image_processor.py:145-165) and the resize -> uint8 front
    half of restore_video (lipsync_pipeline.py:350-355) on the GPU against their torch / torchvision statements"""
    from latentsync_b200 import synthetic as syn
    from oracle import pipeline_ref as P

    pipe, _ = get_pipe("tiny")
    g = torch.Generator().manual_seed(31)
    faces = torch.randint(0, 256, (5, 256, 256, 3), generator=g, dtype=torch.uint8)
    mask3 = syn.fixed_mask(256, 256).repeat(3, 1, 1).double()
    mask3[:, 94:96] = 0.37  # the Lanczos-resized mask.png has fractional rows at the box edges
    for frames in (faces, faces.permute(0, 3, 1, 2).contiguous()):
        px, masked, masks = pipe.prepare_masks_and_masked_images(frames, mask3)
        wpx, wmasked, wmasks = P.preprocess_fixed_mask(frames, mask3)
        assert torch.equal(px.cpu(), wpx)  # same fp32 operations: bit-exact
        assert torch.allclose(masked.cpu().double(), wmasked, atol=1e-6) and masked.shape == (5, 3, 256, 256)
        assert torch.allclose(masks.cpu().double(), wmasks, atol=1e-7) and masks.shape == (5, 1, 256, 256)
    # post: decoded faces in [-1.2, 1.2] (clamp exercised) -> box-sized uint8 HWC, BYTE-exact: against the numpy
    # restatement of aten::_upsample_bilinear2d_aa's CPU kernel (oracle/resize_ref.py, pinned to torch in
    # tests/test_cpu.py) on every host, and against torch / torchvision itself where this host's PyTorch dispatches to the
    # FMA builds of that kernel (AVX2 / AVX512: every x86 server CPU) whose accumulation order the kernel reproduces
    from oracle import resize_ref as R

    dec = (torch.rand(4, 3, 256, 256, generator=g) * 2.4 - 1.2)
    fma_host = torch.backends.cpu.get_cpu_capability() in ("AVX2", "AVX512")
    for (h, w) in ((210, 280), (256, 256), (311, 287), (96, 128), (256, 300), (200, 256), (37, 41), (700, 640)):
        got = pipe.faces_to_uint8(dec, h, w).cpu()
        assert got.shape == (4, h, w, 3)
        assert np.array_equal(got.numpy(), R.restore_faces_u8(dec.numpy(), h, w)), (h, w)
        want = P.restore_faces_u8(dec, h, w)
        diff = (got.int() - want.int()).abs()
        print(f"resize 256x256 -> {h}x{w}: {int((diff > 0).sum())} bytes differ from torch on this host "
              f"({torch.backends.cpu.get_cpu_capability()})")
        assert diff.max().item() <= (0 if fma_host else 1)


@pytest.mark.parametrize("name,hw", [("tiny", 16), ("stage2", 32)])
def test_null_audio_shortcut_is_bitwise_identical(name, hw):
    """engine.UNetEngine.plan(uncond_zero=True): with all-zero audio in the first half of the CFG batch (what the pipeline
    builds, lipsync_pipeline.py:503-507) the cross-attention of those rows is exactly `to_out.bias`; the shortcut plan
    must give the SAME BITS as the full plan, and fewer attention rows"""
    unet, _, _ = get_unet(name)
    full = unet.plan(2, 16, hw, hw, 50)
    short = unet.plan(2, 16, hw, hw, 50, uncond_zero=True)
    assert short is not full and short.uncond_zero and not full.uncond_zero
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(full.x_in.tensor().shape, generator=g, device="cuda").half()
    a = torch.randn(full.audio_in.tensor().shape, generator=g, device="cuda").half()
    a[: a.shape[0] // 2] = 0  # null audio for the unconditional half
    outs = []
    for plan in (full, short):
        plan.x_in.tensor().copy_(x)
        plan.audio_in.tensor().copy_(a)
        plan.t_in.tensor().fill_(501.0)
        plan.run_hoisted()
        plan.replay()
        torch.cuda.synchronize()
        outs.append(plan.eps_out.tensor().clone())
    assert torch.isfinite(outs[0]).all() and outs[0].abs().max() > 0
    assert torch.equal(outs[0], outs[1])
    assert short.flops("attention") < full.flops("attention")
    # and the pipeline uses it: same latents from both settings of the switch
    pipe, _ = get_pipe(name)
    from latentsync_b200 import synthetic as syn
    seg = {k: v.cuda() for k, v in syn.segment_inputs(3, 0, 16, hw * 8, hw * 8).items()}
    lat = {}
    try:
        for key, (shortcut, prefix) in {"full": (False, False), "null": (True, False), "prefix": (True, True)}.items():
            pipe.cfg_null_audio_shortcut, pipe.cfg_shared_prefix = shortcut, prefix
            lat[key] = pipe.denoise_segment(seg["latents"], seg["audio_embeds"], seg["mask_latents"],
                                            seg["masked_image_latents"], seg["ref_latents"], 2, 1.5).clone()
    finally:
        pipe.cfg_null_audio_shortcut, pipe.cfg_shared_prefix = True, True
    assert torch.equal(lat["null"], lat["full"])
    rel = ((lat["prefix"] - lat["full"]).norm() / lat["full"].norm()).item()
    print(f"pipeline, shared prefix vs full plan: latents rel-L2 {rel:.2e}")
    assert rel < 2e-3, rel  # shared prefix: GroupNorm chunking differs, fp16 rounding noise only


@pytest.mark.parametrize("name,hw", [("tiny", 16), ("stage2", 32)])
def test_shared_prefix_plan_matches_full_plan(name, hw):
    """engine.UNetEngine.plan(uncond_zero=True, same_sample=True): when both halves of the CFG batch carry the same sample
    and timestep (lipsync_pipeline.py:542-549) everything before the first audio cross-attention is computed once.  Same
    arithmetic per element; only the GroupNorm partial-sum chunking differs (fp32 summation order), so the outputs agree
    to fp16 rounding noise"""
    unet, _, _ = get_unet(name)
    full = unet.plan(2, 16, hw, hw, 50)
    short = unet.plan(2, 16, hw, hw, 50, uncond_zero=True, same_sample=True)
    assert short is not full and short.same_sample and short.launches != full.launches
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randn(full.x_in.tensor().shape, generator=g, device="cuda").half()
    x[x.shape[0] // 2:] = x[: x.shape[0] // 2]  # the duplicated sample
    a = torch.randn(full.audio_in.tensor().shape, generator=g, device="cuda").half()
    a[: a.shape[0] // 2] = 0
    outs = []
    for plan in (full, short):
        plan.x_in.tensor().copy_(x)
        plan.audio_in.tensor().copy_(a)
        plan.t_in.tensor().fill_(301.0)
        plan.run_hoisted()
        plan.replay()
        torch.cuda.synchronize()
        outs.append(plan.eps_out.tensor().clone())
    rel = ((outs[0] - outs[1]).norm() / outs[0].norm()).item()
    print(f"{name}: shared-prefix plan vs full plan rel-L2 {rel:.2e}, launches {short.launches} vs {full.launches}")
    assert torch.isfinite(outs[1]).all() and rel < 2e-3  # two fp16 evaluations of the same network differ by rounding noise
    assert short.flops() < full.flops()


def test_restore_video_stage_vs_opencv():
    """SURVEY.md §8f rank 3 through the pipeline: LipsyncPipeline.restore_video (lipsync_pipeline.py:343-358) = resize
    + uint8 (rank 2 kernel) + AlignRestore.restore_img per frame; the GPU stage must give the bytes OpenCV gives for
    the same uint8 faces, for frames with two different box sizes and more frames than one launch takes"""
    import numpy as np
    pytest.importorskip("cv2")
    from oracle import restore_ref as RR

    pipe, _ = get_pipe("tiny")
    g = torch.Generator().manual_seed(77)
    n = 7
    dec = torch.rand(n, 3, 256, 256, generator=g) * 2.2 - 1.1
    cases = [RR.synthetic_case(400 + i, 360, 640, (0.9, 1.5), (40.0, 200.0)) for i in range(n)]
    frames = np.stack([c[0] for c in cases])
    boxes = [[0, 0, 210, 280]] * 5 + [[0, 0, 200, 260]] * 2
    mats = [c[2] for c in cases]
    out = pipe._restore_video(dec, frames, boxes, mats, frames_per_call=3)
    assert out.shape == frames.shape and out.dtype == np.uint8
    for i in range(n):
        h, w = boxes[i][3], boxes[i][2]
        face = pipe.faces_to_uint8(dec[i:i + 1], h, w).cpu().numpy()[0]
        ref = RR.restore_img_cv2(frames[i], face, mats[i])  # off-size boxes: the mask stays ones(280, 210) (:97)
        assert np.array_equal(out[i], ref), f"frame {i}: {(out[i] != ref).sum()} bytes differ"


def _need(path):
    if not os.path.exists(path):
        pytest.skip(f"{os.path.basename(path)} not generated yet (python -m oracle.make_golden stage2)")
    return torch.load(path)


def test_unet_stage2_forward_vs_reference_golden():
    """BASELINE config 2: one CFG-batched forward of the full 1.27 B-parameter UNet vs the reference's fp32 output"""
    from latentsync_b200 import synthetic as syn

    gold = _need(os.path.join(GOLDEN, "unet_stage2_fwd.pt"))
    unet, _, _ = get_unet("stage2")
    seg = syn.segment_inputs(INPUT_SEED, 0, 16, 256, 256)
    x, a = cfg_batch(seg)
    y = unet(x.cuda(), 951, encoder_hidden_states=a.cuda()).sample
    e = rel_l2(y, gold["noise_pred"])
    print(f"stage2 forward vs reference golden: rel-L2 {e:.3e}")
    assert torch.isfinite(y).all() and e < TOL


def test_loop_stage2_vs_reference_golden():
    """BASELINE config 1/2: 20 DDIM steps, CFG 1.5.  (i) teacher-forced per-step guided noise at steps 0,5,10,15,19
    (fp32 reference latents as input), (ii) free-running per-step noise/latents, (iii) final frames PSNR >= 40 dB
    against the oracle VAE decode of the reference's final latents."""
    from latentsync_b200 import synthetic as syn
    from oracle import pipeline_ref as P

    gold = _need(os.path.join(GOLDEN, "loop_stage2.pt"))
    get_unet("stage2")
    pipe, vsd = get_pipe("stage2")
    seg = syn.segment_inputs(INPUT_SEED, 0, 16, 256, 256)
    args = (seg["audio_embeds"], seg["mask_latents"], seg["masked_image_latents"], seg["ref_latents"])
    # (i) teacher forced
    tf = {j: gold["latents_in_f32"][i] for i, j in enumerate(gold["tf_steps"])}
    teacher = [tf.get(j, gold["latents"][j - 1].float() if j else seg["latents"]) for j in range(20)]
    trace = {}
    pipe.denoise_segment(seg["latents"], *args, num_inference_steps=20, guidance_scale=1.5, trace=trace,
                         teacher_latents=teacher)
    for i, j in enumerate(gold["tf_steps"]):
        e = rel_l2(trace["noise_pred"][j], gold["noise_pred_f32"][i])
        print(f"teacher-forced step {j}: guided noise rel-L2 {e:.3e}")
        assert e < TOL
    # (ii) free running
    trace = {}
    lat = pipe.denoise_segment(seg["latents"], *args, num_inference_steps=20, guidance_scale=1.5, trace=trace)
    worst = 0.0
    for j in range(20):
        e_n = rel_l2(trace["noise_pred"][j], gold["noise_pred"][j])
        e_l = rel_l2(trace["latents"][j], gold["latents"][j])
        worst = max(worst, e_n)
        print(f"free-running step {j}: noise {e_n:.3e} latents {e_l:.3e}")
    assert worst < TOL
    # (iii) frames
    frames = pipe.decode_and_paste(lat, seg["ref_pixel_values"], seg["masks"])
    want = P.decode_and_paste(vsd, gold["final_latents"], seg)
    p = P.psnr(frames.cpu(), want)
    repaint = (~seg["masks"].bool()).expand_as(want)  # the mouth region the model actually generates
    p_in = P.psnr(frames.cpu()[repaint], want[repaint])
    print(f"final frames PSNR {p:.1f} dB (repainted region only: {p_in:.1f} dB)")
    assert p >= 40.0 and p_in >= 40.0


def test_loop_graph_is_bitwise_identical_to_the_per_step_loop():
    """LipsyncPipeline.loop_graph: the whole denoising loop of a segment (lipsync_pipeline.py:537-568) replayed as ONE CUDA
    graph gives the same bits as the per-step loop (concat13 / time row / UNet graph / CFG + DDIM launched one by one),
    for several segments through the same captured graph, with and without guidance"""
    from latentsync_b200 import synthetic as syn

    pipe, _ = get_pipe("tiny")
    segs = [syn.segment_inputs(INPUT_SEED, s, 16, 128, 128) for s in range(3)]
    for steps, g in ((4, 1.5), (3, 1.0)):
        outs = {}
        for mode in (True, False, True):
            pipe.loop_graph = mode
            outs.setdefault(mode, []).append([
                pipe.denoise_segment(sg["latents"], sg["audio_embeds"], sg["mask_latents"], sg["masked_image_latents"],
                                     sg["ref_latents"], num_inference_steps=steps, guidance_scale=g) for sg in segs])
        pipe.loop_graph = True
        for a, b, c in zip(outs[True][0], outs[False][0], outs[True][1]):
            assert torch.equal(a, b) and torch.equal(a, c)
        assert not torch.equal(outs[True][0][0], outs[True][0][1])
    assert len(pipe._loop_graphs) >= 2


def test_batched_segments_match_single_segment_path():
    """run_segments(..., segments_per_batch=2): two different segments advanced as one UNet batch give the same frames
    as one at a time (segments never interact; only GEMM tile shapes change, so agreement is to fp16 rounding)"""
    from latentsync_b200 import synthetic as syn
    from oracle import pipeline_ref as P

    pipe, _ = get_pipe("tiny")
    segs = [syn.segment_inputs(INPUT_SEED, s, 16, 128, 128) for s in range(3)]
    one = pipe.run_segments(segs, num_inference_steps=3, guidance_scale=1.5, segments_per_batch=1)
    two = pipe.run_segments(segs, num_inference_steps=3, guidance_scale=1.5, segments_per_batch=2)  # 2 + 1
    assert len(one) == len(two) == 3
    for a, b in zip(one, two):
        p = P.psnr(a.cpu(), b.cpu())
        print(f"batched vs single: PSNR {p:.1f} dB")
        assert p >= 55.0
    assert not torch.equal(one[0], one[1])  # the segments really are different
