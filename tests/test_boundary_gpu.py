"""The drop-in boundary executed end to end on the GPU:

* `prepare_latents` against an inline restatement of lipsync_pipeline.py:182-196, incl. the fp16 draw the reference's
  call site makes (:489-498);
* the BODY of `LipsyncPipeline.__call__` (lipsync_pipeline.py:361-604) with the reference's untouched pre / post stages
  (`latentsync.*`: video decode, face alignment, Whisper chunks, ffmpeg) replaced by small stand-ins registered in
  sys.modules - the frames it hands to `write_video` must be the frames the hot-path entry points produce for the same
  inputs, one segment at a time and batched;
* `run_clip` (sharded clip entry) on one GPU, and `gather_frames` / `run_clip` under NCCL when >= 2 GPUs are visible.
"""
import os
import socket
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from test_model_gpu import get_pipe, get_unet  # noqa: E402

H = W = 128
FRAMES = 16


# ------------------------------------------------------------------------------------------------ prepare_latents
def test_prepare_latents_matches_reference_statement():
    """ONE (b, 4, 1, h, w) draw in `dtype` on `device` from the generator, repeated over all frames of the clip, times
    init_noise_sigma (lipsync_pipeline.py:182-196); __call__ passes weight_dtype = fp16 (:489-498)"""
    pipe, _ = get_pipe("tiny")
    dev = torch.device("cuda")
    for dtype in (torch.float16, torch.float32):
        for nframes in (16, 48):
            g1 = torch.Generator(device=dev).manual_seed(1247)
            got = pipe.prepare_latents(1, nframes, 4, 256, 256, dtype, dev, g1)
            g2 = torch.Generator(device=dev).manual_seed(1247)
            shape = (1, 4, 1, 256 // pipe.vae_scale_factor, 256 // pipe.vae_scale_factor)
            want = torch.randn(shape, generator=g2, device=dev, dtype=dtype).to(dev).repeat(1, 1, nframes, 1, 1)
            want = want * pipe.scheduler.init_noise_sigma
            assert got.dtype == dtype and got.shape == (1, 4, nframes, 32, 32)
            assert torch.equal(got, want)
            assert torch.equal(got[:, :, 0], got[:, :, nframes - 1])  # the same noise in every frame (and segment)
    # a CPU generator draws on the CPU like the reference's randn(..., device=generator's device) would fail: the
    # reference passes `device`; a CPU device + CPU generator works too
    gc = torch.Generator().manual_seed(5)
    cpu = pipe.prepare_latents(1, 16, 4, 64, 64, torch.float32, torch.device("cpu"), gc)
    assert cpu.device.type == "cpu" and cpu.shape == (1, 4, 16, 8, 8)


# ------------------------------------------------------------------------------------- __call__ with stubbed stages
class _Recorder:
    def __init__(self):
        self.video = None
        self.fps = None
        self.audio = None
        self.commands = []


def _install_stubs(monkeypatch, rec, nframes_video, nchunks, box=(0, 0, 210, 280)):
    """stand-ins for the reference's untouched stages, shaped like the real ones (file:line of what they replace)"""
    from latentsync_b200 import synthetic as syn

    cases = [syn.restore_case(900 + i, 360, 640, (1.2, 1.6), (40.0, 200.0)) for i in range(nframes_video)]
    frames = np.stack([c[0] for c in cases])
    mats = [c[2] for c in cases]
    g = torch.Generator().manual_seed(321)
    faces = torch.randint(0, 256, (nframes_video, 3, H, W), generator=g, dtype=torch.uint8)

    def mod(name):
        m = types.ModuleType(name)
        monkeypatch.setitem(sys.modules, name, m)
        return m

    mod("latentsync")
    mod("latentsync.pipelines")
    mod("latentsync.utils")
    atv = mod("latentsync.pipelines.affine_transform_video")
    # affine_transform_video.py:8-21 -> (faces (n,3,H,W) uint8 tensor, video frames, boxes, affine matrices)
    atv.affine_transform_video = lambda image_processor, video_path: (faces, frames, [list(box)] * nframes_video, mats)
    ip = mod("latentsync.utils.image_processor")

    class ImageProcessor:  # image_processor.py:30-60: only the attributes __call__ touches on the fix_mask path
        def __init__(self, resolution, mask="fix_mask", device="cpu", mask_image=None):
            self.resolution, self.mask_image, self.restorer = resolution, mask_image, object()

        def prepare_masks_and_masked_images(self, images, affine_transform=False):
            raise AssertionError("fix_mask frames at the working resolution must take the CUDA pre-processing path")

    ip.ImageProcessor = ImageProcessor
    ip.load_fixed_mask = lambda resolution, path: syn.fixed_mask(resolution, resolution).expand(3, -1, -1).contiguous()
    rp = mod("latentsync.utils.repeat")

    def pad_whisper_chunks_end(chunks, shape, audio, sr, fps=25, divisible_by=16):  # repeat.py:164-209
        add = (divisible_by - len(chunks) % divisible_by) % divisible_by
        dur = add / fps
        out = list(chunks) + [torch.zeros(shape) for _ in range(add)]
        return out, torch.cat([audio, torch.zeros(int(dur * sr), dtype=audio.dtype)]), dur

    def repeat_to_length(x, n):  # repeat.py: tile the clip until it has n entries
        reps = -(-n // len(x))
        if isinstance(x, torch.Tensor):
            return x.repeat(reps, *([1] * (x.dim() - 1)))[:n]
        if isinstance(x, np.ndarray):
            return np.concatenate([x] * reps)[:n]
        return (list(x) * reps)[:n]

    rp.pad_whisper_chunks_end = pad_whisper_chunks_end
    rp.repeat_to_length = repeat_to_length
    rp.truncate_to_length = lambda x, n: x[:n]
    rp.pad_whisper_chunks = rp.pad_whisper_chunks_to_target = None
    ut = mod("latentsync.utils.util")
    ut.read_audio = lambda path: torch.zeros(16000 * 2)  # util.py: 16 kHz mono samples
    ut.read_video = lambda path, use_decord=False: frames

    def write_video(path, video, fps=25, use_darken=False, brightness_factor=1.0):  # util.py:115-210 (PNG -> ffmpeg)
        rec.video, rec.fps = np.array(video), fps

    ut.write_video = write_video
    sfm = mod("soundfile")
    sfm.write = lambda path, samples, sr: setattr(rec, "audio", (len(samples), sr))
    import subprocess

    monkeypatch.setattr(subprocess, "run", lambda cmd, shell=False, **k: rec.commands.append(cmd))

    class AudioEncoder:  # audio2feature.py:24-115: one (50, 384) Whisper chunk per video frame
        def audio2feat(self, path):
            return "features"

        def feature2chunks(self, feature_array, fps):
            gg = torch.Generator().manual_seed(99)
            return [torch.randn(50, 384, generator=gg) for _ in range(nchunks)]

    return AudioEncoder(), faces, frames, mats


def _full_pipe():
    from latentsync_b200 import synthetic as syn
    from latentsync_b200.pipeline import LipsyncPipeline
    from latentsync_b200.scheduler import DDIMScheduler
    from latentsync_b200.vae import AutoencoderKLDecoder

    unet, _, _ = get_unet("tiny")
    vae = AutoencoderKLDecoder({**syn.vae_decoder_state_dict(seed=0), **syn.vae_encoder_state_dict(seed=0)},
                               device="cuda")
    return LipsyncPipeline(vae, None, unet, DDIMScheduler()).to("cuda")


@pytest.mark.parametrize("spb", [1, 2])
def test_call_end_to_end_with_stubbed_stages(monkeypatch, tmp_path, spb):
    """30 Whisper chunks / 20 video frames -> padded to 32 chunks, faces repeated to 32 -> 2 segments of 16; the clip
    __call__ writes must equal the clip assembled from the hot-path entry points on the same inputs"""
    rec = _Recorder()
    pipe = _full_pipe()
    enc, faces, frames, mats = _install_stubs(monkeypatch, rec, nframes_video=20, nchunks=30)
    pipe.audio_encoder = enc
    monkeypatch.chdir(tmp_path)
    steps, gs = 3, 1.5
    gen = torch.Generator(device="cuda").manual_seed(2024)
    out = pipe("in.mp4", "in.wav", "out.mp4", num_frames=FRAMES, height=H, width=W, num_inference_steps=steps,
               guidance_scale=gs, weight_dtype=torch.float16, generator=gen, mask_image_path="mask.png",
               data_file_url="ignored", segments_per_batch=spb)
    assert out is None  # like the reference: the side effect is the mp4
    assert rec.video is not None and rec.video.shape == (32, 360, 640, 3) and rec.video.dtype == np.uint8
    assert rec.fps == 25 and rec.audio == (int(32 / 25 * 16000), 16000)
    assert len(rec.commands) == 1 and "out.mp4" in rec.commands[0] and "ffmpeg" in rec.commands[0]

    # the same clip from the entry points __call__ is made of, replaying its generator draws in the reference's order
    from latentsync_b200 import synthetic as syn

    gen = torch.Generator(device="cuda").manual_seed(2024)
    dev = torch.device("cuda")
    lat_all = pipe.prepare_latents(1, 32, 4, H, W, torch.float16, dev, gen).float()
    chunks = enc.feature2chunks(None, 25) + [torch.zeros(50, 384)] * 2
    faces32 = faces.repeat(2, 1, 1, 1)[:32]
    mask_image = syn.fixed_mask(H, W).expand(3, -1, -1).contiguous()
    segs = []
    for i in range(2):
        sl = slice(i * FRAMES, (i + 1) * FRAMES)
        ref_px, masked_px, masks = pipe.prepare_masks_and_masked_images(faces32[sl], mask_image)
        mask_lat, masked_lat = pipe.prepare_mask_latents(masks, masked_px, H, W, torch.float16, dev, gen, False)
        ref_lat = pipe.prepare_image_latents(ref_px, dev, torch.float16, gen, False)
        segs.append(dict(latents=lat_all[:, :, sl], audio_embeds=torch.stack(chunks[sl]), mask_latents=mask_lat,
                         masked_image_latents=masked_lat, ref_latents=ref_lat, ref_pixel_values=ref_px, masks=masks))
    dec = torch.cat([f.to(torch.float16) for f in pipe.run_segments(segs, steps, gs, segments_per_batch=spb)])
    vid = np.concatenate([frames, frames])[:32]
    want = pipe._restore_video(dec, vid, [[0, 0, 210, 280]] * 32, (mats * 2)[:32])
    assert np.array_equal(rec.video, want), f"{(rec.video != want).sum()} bytes differ"
    # and the loop really changed the faces (the paste-back region is not the input)
    assert (rec.video != vid).mean() > 0.01


def test_call_rejects_what_the_reference_rejects(monkeypatch, tmp_path):
    rec = _Recorder()
    pipe = _full_pipe()
    pipe.audio_encoder = _install_stubs(monkeypatch, rec, 16, 16)[0]
    monkeypatch.chdir(tmp_path)
    with pytest.raises(ValueError):  # check_inputs, lipsync_pipeline.py:168-180
        pipe("a", "b", "c", height=H, width=W, callback_steps=0, num_inference_steps=1)
    with pytest.raises(NotImplementedError):
        pipe("a", "b", "c", height=H, width=W, eta=0.5, num_inference_steps=1)


def test_delegated_encoder_follows_pipeline_to(monkeypatch):
    """INTEGRATION.md recipe: decoder weights here, `encoder=` a foreign AutoencoderKL built on the CPU; pipeline.to(cuda)
    must move it (the reference's pipeline.to moves the VAE), and prepare_* must then run"""
    from latentsync_b200 import synthetic as syn
    from latentsync_b200.pipeline import LipsyncPipeline
    from latentsync_b200.scheduler import DDIMScheduler
    from latentsync_b200.vae import AutoencoderKLDecoder

    class ForeignVAE(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(1))

        def encode(self, x):
            assert x.device == self.w.device, "encoder was not moved with the pipeline"
            mean = torch.nn.functional.avg_pool2d(x[:, :1].float(), 8).repeat(1, 4, 1, 1) * self.w

            class D:
                def sample(self, generator=None):
                    return mean

            return types.SimpleNamespace(latent_dist=D())

    unet, _, _ = get_unet("tiny")
    foreign = ForeignVAE()
    vae = AutoencoderKLDecoder(syn.vae_decoder_state_dict(seed=0), device="cuda", encoder=foreign)
    pipe = LipsyncPipeline(vae, None, unet, DDIMScheduler()).to("cuda")
    assert next(foreign.parameters()).device.type == "cuda"
    px = torch.rand(4, 3, 64, 64, device="cuda") * 2 - 1
    lat = pipe.prepare_image_latents(px, torch.device("cuda"), torch.float16, None, True)
    assert lat.shape == (2, 4, 4, 8, 8) and lat.is_cuda


# ----------------------------------------------------------------------------------------------- sharded clip entry
def test_run_clip_single_process_equals_run_segments():
    from latentsync_b200 import synthetic as syn

    pipe, _ = get_pipe("tiny")
    segs = [{k: v.cuda() for k, v in syn.segment_inputs(31, s, FRAMES, H, W).items()} for s in range(3)]
    want = torch.cat(pipe.run_segments(segs, 2, 1.5)).to(torch.float16)
    got = pipe.run_clip(lambda i: segs[i], num_segments=3, num_inference_steps=2, guidance_scale=1.5)
    assert got.dtype == torch.float16 and torch.equal(got, want)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from latentsync_b200 import synthetic as syn
    from latentsync_b200.pipeline import LipsyncPipeline
    from latentsync_b200.scheduler import DDIMScheduler
    from latentsync_b200.spec import TINY_UNET_CONFIG
    from latentsync_b200.unet import UNet3DConditionModel
    from latentsync_b200.vae import AutoencoderKLDecoder

    unet = UNet3DConditionModel.from_config(TINY_UNET_CONFIG)
    unet.load_state_dict(syn.unet_state_dict(TINY_UNET_CONFIG, seed=0))
    unet = unet.to(dev).eval()
    pipe = LipsyncPipeline(AutoencoderKLDecoder(syn.vae_decoder_state_dict(seed=0), device=dev), None, unet,
                           DDIMScheduler()).to(dev)
    nseg = 3  # ragged: rank 0 holds 2 segments, rank 1 one

    def seg(i):
        return {k: v.to(dev) for k, v in syn.segment_inputs(31, i, FRAMES, H, W).items()}

    clip = pipe.run_clip(seg, num_segments=nseg, num_inference_steps=2, guidance_scale=1.5)
    if rank == 0:
        want = torch.cat(pipe.run_segments([seg(i) for i in range(nseg)], 2, 1.5)).to(torch.float16)
        q.put(bool(clip is not None and clip.shape == want.shape and torch.equal(clip, want)))
    else:
        assert clip is None
    # uint8 payload through gather_frames directly, one empty shard
    u8 = torch.full((4 if rank == 0 else 0, 3, 8, 8), 7, dtype=torch.uint8, device=dev)
    g = LipsyncPipeline.gather_frames(u8)
    if rank == 0:
        q.put(bool(g.shape == (4, 3, 8, 8) and g.dtype == torch.uint8 and int(g.sum()) == 7 * 4 * 3 * 64))
    dist.barrier()
    dist.destroy_process_group()


def test_run_clip_and_gather_frames_under_nccl():
    """2 ranks on 2 GPUs: sharded clip == single-GPU clip, gathered over NCCL (skipped on a 1-GPU box; the gloo version of
    the same host logic runs in tests/test_dist_cpu.py)"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True and q.get(timeout=5) is True


# ------------------------------------------------------------------------------------------- CTA-pair GEMM (cta_group::2)
@pytest.mark.parametrize("M,N,K,bn", [(2048, 640, 1280, 160), (1152, 320, 640, 160), (4096, 1280, 2560, 256),
                                      (2048, 960, 320, 192), (384, 320, 320, 128)])
def test_gemm_cta_pair_matches_fp32_reference(M, N, K, bn):
    """gemm_tc_pair_kernel (256 x BN tiles, each CTA stages half of the B tile): bias + residual epilogue, an odd number
    of 128-row tiles (1152 = 9 tiles: the last pair's second half is out of range), ragged N"""
    from latentsync_b200 import _lib as L

    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).half()
    out = torch.empty(M, N, dtype=torch.float16, device="cuda")
    L.gemm([L.Seg(a, K, K, 1)], 1, 1, M, w, N, out, N, bias=bias, residual=res, ldr=N, tile_n=bn, cta_pair=2)
    want = a.float() @ w.float().t() + bias + res.float()
    err = ((out.float() - want).norm() / want.norm()).item()
    assert err < 2e-3, err
    single = torch.empty_like(out)
    L.gemm([L.Seg(a, K, K, 1)], 1, 1, M, w, N, single, N, bias=bias, residual=res, ldr=N, tile_n=bn, cta_pair=1)
    # same products; only split-K (single-CTA launches with few tiles) may change the summation order
    assert ((out.float() - single.float()).norm() / want.norm()).item() < 5e-4


def test_conv3x3_cta_pair_and_geglu_pair():
    from latentsync_b200 import _lib as L

    g = torch.Generator(device="cuda").manual_seed(5)
    nimg, Hh, Ww, cin, cout = 6, 16, 16, 128, 320
    x = torch.randn(nimg, Hh, Ww, cin, device="cuda", generator=g).half()
    wt = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (9 * cin) ** 0.5)
    from latentsync_b200.engine import pack_conv3x3

    wp = pack_conv3x3(wt)
    bias = torch.randn(cout, device="cuda", generator=g)
    out = torch.empty(nimg * Hh * Ww, cout, dtype=torch.float16, device="cuda")
    L.gemm([L.Seg(x.view(-1, cin), cin, cin, 9)], nimg, Hh, Ww, wp, cout, out, cout, bias=bias, tile_n=160, cta_pair=2)
    want = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.half().float(), bias, padding=1)
    want = want.permute(0, 2, 3, 1).reshape(-1, cout)
    assert ((out.float() - want).norm() / want.norm()).item() < 2e-3
    # GEGLU epilogue in pair mode
    M, C = 1024, 320
    a = torch.randn(M, C, device="cuda", generator=g).half()
    wg = torch.randn(8 * C, C, device="cuda", generator=g) / C ** 0.5
    bg = torch.randn(8 * C, device="cuda", generator=g)
    wpk, bpk = L.pack_geglu(wg, bg, 256)
    og = torch.empty(M, 4 * C, dtype=torch.float16, device="cuda")
    L.gemm([L.Seg(a, C, C, 1)], 1, 1, M, wpk.half().contiguous(), 8 * C, og, 4 * C, bias=bpk, flags=L.EPI_GEGLU,
           tile_n=256, cta_pair=2)
    hcat = a.float() @ wg.half().float().t() + bg
    wantg = hcat[:, : 4 * C] * torch.nn.functional.gelu(hcat[:, 4 * C:])
    assert ((og.float() - wantg).norm() / wantg.norm()).item() < 2e-3
