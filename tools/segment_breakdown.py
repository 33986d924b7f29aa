"""where the non-UNet milliseconds of a segment go: denoising loop, VAE decode graph, decode_and_paste, whole segment"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from latentsync_b200 import synthetic as syn
from latentsync_b200.pipeline import LipsyncPipeline
from latentsync_b200.scheduler import DDIMScheduler
from latentsync_b200.spec import STAGE2_UNET_CONFIG as cfg
from latentsync_b200.unet import UNet3DConditionModel
from latentsync_b200.vae import AutoencoderKLDecoder
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
unet = UNet3DConditionModel.from_config(cfg); unet.load_state_dict(syn.unet_state_dict(cfg, seed=0)); unet = unet.to(dev).eval()
vae = AutoencoderKLDecoder(syn.vae_decoder_state_dict(seed=0), device=dev)
pipe = LipsyncPipeline(vae, None, unet, DDIMScheduler()).to(dev)
seg = {k: v.to(dev) for k, v in syn.segment_inputs(100, 0, 16, 256, 256).items()}
uplan = unet.plan(2, 16, 32, 32, 50, uncond_zero=True, same_sample=True); vplan = vae.plan(16, 32, 32)
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
den = lambda: pipe.denoise_segment(seg["latents"], seg["audio_embeds"], seg["mask_latents"], seg["masked_image_latents"], seg["ref_latents"], 20, 1.5)
lat = den()
print(f"unet graph replay      : {t(uplan.replay, 20):8.3f} ms")
print(f"denoise_segment (20)   : {t(den):8.3f} ms")
print(f"vae graph replay       : {t(vplan.replay, 5):8.3f} ms")
print(f"decode_and_paste       : {t(lambda: pipe.decode_and_paste(lat, seg['ref_pixel_values'], seg['masks'])):8.3f} ms")
print(f"run_segments (1 seg)   : {t(lambda: pipe.run_segments([seg], 20, 1.5)):8.3f} ms")
for k, (n, ms) in sorted(vplan.run_timed().items(), key=lambda kv: -kv[1][1]):
    print(f"   vae eager {k:14s} {n:3d} launches {ms:7.3f} ms")
