"""where does the multi-GPU pass spend its time?  torchrun --nproc-per-node 2 tools/gather_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from latentsync_b200 import synthetic as syn
from latentsync_b200.pipeline import LipsyncPipeline
from latentsync_b200.scheduler import DDIMScheduler
from latentsync_b200.spec import STAGE2_UNET_CONFIG
from latentsync_b200.unet import UNet3DConditionModel
from latentsync_b200.vae import AutoencoderKLDecoder
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
unet = UNet3DConditionModel.from_config(STAGE2_UNET_CONFIG); unet.load_state_dict(syn.unet_state_dict(STAGE2_UNET_CONFIG, seed=0)); unet = unet.to(dev).eval()
pipe = LipsyncPipeline(AutoencoderKLDecoder(syn.vae_decoder_state_dict(seed=0), device=dev), None, unet, DDIMScheduler()).to(dev)
segs = [{k: v.to(dev) for k, v in syn.segment_inputs(100 + rank, s, 16, 256, 256).items()} for s in range(3)]
def one_pass(tag):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    frames = torch.cat([f.half() for f in pipe.run_segments(segs, 20, 1.5)])
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out = pipe.gather_frames(frames)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"rank {rank} {tag}: run_segments(3) {1e3 * (t1 - t0):.1f} ms, gather {1e3 * (t2 - t1):.1f} ms", flush=True)
for i in range(4): one_pass(f"pass {i}")
dist.destroy_process_group()
