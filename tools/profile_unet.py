"""ncu target: one eager UNet forward (+ optionally the VAE decode) of the stage2 config between cudaProfilerStart/Stop.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_unet.py
Prints the plan's launch descriptors in order to gpurun_out/launch_descs.txt so that rows can be joined by index.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from latentsync_b200 import synthetic as syn  # noqa: E402
from latentsync_b200.engine import UNetEngine, VAEDecoderEngine  # noqa: E402
from latentsync_b200.spec import STAGE2_UNET_CONFIG, TINY_UNET_CONFIG  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "unet"
cfg = TINY_UNET_CONFIG if "tiny" in sys.argv else STAGE2_UNET_CONFIG
torch.cuda.set_device(0)
if which == "unet":
    eng = UNetEngine({k: v.cuda() for k, v in syn.unet_state_dict(cfg, 0).items()}, cfg, "cuda")
    hw = 16 if "tiny" in sys.argv else 32
    # "short": the plan the pipeline (and bench.py) runs - null-audio shortcut + shared CFG prefix; default: the full plan
    short = "short" in sys.argv
    plan = eng.plan(2, 16, hw, hw, 50, uncond_zero=short, same_sample=short)
    plan.x_in.tensor().normal_()
    plan.audio_in.tensor().normal_()
    plan.t_in.tensor().fill_(951.0)
else:
    eng = VAEDecoderEngine({k: v.cuda() for k, v in syn.vae_decoder_state_dict().items()}, device="cuda")
    plan = eng.plan(16, 32, 32)
    plan.z_in.tensor()[:, :4].normal_()
plan.run()
plan.run()
torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"launch_descs_{which}.txt"), "w") as f:
    for i, (k, d, fl) in enumerate(zip(plan.kinds, plan.descs, plan.op_flops)):
        f.write(f"{i}\t{k}\t{fl:.0f}\t{d}\n")
torch.cuda.profiler.start()
plan.run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", len(plan.ops), "launches")
