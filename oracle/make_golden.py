"""TEST INFRASTRUCTURE - generates tests/golden/*.pt by executing the REFERENCE's own modules
(/root/reference/latentsync/models/*.py, imported unmodified behind oracle/diffusers_shim.py) on the synthetic
weights / inputs of latentsync_b200/synthetic.py.  Runs only in the build container (the reference tree does not
travel to the GPU box); the fixtures it writes are what the GPU parity tests compare against.

    python -m oracle.make_golden tiny        # seconds
    python -m oracle.make_golden tiny_stage1 # seconds: stage1.yaml variant (use_motion_module = false)
    python -m oracle.make_golden stage2      # ~12 min on 8 cores: one forward + the 20-step loop trace (config 1)

It also pins the CPU port (oracle/unet_ref.py) to the reference: rel-L2 between the two is asserted < 1e-5.
"""
from __future__ import annotations

import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from latentsync_b200 import synthetic as syn  # noqa: E402
from latentsync_b200.spec import STAGE2_UNET_CONFIG, TINY_UNET_CONFIG as TINY_CONFIG  # noqa: E402
from oracle import diffusers_shim as shim  # noqa: E402
from oracle import pipeline_ref as P  # noqa: E402
from oracle.unet_ref import unet_forward  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED = 0
INPUT_SEED = 11


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def reference_model(cfg):
    ref = shim.reference_unet(cfg).eval()
    sd = syn.unet_state_dict(cfg, seed=WEIGHT_SEED)
    ref.load_state_dict(dict(sd), strict=True)
    return ref, sd


def unet_inputs(seg, t_lat=None):
    """the (2,13,f,h,w) CFG batch of lipsync_pipeline.py:542-549 and the (2,f,S,D) audio batch of :503-507"""
    lat = seg["latents"] if t_lat is None else t_lat
    x = torch.cat([lat] * 2)
    x = torch.cat([x, torch.cat([seg["mask_latents"]] * 2), torch.cat([seg["masked_image_latents"]] * 2),
                   torch.cat([seg["ref_latents"]] * 2)], dim=1)
    a = seg["audio_embeds"][None]
    return x, torch.cat([torch.zeros_like(a), a])


def make_tiny():
    ref, sd = reference_model(TINY_CONFIG)
    seg = syn.segment_inputs(INPUT_SEED, 0, 16, 128, 128)
    x, a = unet_inputs(seg)
    taps = {}
    with torch.no_grad():
        y = ref(x, 951, encoder_hidden_states=a).sample
    yp = unet_forward(sd, TINY_CONFIG, x, 951, a, taps=taps)
    err = rel_l2(yp, y)
    print("tiny: port vs reference rel-L2", err)
    assert err < 1e-5
    trace = {}
    P.denoise_segment(lambda s, t, au: ref(s, t, encoder_hidden_states=au).sample, seg, steps=4, guidance=1.5,
                      trace=trace)
    torch.save({"config": "tiny", "noise_pred": y, "t": 951,
                "loop4_noise_pred": torch.stack(trace["noise_pred"]), "loop4_latents": torch.stack(trace["latents"]),
                "taps": {k: v.flatten()[:: max(1, v.numel() // 4096)].clone() for k, v in taps.items()}},
               os.path.join(GOLDEN, "unet_tiny.pt"))


def make_tiny_stage1():
    """configs/unet/stage1.yaml variant: use_motion_module = false (no temporal layers), quarter width"""
    cfg = dict(TINY_CONFIG)
    cfg["use_motion_module"] = False
    ref, sd = reference_model(cfg)
    seg = syn.segment_inputs(INPUT_SEED, 1, 16, 128, 128)
    x, a = unet_inputs(seg)
    with torch.no_grad():
        y = ref(x, 701, encoder_hidden_states=a).sample
    err = rel_l2(unet_forward(sd, cfg, x, 701, a), y)
    print("tiny stage1 (no motion modules): port vs reference rel-L2", err)
    assert err < 1e-5
    torch.save({"config": "tiny_stage1", "noise_pred": y, "t": 701}, os.path.join(GOLDEN, "unet_tiny_stage1.pt"))


def make_stage2():
    t0 = time.time()
    cfg = STAGE2_UNET_CONFIG
    ref, sd = reference_model(cfg)
    print("model + weights", time.time() - t0)
    seg = syn.segment_inputs(INPUT_SEED, 0, 16, 256, 256)
    x, a = unet_inputs(seg)
    with torch.no_grad():
        t1 = time.time()
        y = ref(x, 951, encoder_hidden_states=a).sample
        print("reference forward s:", time.time() - t1)
    yp = unet_forward(sd, cfg, x, 951, a)
    err = rel_l2(yp, y)
    print("stage2: port vs reference rel-L2", err)
    assert err < 1e-5
    torch.save({"config": "stage2", "noise_pred": y, "t": 951, "port_vs_reference": err},
               os.path.join(GOLDEN, "unet_stage2_fwd.pt"))
    trace = {}
    t1 = time.time()
    lat = P.denoise_segment(lambda s, t, au: ref(s, t, encoder_hidden_states=au).sample, seg, steps=20, guidance=1.5,
                            trace=trace)
    print("20-step loop s:", time.time() - t1)
    torch.save({"config": "stage2", "steps": 20, "guidance": 1.5,
                "noise_pred": torch.stack(trace["noise_pred"]).half(),
                "tf_steps": [0, 5, 10, 15, 19],
                "latents_in_f32": torch.stack([trace["latents_in"][j] for j in (0, 5, 10, 15, 19)]),
                "noise_pred_f32": torch.stack([trace["noise_pred"][j] for j in (0, 5, 10, 15, 19)]),
                "latents": torch.stack(trace["latents"]).half(), "final_latents": lat},
               os.path.join(GOLDEN, "loop_stage2.pt"))


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    if which == "tiny":
        make_tiny()
    elif which == "tiny_stage1":
        make_tiny_stage1()
    elif which == "stage2":
        make_stage2()
    else:
        raise SystemExit(f"unknown target {which}")
