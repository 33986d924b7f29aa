"""latentsync_b200 - B200-native (sm_100a) implementation of LatentSync's inference hot path:
the audio-conditioned UNet3DConditionModel denoising loop (CFG + DDIM) and the AutoencoderKL decode.

Host code mirrors the reference's Python interfaces (UNet3DConditionModel.forward, LipsyncPipeline.__call__);
all arithmetic runs in hand-written CUDA kernels behind the C-ABI in include/latentsync_b200.h.
"""
__version__ = "0.1.0"
