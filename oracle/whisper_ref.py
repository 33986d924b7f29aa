"""TEST INFRASTRUCTURE - CPU restatement (plain PyTorch fp32) of the Whisper audio front end the reference vendors:
`AudioEncoder.forward(x, include_embeddings=True)` (latentsync/whisper/whisper/model.py:129-171 with
ResidualAttentionBlock :110-126 and MultiHeadAttention :58-107), the window loop of `transcribe`
(whisper/transcribe.py:100-127), `_audio2feat` / `get_sliced_feature` / `feature2chunks`
(latentsync/whisper/audio2feature.py:24-48,85-115).

Pinned to the reference's OWN modules: oracle/make_golden_whisper.py imports model.py and audio2feature.py from
/root/reference, runs them on latentsync_b200/synthetic.py weights and inputs and commits the outputs under
tests/golden/whisper_*.pt; tests/test_cpu.py checks this restatement against them.  Only tests/ (and bench legs) may
import this module; the product path (latentsync_b200/whisper.py) never does.
"""
from typing import Dict, List

import torch
import torch.nn.functional as F


def encoder_embeddings(sd: Dict[str, torch.Tensor], dims: Dict[str, int], mel: torch.Tensor) -> torch.Tensor:
    """mel [B, n_mels, 2 n_ctx] fp32 -> [B, n_layer + 1, n_ctx, n_state]: the `embeddings` stack of model.py:157-168"""
    p = lambda k: sd["encoder." + k].float()  # noqa: E731
    heads = dims["n_audio_head"]
    x = F.gelu(F.conv1d(mel.float(), p("conv1.weight"), p("conv1.bias"), padding=1))
    x = F.gelu(F.conv1d(x, p("conv2.weight"), p("conv2.bias"), stride=2, padding=1))
    x = x.permute(0, 2, 1)
    x = x + p("positional_embedding")
    out = [x]
    for i in range(dims["n_audio_layer"]):
        b = f"blocks.{i}."
        n_state = x.shape[-1]
        h = F.layer_norm(x, (n_state,), p(b + "attn_ln.weight"), p(b + "attn_ln.bias"))
        q = F.linear(h, p(b + "attn.query.weight"), p(b + "attn.query.bias"))
        k = F.linear(h, p(b + "attn.key.weight"))
        v = F.linear(h, p(b + "attn.value.weight"), p(b + "attn.value.bias"))
        scale = (n_state // heads) ** -0.25
        B, T, _ = q.shape
        q = q.view(B, T, heads, -1).permute(0, 2, 1, 3) * scale
        k = k.view(B, T, heads, -1).permute(0, 2, 3, 1) * scale
        v = v.view(B, T, heads, -1).permute(0, 2, 1, 3)
        w = F.softmax((q @ k).float(), dim=-1)
        a = (w @ v).permute(0, 2, 1, 3).flatten(start_dim=2)
        x = x + F.linear(a, p(b + "attn.out.weight"), p(b + "attn.out.bias"))
        h = F.layer_norm(x, (n_state,), p(b + "mlp_ln.weight"), p(b + "mlp_ln.bias"))
        h = F.linear(F.gelu(F.linear(h, p(b + "mlp.0.weight"), p(b + "mlp.0.bias"))), p(b + "mlp.2.weight"), p(b + "mlp.2.bias"))
        x = x + h
        out.append(x)
    return torch.stack(out, dim=1)


def audio2feat(sd, dims, mel: torch.Tensor) -> torch.Tensor:
    """transcribe's 3000-frame windows (zero padded) + _audio2feat's concatenation: mel [n_mels, n_frames] ->
    [T, n_layer + 1, n_state]"""
    win = 2 * dims["n_audio_ctx"]
    n_frames = mel.shape[-1]
    parts = []
    for seek in range(0, n_frames, win):
        end = min(seek + win, n_frames)
        seg = mel[:, seek:seek + win]
        if seg.shape[1] < win:
            seg = F.pad(seg, (0, win - seg.shape[1]))
        emb = encoder_embeddings(sd, dims, seg[None])  # [1, L, T, C]
        emb = emb.permute(0, 2, 1, 3)[0]              # [T, L, C]
        parts.append(emb[: int((end - seek) / 2)])
    return torch.cat(parts, dim=0)


def sliced_indices(length: int, vid_idx: int, fps, audio_feat_length=(2, 2)) -> List[int]:
    center = int(vid_idx * 50 / fps)
    left, right = center - audio_feat_length[0] * 2, center + (audio_feat_length[1] + 1) * 2
    return [min(max(i, 0), length - 1) for i in range(left, right)]


def feature2chunks(feature_array: torch.Tensor, fps, audio_feat_length=(2, 2)) -> List[torch.Tensor]:
    chunks, i = [], 0
    while True:
        start_idx = int(i * (50.0 / fps))
        idx = sliced_indices(len(feature_array), i, fps, audio_feat_length)
        chunks.append(torch.cat([feature_array[j] for j in idx], dim=0).reshape(-1, feature_array.shape[-1]))
        i += 1
        if start_idx > len(feature_array):
            break
    return chunks
