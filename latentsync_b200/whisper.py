"""Whisper-tiny audio front end on the GPU: the `AudioEncoder` of the reference's vendored Whisper
(latentsync/whisper/whisper/model.py:129-171) as a CUDA-graph plan over this package's own kernels, and the
`Audio2Feature` class of latentsync/whisper/audio2feature.py behind the same methods (`audio2feat`, `feature2chunks`,
`get_sliced_feature`, `crop_overlap_audio_window`).

What the reference does per clip (audio2feature.py:102-115, whisper/transcribe.py:84-127): log-mel spectrogram of the whole
track, 3000-frame windows (30 s, zero padded), `model.encoder(window, include_embeddings=True)`, and of the five hidden
states per 20 ms position (embedding + 4 block outputs) the first (end - start) / 2 positions of each window are kept and
concatenated: feature_array [T, 5, 384].  Video frame i then takes 10 consecutive positions around int(i * 50 / fps)
(clamped) -> a [50, 384] chunk (audio2feature.py:24-48,85-100) - the `encoder_hidden_states` of the UNet.

Here: the two Conv1d become GEMMs over an explicit 3-tap im2col (ls_im2col1d), their GELU and the MLP's GELU run in the
GEMM's GEGLU epilogue with an all-ones value half (value * gelu(gate) with zero value weights and bias 1 - no extra
kernel, no change to the hot GEMM), the positional embedding is the residual operand of the second convolution's GEMM,
attention is ls_attention (6 heads x 64: q k^T / sqrt(64) equals the reference's q / 64^.25 times k / 64^.25), the
residual adds are GEMM epilogues, every block writes its output straight into its slice of the [5][T][384] layer array,
and all chunks of a clip are gathered by ONE launch (ls_whisper_chunks).  The log-mel spectrogram itself (torch.stft on
the host in the reference, whisper/audio.py:92-124) stays with the caller: `Audio2Feature.audio2feat` takes the mel
tensor, a waveform (then `log_mel` below restates audio.py with torch.stft - cuFFT, a library FFT) or a path (then the
reference's own `load_audio` must be importable: it shells out to ffmpeg).

fp16 tensor-core operands with fp32 accumulation like the rest of the package; parity against the reference's own
model.py is measured in tests/test_whisper_gpu.py.  No CPU fallback.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .engine import KPAD, Plan, _chk, _h, _stream

N_FRAMES = 3000  # mel frames per 30-second window (whisper/audio.py:19)
SAMPLE_RATE, N_FFT, HOP_LENGTH, N_MELS = 16000, 400, 160, 80
GELU_TILE = 256


def whisper_tiny_dims() -> Dict[str, int]:
    """ModelDimensions of checkpoints/whisper/tiny.pt (model.py:15-26), encoder half"""
    return dict(n_mels=80, n_audio_ctx=1500, n_audio_state=384, n_audio_head=6, n_audio_layer=4)


def encoder_param_spec(dims: Dict[str, int]) -> Dict[str, Tuple[int, ...]]:
    """state_dict keys / shapes of `Whisper.encoder` (model.py:129-141): the on-disk contract of tiny.pt"""
    d, m = dims["n_audio_state"], dims["n_mels"]
    spec = {
        "encoder.conv1.weight": (d, m, 3), "encoder.conv1.bias": (d,),
        "encoder.conv2.weight": (d, d, 3), "encoder.conv2.bias": (d,),
        "encoder.positional_embedding": (dims["n_audio_ctx"], d),
    }
    for i in range(dims["n_audio_layer"]):
        b = f"encoder.blocks.{i}"
        spec.update({
            f"{b}.attn.query.weight": (d, d), f"{b}.attn.query.bias": (d,),
            f"{b}.attn.key.weight": (d, d),
            f"{b}.attn.value.weight": (d, d), f"{b}.attn.value.bias": (d,),
            f"{b}.attn.out.weight": (d, d), f"{b}.attn.out.bias": (d,),
            f"{b}.attn_ln.weight": (d,), f"{b}.attn_ln.bias": (d,),
            f"{b}.mlp.0.weight": (4 * d, d), f"{b}.mlp.0.bias": (4 * d,),
            f"{b}.mlp.2.weight": (d, 4 * d), f"{b}.mlp.2.bias": (d,),
            f"{b}.mlp_ln.weight": (d,), f"{b}.mlp_ln.bias": (d,),
        })
    spec.update({"encoder.ln_post.weight": (d,), "encoder.ln_post.bias": (d,)})
    return spec


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """the `positional_embedding` buffer (model.py:48-55)"""
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    t = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([torch.sin(t), torch.cos(t)], dim=1)


def _gelu_as_geglu(w: torch.Tensor, b: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """GELU(x W^T + b) through the GEGLU epilogue (value * gelu(gate)): value rows with zero weights and bias 1"""
    n = w.shape[0]
    wf = torch.cat([torch.zeros_like(w), w])
    bf = torch.cat([torch.ones(n, dtype=torch.float32, device=w.device), b.float()])
    wp, bp = L.pack_geglu(wf, bf, GELU_TILE)
    return _h(wp), bp.float().contiguous()


def _conv1d_as_gemm(w: torch.Tensor, cpad: int) -> torch.Tensor:
    """Conv1d weight [N, C, 3] -> [N, 3 * cpad], K index = (tap, channel), channels zero padded to cpad"""
    n, c, k = w.shape
    p = torch.zeros(n, k, cpad, dtype=torch.float32, device=w.device)
    p[:, :, :c] = w.permute(0, 2, 1)
    return p.reshape(n, k * cpad)


class WhisperEncoderEngine:
    """packed weights of `Whisper.encoder`; one plan per window count"""

    def __init__(self, state_dict: Dict[str, torch.Tensor], dims: Optional[Dict[str, int]] = None, device="cuda"):
        self.dims = dict(dims or whisper_tiny_dims())
        self.device = torch.device(device)
        d = self.dims
        assert d["n_audio_state"] % d["n_audio_head"] == 0
        self.head_dim = d["n_audio_state"] // d["n_audio_head"]
        if self.head_dim not in (16, 32, 40, 64, 80, 160):
            raise NotImplementedError(f"attention head_dim {self.head_dim} is not built")
        if d["n_audio_state"] % (GELU_TILE // 2) != 0:
            raise NotImplementedError("n_audio_state must be a multiple of 128 (GEGLU tile of the GELU epilogue)")
        missing = [k for k in encoder_param_spec(d) if k not in state_dict and k != "encoder.positional_embedding"]
        if missing:
            raise KeyError(f"Whisper encoder weights missing from the state_dict: {missing[:4]} ...")
        self.plans: Dict[int, "WhisperPlan"] = {}
        self.w: Dict[str, torch.Tensor] = {}
        self._pack({k: v.detach().to(self.device, torch.float32) for k, v in state_dict.items() if k.startswith("encoder.")})

    def _pack(self, sd: Dict[str, torch.Tensor]) -> None:
        d, w = self.dims, self.w
        C = d["n_audio_state"]
        self.mel_pad = (d["n_mels"] + KPAD - 1) // KPAD * KPAD
        w["conv1.w"], w["conv1.b"] = _gelu_as_geglu(_conv1d_as_gemm(sd["encoder.conv1.weight"], self.mel_pad),
                                                    sd["encoder.conv1.bias"])
        w["conv2.w"], w["conv2.b"] = _gelu_as_geglu(_conv1d_as_gemm(sd["encoder.conv2.weight"], C), sd["encoder.conv2.bias"])
        pos = sd.get("encoder.positional_embedding")
        if pos is None:  # a buffer: computed like model.py:135 when the checkpoint does not carry it
            pos = sinusoids(d["n_audio_ctx"], C).to(self.device)
        w["pos"] = _h(pos)
        for i in range(d["n_audio_layer"]):
            b = f"encoder.blocks.{i}"
            w[f"{i}.qkv.w"] = _h(torch.cat([sd[f"{b}.attn.query.weight"], sd[f"{b}.attn.key.weight"],
                                            sd[f"{b}.attn.value.weight"]]))
            w[f"{i}.qkv.b"] = torch.cat([sd[f"{b}.attn.query.bias"], torch.zeros(C, device=self.device),
                                         sd[f"{b}.attn.value.bias"]]).float().contiguous()
            w[f"{i}.out.w"], w[f"{i}.out.b"] = _h(sd[f"{b}.attn.out.weight"]), sd[f"{b}.attn.out.bias"].contiguous()
            w[f"{i}.mlp0.w"], w[f"{i}.mlp0.b"] = _gelu_as_geglu(sd[f"{b}.mlp.0.weight"], sd[f"{b}.mlp.0.bias"])
            w[f"{i}.mlp2.w"], w[f"{i}.mlp2.b"] = _h(sd[f"{b}.mlp.2.weight"]), sd[f"{b}.mlp.2.bias"].contiguous()
            for ln in ("attn_ln", "mlp_ln"):
                w[f"{i}.{ln}.g"], w[f"{i}.{ln}.b"] = sd[f"{b}.{ln}.weight"].contiguous(), sd[f"{b}.{ln}.bias"].contiguous()

    def plan(self, nwin: int) -> "WhisperPlan":
        if nwin not in self.plans:
            self.plans[nwin] = WhisperPlan(self, nwin)
        return self.plans[nwin]


class WhisperPlan(Plan):
    """AudioEncoder.forward(mel, include_embeddings=True) for `nwin` windows (model.py:143-171).
    mel_in: fp32 [nwin, n_mels, 2 n_ctx]; layers: fp16 [n_layer + 1][nwin * n_ctx][n_state] - the reference's
    `embeddings` list (before ln_post, which transcribe's caller never reads: audio2feature.py:102-115)."""

    def __init__(self, eng: WhisperEncoderEngine, nwin: int):
        super().__init__(eng.device)
        self.eng, self.nwin = eng, nwin
        self._build()

    def _build(self) -> None:
        eng, w, d, n = self.eng, self.eng.w, self.eng.dims, self.nwin
        C, T, T0, nl = d["n_audio_state"], d["n_audio_ctx"], 2 * d["n_audio_ctx"], d["n_audio_layer"]
        heads, hd = d["n_audio_head"], eng.head_dim
        rows = n * T
        self.mel_in = self.static(n * d["n_mels"], T0, torch.float32)
        self.layers = self.static((nl + 1) * rows, C)
        self.pos = self.static(rows, C)
        self.pos.tensor().copy_(w["pos"].repeat(n, 1))
        lay = lambda i: self.layers.ptr + i * rows * C * 2  # noqa: E731
        # conv1 + GELU, conv2 (stride 2) + GELU + positional embedding (model.py:149-154)
        x0 = self.buf(n * T0, eng.mel_pad)
        self.call("ls_ncfhw_to_cl", self.mel_in.ptr, n, d["n_mels"], 1, T0, eng.mel_pad, 1.0, x0.ptr)
        c1 = self.buf(n * T0, 3 * eng.mel_pad)
        self.call("ls_im2col1d", x0.ptr, n, T0, eng.mel_pad, 1, c1.ptr)
        h1 = self.buf(n * T0, C)
        self.gemm([(c1.ptr, 3 * eng.mel_pad, 3 * eng.mel_pad, 1)], 1, 1, n * T0, w["conv1.w"], 2 * C, h1.ptr, C,
                  bias_ptr=w["conv1.b"].data_ptr(), flags=L.EPI_GEGLU, tile_n=GELU_TILE)
        del x0, c1
        c2 = self.buf(rows, 3 * C)
        self.call("ls_im2col1d", h1.ptr, n, T0, C, 2, c2.ptr)
        del h1
        self.gemm([(c2.ptr, 3 * C, 3 * C, 1)], 1, 1, rows, w["conv2.w"], 2 * C, lay(0), C, bias_ptr=w["conv2.b"].data_ptr(),
                  residual_ptr=self.pos.ptr, ldr=C, flags=L.EPI_GEGLU, tile_n=GELU_TILE)
        del c2
        # ResidualAttentionBlock x n_layer (model.py:110-126): x += attn(attn_ln(x)); x += mlp(mlp_ln(x))
        for i in range(nl):
            xin = lay(i)
            nrm = self.buf(rows, C)
            self.layernorm(xin, rows, C, w[f"{i}.attn_ln.g"], w[f"{i}.attn_ln.b"], nrm.ptr)
            qkv = self.buf(rows, 3 * C)
            self.gemm([(nrm.ptr, C, C, 1)], 1, 1, rows, w[f"{i}.qkv.w"], 3 * C, qkv.ptr, 3 * C,
                      bias_ptr=w[f"{i}.qkv.b"].data_ptr())
            o = nrm
            self.attention(qkv.ptr, qkv.ptr + 2 * C, qkv.ptr + 4 * C, o.ptr, 3 * C, 3 * C, 3 * C, C, n, heads, hd, T, T)
            del qkv
            x1 = self.buf(rows, C)
            self.gemm([(o.ptr, C, C, 1)], 1, 1, rows, w[f"{i}.out.w"], C, x1.ptr, C, bias_ptr=w[f"{i}.out.b"].data_ptr(),
                      residual_ptr=xin, ldr=C)
            self.layernorm(x1.ptr, rows, C, w[f"{i}.mlp_ln.g"], w[f"{i}.mlp_ln.b"], nrm.ptr)
            g = self.buf(rows, 4 * C)
            self.gemm([(nrm.ptr, C, C, 1)], 1, 1, rows, w[f"{i}.mlp0.w"], 8 * C, g.ptr, 4 * C,
                      bias_ptr=w[f"{i}.mlp0.b"].data_ptr(), flags=L.EPI_GEGLU, tile_n=GELU_TILE)
            self.gemm([(g.ptr, 4 * C, 4 * C, 1)], 1, 1, rows, w[f"{i}.mlp2.w"], C, lay(i + 1), C,
                      bias_ptr=w[f"{i}.mlp2.b"].data_ptr(), residual_ptr=x1.ptr, ldr=C)
            del nrm, o, x1, g

    def layer_tensor(self) -> torch.Tensor:
        """fp16 [n_layer + 1, nwin * n_ctx, n_state] view of the layer array"""
        d = self.eng.dims
        return self.layers.tensor().view(d["n_audio_layer"] + 1, self.nwin * d["n_audio_ctx"], d["n_audio_state"])


def mel_filterbank(n_mels: int = N_MELS, n_fft: int = N_FFT, sr: int = SAMPLE_RATE) -> torch.Tensor:
    """librosa.filters.mel(sr, n_fft, n_mels) (Slaney scale, Slaney normalisation) - the matrix the reference ships as
    whisper/assets/mel_filters.npz (audio.py:77-89); [n_mels, n_fft // 2 + 1] fp32"""
    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        mel = f / (200.0 / 3)
        log_t = f >= 1000.0
        return np.where(log_t, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) / (np.log(6.4) / 27.0), mel)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        f = m * (200.0 / 3)
        log_t = m >= 15.0
        return np.where(log_t, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), f)

    fft_f = np.linspace(0, sr / 2, n_fft // 2 + 1)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_f[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    wts = np.maximum(0, np.minimum(lower, upper))
    wts *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return torch.from_numpy(wts.astype(np.float32))


def log_mel(audio: torch.Tensor, filters: Optional[torch.Tensor] = None) -> torch.Tensor:
    """whisper/audio.py:92-124 `log_mel_spectrogram` for a 16 kHz waveform tensor -> [80, n_frames] fp32 (torch.stft)"""
    audio = audio.float()
    window = torch.hann_window(N_FFT, device=audio.device)
    stft = torch.stft(audio, N_FFT, HOP_LENGTH, window=window, return_complex=True)
    mag = stft[:, :-1].abs() ** 2
    f = (filters if filters is not None else mel_filterbank()).to(audio.device)
    spec = torch.clamp(f @ mag, min=1e-10).log10()
    spec = torch.maximum(spec, spec.max() - 8.0)
    return (spec + 4.0) / 4.0


class Audio2Feature:
    """latentsync/whisper/audio2feature.py:9-147 with the encoder and the chunk slicing on the GPU.

    `model_path`: a Whisper checkpoint as `whisper.load_model` reads it (torch.load -> {"dims", "model_state_dict"},
    whisper/__init__.py:95-117), or pass `state_dict=` (+ `dims=`) directly."""

    def __init__(self, model_path: str = "checkpoints/whisper/tiny.pt", device=None, audio_embeds_cache_dir=None,
                 num_frames: int = 16, audio_feat_length: Sequence[int] = (2, 2), state_dict=None, dims=None,
                 feature_dtype=torch.float16, output_device="cpu"):
        device = device or "cuda"
        if state_dict is None:
            ckpt = torch.load(model_path, map_location="cpu", weights_only=False)
            state_dict = ckpt["model_state_dict"]
            cd = ckpt["dims"]
            dims = dims or {k: (cd[k] if isinstance(cd, dict) else getattr(cd, k)) for k in whisper_tiny_dims()}
        self.engine = WhisperEncoderEngine(state_dict, dims, device)
        self.device = self.engine.device
        self.audio_embeds_cache_dir = audio_embeds_cache_dir
        self.num_frames = num_frames
        self.embedding_dim = self.engine.dims["n_audio_state"]
        self.audio_feat_length = list(audio_feat_length)
        self.feature_dtype = feature_dtype
        # where the feature array and the chunks are handed out: the reference returns CPU tensors (numpy round trip,
        # audio2feature.py:113-114) and latentsync/utils/repeat.py pads the chunk list with CPU zeros; None = leave them
        # on the GPU (one copy less per segment when the caller does not use those helpers)
        self.output_device = output_device

    def _out(self, t: torch.Tensor) -> torch.Tensor:
        return t if self.output_device is None else t.to(self.output_device)

    # ---- encoder ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def encode_mel(self, mel: torch.Tensor) -> torch.Tensor:
        """transcribe()'s window loop + _audio2feat's concatenation (transcribe.py:100-127, audio2feature.py:102-115):
        mel [n_mels, n_frames] -> feature_array [T, n_layer + 1, n_state] on the device, T = sum over windows of
        int((end - start) / 2)"""
        d = self.engine.dims
        win = 2 * d["n_audio_ctx"]
        n_frames = mel.shape[-1]
        nwin = max(1, (n_frames + win - 1) // win)
        plan = self.engine.plan(nwin)
        if plan.graph is None:
            plan.capture()
        m = torch.zeros(nwin, d["n_mels"], win, dtype=torch.float32, device=self.device)  # pad_or_trim: zero padding
        mm = mel.to(self.device, torch.float32)
        for i in range(nwin):
            seg = mm[:, i * win:(i + 1) * win]
            m[i, :, : seg.shape[1]] = seg
        plan.mel_in.tensor().view(nwin, d["n_mels"], win).copy_(m)
        plan.replay()
        valid = sum(int((min(s + win, n_frames) - s) / 2) for s in range(0, n_frames, win))
        self._plan, self._valid = plan, valid
        return self._out(plan.layer_tensor()[:, :valid].permute(1, 0, 2).to(self.feature_dtype).contiguous())

    def _audio2feat(self, audio) -> torch.Tensor:
        if isinstance(audio, str):
            try:
                from latentsync.whisper.whisper.audio import load_audio  # the reference's ffmpeg reader
            except Exception as e:  # pragma: no cover - needs the reference tree, ffmpeg and ffmpeg-python
                raise RuntimeError("reading an audio FILE needs the reference's whisper.audio.load_audio (ffmpeg); pass "
                                   "the waveform or the mel spectrogram instead") from e
            audio = torch.from_numpy(load_audio(audio))
        audio = torch.as_tensor(audio)
        mel = audio if audio.dim() == 2 else log_mel(audio.to(self.device))
        return self.encode_mel(mel)

    def audio2feat(self, audio_path) -> torch.Tensor:
        """audio2feature.py:117-135 incl. the on-disk cache of the feature array"""
        cache = self.audio_embeds_cache_dir
        if not cache or not isinstance(audio_path, str):
            return self._audio2feat(audio_path)
        path = os.path.join(cache, os.path.basename(audio_path) + ".pt")
        if os.path.isfile(path):
            try:
                return torch.load(path, weights_only=True)
            except Exception as e:
                print(f"{type(e).__name__} - {e} - {path}")
                os.remove(path)
        feat = self._audio2feat(audio_path)
        torch.save(feat, path)
        return feat

    # ---- slicing ------------------------------------------------------------------------------------------
    def _first(self, vid_idx: int, fps) -> int:
        return int(vid_idx * 50 / fps) - self.audio_feat_length[0] * 2  # audio2feature.py:36-37

    def _window(self) -> int:
        return (self.audio_feat_length[0] + self.audio_feat_length[1] + 1) * 2  # left_idx .. right_idx - 1

    def _gather(self, feature_array: torch.Tensor, firsts: List[int]) -> torch.Tensor:
        """[len(firsts), K * L, C] by one ls_whisper_chunks launch; feature_array: [T, L, C] (any float dtype, device)"""
        T, Lr, C = feature_array.shape
        fa = feature_array.to(self.device, torch.float16)
        lay = fa.permute(1, 0, 2).contiguous()  # [L][T][C]
        first = torch.tensor(firsts, dtype=torch.int32, device=self.device)
        K = self._window()
        f32 = self.feature_dtype == torch.float32
        out = torch.empty(len(firsts), K * Lr, C, dtype=torch.float32 if f32 else torch.float16, device=self.device)
        _chk(L.lib().ls_whisper_chunks(lay.data_ptr(), T, Lr, T, C, first.data_ptr(), len(firsts), K, int(f32),
                                       out.data_ptr(), _stream()), "ls_whisper_chunks")
        return self._out(out.to(self.feature_dtype))

    def get_sliced_feature(self, feature_array, vid_idx, fps=25):
        """audio2feature.py:24-48 -> (selected_feature [50, 384], selected_idx)"""
        length = len(feature_array)
        first = self._first(vid_idx, fps)
        idx = [min(max(first + k, 0), length - 1) for k in range(self._window())]
        return self._gather(feature_array, [first])[0], idx

    def feature2chunks(self, feature_array, fps):
        """audio2feature.py:85-100: the reference's loop emits chunks for i = 0, 1, ... until int(i * 50 / fps) exceeds
        len(feature_array) (that last chunk included) - here all of them in one launch"""
        mult = 50.0 / fps
        n = 0
        while True:
            start_idx = int(n * mult)
            n += 1
            if start_idx > len(feature_array):
                break
        chunks = self._gather(feature_array, [self._first(i, fps) for i in range(n)])
        return list(chunks.unbind(0))

    def crop_overlap_audio_window(self, audio_feat, start_index):
        """audio2feature.py:137-143"""
        return self._gather(audio_feat, [self._first(i, 25) for i in range(start_index, start_index + self.num_frames)])
