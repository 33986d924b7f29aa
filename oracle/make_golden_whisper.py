"""Golden vectors for the Whisper front end, produced by the REFERENCE's own modules (run in the build container only:
/root/reference does not exist on the GPU box).

    python -m oracle.make_golden_whisper

writes tests/golden/whisper_small.pt (a reduced-width encoder, full output), whisper_tiny.pt (the real whisper-tiny
dimensions, every 50th position of a 30 s window) and whisper_chunks.pt (index lists and chunk tensors of the reference's
Audio2Feature.get_sliced_feature / feature2chunks for several frame rates).  Weights and inputs are functions of a seed
(latentsync_b200/synthetic.py), so only outputs are stored.
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("LS_REFERENCE", "/root/reference")
GOLDEN = os.path.join(ROOT, "tests", "golden")
SMALL = dict(n_mels=80, n_audio_ctx=96, n_audio_state=128, n_audio_head=2, n_audio_layer=2)


def ref_modules():
    sys.modules.setdefault("ffmpeg", types.ModuleType("ffmpeg"))  # whisper/audio.py imports it at module level
    sys.path.insert(0, REF)
    from latentsync.whisper import audio2feature as A
    from latentsync.whisper.whisper import model as M

    return M, A


def ref_embeddings(M, dims, sd, mel):
    enc = M.AudioEncoder(dims["n_mels"], dims["n_audio_ctx"], dims["n_audio_state"], dims["n_audio_head"],
                         dims["n_audio_layer"]).eval()
    missing, unexpected = enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items()}, strict=True), None
    with torch.no_grad():
        _, emb = enc(mel, include_embeddings=True)
    return torch.from_numpy(emb)  # [B, L, T, C]


def main():
    from latentsync_b200 import synthetic as syn
    from latentsync_b200.whisper import whisper_tiny_dims
    from oracle import whisper_ref as W

    M, A = ref_modules()
    os.makedirs(GOLDEN, exist_ok=True)
    # (1) reduced width, two windows, full output
    sd = syn.whisper_encoder_state_dict(SMALL, seed=3)
    mel = torch.stack([syn.mel_like(21 + i, 80, 2 * SMALL["n_audio_ctx"]) for i in range(2)])
    emb = ref_embeddings(M, SMALL, sd, mel)
    port = W.encoder_embeddings(sd, SMALL, mel)
    print("small: port vs reference rel-L2", ((port - emb).norm() / emb.norm()).item())
    torch.save({"dims": SMALL, "seed": 3, "mel_seeds": [21, 22], "embeddings": emb.half()}, os.path.join(GOLDEN, "whisper_small.pt"))
    # (2) whisper-tiny dimensions, one 30 s window, every 50th position
    dims = whisper_tiny_dims()
    sd = syn.whisper_encoder_state_dict(dims, seed=0)
    mel = syn.mel_like(31, 80, 3000)[None]
    emb = ref_embeddings(M, dims, sd, mel)
    port = W.encoder_embeddings(sd, dims, mel)
    print("tiny: port vs reference rel-L2", ((port - emb).norm() / emb.norm()).item(), "per layer rms",
          [round(emb[0, l].pow(2).mean().sqrt().item(), 3) for l in range(emb.shape[1])])
    torch.save({"dims": dims, "seed": 0, "mel_seed": 31, "stride": 50, "embeddings": emb[:, :, ::50].clone()},
               os.path.join(GOLDEN, "whisper_tiny.pt"))
    # (3) the slicing of the reference's Audio2Feature (no model needed for these methods)
    a2f = object.__new__(A.Audio2Feature)
    a2f.num_frames, a2f.embedding_dim, a2f.audio_feat_length = 16, 8, [2, 2]
    feat = syn.approx_normal(5, "feat", (137, 5, 8))
    out = {"T": 137, "feat_seed": 5, "cases": {}}
    for fps in (25, 30, 24, 29.97, 50):
        chunks = a2f.feature2chunks(feature_array=feat, fps=fps)
        idx = [a2f.get_sliced_feature(feature_array=feat, vid_idx=i, fps=fps)[1] for i in range(len(chunks))]
        mine = W.feature2chunks(feat, fps)
        assert len(mine) == len(chunks) and all(torch.equal(a, b) for a, b in zip(mine, chunks)), fps
        out["cases"][str(fps)] = {"n": len(chunks), "idx": idx, "checksum": torch.stack(chunks).double().sum().item()}
    # (4) three rows of the mel filterbank the reference ships (whisper/assets/mel_filters.npz, audio.py:77-89): the
    #     product computes the matrix by librosa's formula and is checked against these
    import numpy as np

    mf = np.load(os.path.join(REF, "latentsync", "whisper", "whisper", "assets", "mel_filters.npz"))["mel_80"]
    out["mel_rows"] = {r: torch.from_numpy(mf[r].copy()) for r in (0, 40, 79)}
    torch.save(out, os.path.join(GOLDEN, "whisper_chunks.pt"))
    for f in ("whisper_small.pt", "whisper_tiny.pt", "whisper_chunks.pt"):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)) // 1024, "KB")


if __name__ == "__main__":
    main()
