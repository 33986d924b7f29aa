"""debug: the first launches of the Whisper encoder plan at whisper-tiny size, each checked against torch"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from latentsync_b200 import _lib as L  # noqa: E402
from latentsync_b200 import synthetic as syn  # noqa: E402
from latentsync_b200.whisper import WhisperEncoderEngine, whisper_tiny_dims  # noqa: E402


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


dims = whisper_tiny_dims()
sd = syn.whisper_encoder_state_dict(dims, seed=0)
eng = WhisperEncoderEngine(sd, dims, "cuda")
w = eng.w
mel = syn.mel_like(31, 80, 3000)[None].cuda()
n, T0, C = 1, 3000, 384
st = torch.cuda.current_stream().cuda_stream
x0 = torch.empty(n * T0, 128, dtype=torch.float16, device="cuda")
L._check(L.lib().ls_ncfhw_to_cl(mel.data_ptr(), n, 80, 1, T0, 128, 1.0, x0.data_ptr(), st), "cl")
want_x0 = torch.zeros(T0, 128, device="cuda")
want_x0[:, :80] = mel[0].t()
print("x0", rel(x0, want_x0))
c1 = torch.empty(n * T0, 384, dtype=torch.float16, device="cuda")
L._check(L.lib().ls_im2col1d(x0.data_ptr(), n, T0, 128, 1, c1.data_ptr(), st), "im2col")
pad = F.pad(x0.float(), (0, 0, 1, 1))
want_c1 = torch.cat([pad[0:T0], pad[1:T0 + 1], pad[2:T0 + 2]], dim=1)
print("c1", rel(c1, want_c1))
h1 = torch.empty(n * T0, C, dtype=torch.float16, device="cuda")
for tile in (256, 128, 64):
    try:
        L.gemm([L.Seg(c1, 384, 384, 1)], 1, 1, n * T0, w["conv1.w"], 2 * C, h1, C, bias=w["conv1.b"], flags=L.EPI_GEGLU, tile_n=tile)
        want_h1 = F.gelu(F.conv1d(mel.float(), sd["encoder.conv1.weight"].cuda(), sd["encoder.conv1.bias"].cuda(), padding=1))[0].t()
        print("h1 tile", tile, rel(h1, want_h1), "abs mean got", h1.float().abs().mean().item(), "want", want_h1.abs().mean().item())
    except RuntimeError as e:
        print("h1 tile", tile, "error", e)
for pair in (1, 2):
    L.gemm([L.Seg(c1, 384, 384, 1)], 1, 1, n * T0, w["conv1.w"], 2 * C, h1, C, bias=w["conv1.b"], flags=L.EPI_GEGLU, tile_n=256, cta_pair=pair)
    print("h1 pair", pair, rel(h1, want_h1))
# plain GEMM of the gate half for reference
gate = torch.empty(n * T0, C, dtype=torch.float16, device="cuda")
wg = eng.w["conv1.w"]
plan = eng.plan(1)
plan.mel_in.tensor().view(1, 80, T0).copy_(mel)
plan.run()
torch.cuda.synchronize()
lay = plan.layer_tensor()
import oracle.whisper_ref as W  # noqa: E402
port = W.encoder_embeddings(sd, dims, mel.cpu())[0]
for l in range(5):
    print("layer", l, rel(lay[l].cpu(), port[l]), "abs mean got", lay[l].float().abs().mean().item())
