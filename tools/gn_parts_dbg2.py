"""debug: the VAE-encode test's own flow; find the first GroupNorm whose source / partials / output carry a NaN"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from latentsync_b200 import engine as E  # noqa: E402
from latentsync_b200 import synthetic as syn  # noqa: E402
from latentsync_b200 import _lib as L  # noqa: E402

rec = []
orig = E.Plan.gn


def gn(self, srcs, rows, rows_per_inst, groups, gamma, beta, eps, silu, out_ptr):
    idx = len(self.ops)
    orig(self, srcs, rows, rows_per_inst, groups, gamma, beta, eps, silu, out_ptr)
    cc = sum(ch for _, ch in srcs)
    rec.append((id(self), idx, self.descs[idx], [(b.tensor(), b.gnp.tensor() if b.gnp is not None else None, b.gnu) for b, _ in srcs],
                out_ptr, rows, cc))


E.Plan.gn = gn
from latentsync_b200.vae import AutoencoderKL  # noqa: E402

vsd = syn.vae_decoder_state_dict(seed=0)
esd = syn.vae_encoder_state_dict(seed=0)
vae = AutoencoderKL({**vsd, **esd}, device="cuda")
seg = syn.segment_inputs(11, 0, 2, 128, 128)
masked_px = (seg["ref_pixel_values"] * seg["masks"]).cuda()
dist = vae.encode(masked_px).latent_dist
print("after encode(): nan in moments", torch.isnan(dist.mean).sum().item())
plan = vae.encode_plan(2, 128, 128)
n, c, H, W = masked_px.shape
st = torch.cuda.current_stream().cuda_stream
L._check(L.lib().ls_ncfhw_to_cl(masked_px.float().contiguous().data_ptr(), n, c, 1, H * W, plan.x_in.cols, 1.0, plan.x_in.ptr, st), "x")
by_idx = {r[1]: r for r in rec if r[0] == id(plan)}
for i, fn in enumerate(plan.ops):
    fn()
    if i in by_idx:
        torch.cuda.synchronize()
        _, _, desc, srcs, out_ptr, rows, cc = by_idx[i]
        msg = []
        for (x, gp, unit) in srcs:
            xr = x[:rows].float()
            msg.append(f"x nan {torch.isnan(xr).sum().item()} inf {torch.isinf(xr).sum().item()} absmax {xr.abs().max().item():.3g}")
            if gp is not None:
                U = x.shape[1] // unit
                gpv = gp.reshape(gp.shape[0], U, 2)[: rows // 128]
                xf = xr.reshape(rows // 128, 128, U, unit)
                ok = torch.allclose(gpv[..., 1], (xf * xf).sum((1, 3)), rtol=1e-3, atol=1e-3)
                msg.append(f"parts nan {torch.isnan(gpv).sum().item()} inf {torch.isinf(gpv).sum().item()} squares {'ok' if ok else 'BAD'}")
        print(f"op {i:3d} {desc[:58]:58s} " + " | ".join(msg) + f" | prev: {plan.descs[i - 1][:70]}", flush=True)
torch.cuda.synchronize()
print("eager: nan in moments", torch.isnan(plan.mom_out.tensor()).sum().item())
