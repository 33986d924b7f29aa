"""debug: run a plan eagerly and check every GroupNorm-partials array against the tensor it describes"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from latentsync_b200 import engine as E  # noqa: E402
from latentsync_b200 import synthetic as syn  # noqa: E402

rec = []
orig = E.Plan.gn


def gn(self, srcs, rows, rows_per_inst, groups, gamma, beta, eps, silu, out_ptr):
    idx = len(self.ops)
    orig(self, srcs, rows, rows_per_inst, groups, gamma, beta, eps, silu, out_ptr)
    for b, ch in srcs:
        if b.gnp is not None:
            rec.append((idx, self.descs[idx], b.tensor(), b.gnp.tensor(), b.gnu, rows))


E.Plan.gn = gn
which = sys.argv[1] if len(sys.argv) > 1 else "enc"
if which == "enc":
    esd = syn.vae_encoder_state_dict(seed=0)
    eng = E.VAEEncoderEngine(esd, device="cuda")
    plan = eng.plan(2, 128, 128)
    plan.x_in.tensor().copy_(torch.randn_like(plan.x_in.tensor()) * 0.5)
else:
    dsd = syn.vae_decoder_state_dict(seed=0)
    eng = E.VAEDecoderEngine(dsd, device="cuda")
    plan = eng.plan(2, 16, 16)
    plan.z_in.tensor().copy_(torch.randn_like(plan.z_in.tensor()))
# run op by op; after each GroupNorm consumer index check its sources (they are still intact right after the norm)
by_idx = {}
for r in rec:
    by_idx.setdefault(r[0], []).append(r)
for i, fn in enumerate(plan.ops):
    fn()
    if i in by_idx:
        torch.cuda.synchronize()
        for (_, desc, x, gp, unit, rows) in by_idx[i]:
            xr = x[:rows].float()
            U = x.shape[1] // unit
            xf = xr.reshape(rows // 128, 128, U, unit)
            gpv = gp.reshape(gp.shape[0], U, 2)[: rows // 128]
            s_ok = torch.allclose(gpv[..., 0], xf.sum((1, 3)), atol=5e-2, rtol=1e-4)
            q_ok = torch.allclose(gpv[..., 1], (xf * xf).sum((1, 3)), rtol=1e-3, atol=1e-3)
            nan = torch.isnan(gpv).sum().item()
            print(f"op {i:3d} {desc:60s} x {tuple(x.shape)} unit {unit} nan-in-x {torch.isnan(xr).sum().item()} "
                  f"nan-in-parts {nan} sums {'ok' if s_ok else 'BAD'} squares {'ok' if q_ok else 'BAD'}  prev gemm: {plan.descs[i - 1][:90]}",
                  flush=True)
            if not (s_ok and q_ok):
                bad = (~torch.isclose(gpv[..., 1], (xf * xf).sum((1, 3)), rtol=1e-3, atol=1e-3)).nonzero()
                print("   first bad (tile, unit):", bad[:8].tolist(), "of", bad.shape[0], "entries; tiles", gpv.shape[0], "units", U)
