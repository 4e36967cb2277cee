"""Bring-up: per-tensor gradient errors of the CUDA training step against autograd through the oracle, plus the
gradients of the latents (y, z, y_hat).  python tools/train_check.py [q] > gpurun_out/train_check.log"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import crdr_oracle as oracle  # noqa: E402
from compressai import ans  # noqa: E402
import fixtures  # noqa: E402
from crdr_b200.train import CodecTrainer  # noqa: E402

ans.build_lib()
DEV = "cuda:0"
q = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
model, sd = fixtures.build_model(seed=5, calibrated=True, config="crdr_stage_2.yaml")
tr = CodecTrainer(model, device=DEV)
n, h, w = 2, 128, 128
x = fixtures.image(n, h, w, seed=21)
g = torch.Generator().manual_seed(77)
noise = {"z": torch.rand(n, 192, h // 64, w // 64, generator=g) - 0.5, "y": torch.rand(n, 320, h // 16, w // 16, generator=g) - 0.5}
rate_w, lam = 0.8, 150.0

xd = x.to(DEV).contiguous()
out = tr.forward(xd, q, {k: v.to(DEV).contiguous() for k, v in noise.items()})
tr.backward(xd, out, rate_w)
torch.cuda.synchronize()
forced = out["y_sym"].cpu() if os.environ.get("FORCE_SYMBOLS", "1") == "1" else None
sdr = {k: (v.detach().clone().float().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
eb, gc = oracle.entropy_models(sdr)
for p in eb.parameters():
    p.requires_grad_(True)
with torch.enable_grad():
    ro = oracle.forward_train.__wrapped__(sdr, x, q, None, noise, eb, gc, forced_y_symbols=forced)
    for t in (ro["latent_code"]["y"], ro["latent_code"]["z"], ro["quantized_code"]["y"], ro["quantized_code"]["z"], ro["fake_images"]):
        t.retain_grad()
    bits = lambda lik: (-torch.log2(lik)).sum((1, 2, 3))
    bpp = (bits(ro["likelihoods"]["y"]) + bits(ro["likelihoods"]["z"])) / (h * w)
    mse = torch.mean(((x + 1) / 2 - (ro["fake_images"] + 1) / 2) ** 2)
    (rate_w * bpp.mean() + lam * mse).backward()
ref = {k: v.grad for k, v in sdr.items() if v.is_floating_point() and v.grad is not None}
for k, p in eb.named_parameters():
    if p.grad is not None and k != "quantiles":
        ref["entropy_model_z." + k] = p.grad

S = tr.loss_scale
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
nchw = lambda act: act.to_nchw().cpu() / S
G = tr._grads
print("loss scale", S)
yh_g = out["yhat32"].permute(0, 3, 1, 2).cpu()
yh_r = ro["quantized_code"]["y"].detach()
for s_ in range(10):
    dd = (yh_g[:, 32 * s_:32 * s_ + 32] - yh_r[:, 32 * s_:32 * s_ + 32]).abs()
    print("   y_hat slice", s_, "max diff", float(dd.max()), "flips", int((dd > 0.5).sum()), "of", dd.numel())
print("d fake     ", "n/a")
T = out["T"]
GT = nchw(G[T.hi.data_ptr()])
ch = tr.charm
print("d y_hat    ", rel(GT[:, ch.off_y:ch.off_y + 320], ro["quantized_code"]["y"].grad), float(ro["quantized_code"]["y"].grad.norm()))
for s in range(10):
    a, b = GT[:, ch.off_y + 32 * s:ch.off_y + 32 * s + 32], ro["quantized_code"]["y"].grad[:, 32 * s:32 * s + 32]
    print("   slice", s, rel(a, b), float(a.norm()), float(b.norm()))
print("d y        ", rel(nchw(G[out["y_act"].hi.data_ptr()]), ro["latent_code"]["y"].grad), float(ro["latent_code"]["y"].grad.norm()))
gy, ry = nchw(G[out["y_act"].hi.data_ptr()]), ro["latent_code"]["y"].grad
for s in range(10):
    print("   slice", s, rel(gy[:, 32 * s:32 * s + 32], ry[:, 32 * s:32 * s + 32]), float(gy[:, 32 * s:32 * s + 32].norm()), float(ry[:, 32 * s:32 * s + 32].norm()))
print("d z        ", rel(nchw(G[out["z32"].data_ptr()]), ro["latent_code"]["z"].grad), float(ro["latent_code"]["z"].grad.norm()))
print("d z_hat    ", rel(nchw(G[tr._zhat.hi.data_ptr()]), ro["quantized_code"]["z"].grad) if hasattr(tr, "_zhat") else "n/a")
print("d hyper mu/sigma", float(GT[:, ch.off_mean:ch.off_mean + 320].norm()), float(GT[:, ch.off_scale:ch.off_scale + 320].norm()))
got = {k: v.detach().cpu() for k, v in tr.ctx.grads.items()}
for k in sorted(ref):
    if k in got and not k.endswith("quantiles"):
        print(f"{rel(got[k], ref[k]):10.3e}  |ref|={float(ref[k].norm()):.3e} |got|={float(got[k].norm()):.3e}  {k}")
