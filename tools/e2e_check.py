"""Stage-by-stage parity report: CUDA engines vs the CPU oracle on one seeded checkpoint/image.
    python tools/e2e_check.py [H W] [--raw] [--simt]
"""
import os
import sys
import time

import torch

ROOT = __file__.rsplit("/", 2)[0]
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import crdr_oracle as orc  # noqa: E402
import fixtures  # noqa: E402
from crdr_b200 import native as nv  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    h, w = (int(args[0]), int(args[1])) if len(args) >= 2 else (128, 192)
    model, sd = fixtures.build_model(seed=0, calibrated="--raw" not in sys.argv)
    if "--simt" in sys.argv:
        model.set_engine_options(conv_engine=nv.ENGINE_SIMT)
    x = fixtures.image(1, h, w)
    q, beta = 1.5, 2.56
    t0 = time.time()
    eb, gc = orc.entropy_models(sd)
    o = orc.compress(sd, x, q, eb, gc)
    print(f"oracle compress {time.time() - t0:.1f}s; y std {o['y'].std():.3f} nonzero sym {(o['y_sym'] != 0).float().mean():.3f} "
          f"idx range {o['y_idx'].min().item()}..{o['y_idx'].max().item()} |yhat|max {o['y_hat'].abs().max():.1f}")
    eng = model.engine()
    nv.status_reset()
    a = eng.analysis(x.cuda(), q)
    torch.cuda.synchronize()
    nv.status_check()
    y = eng.to_nchw(a["y32"]); z = eng.to_nchw(a["z32"]); yhat = eng.to_nchw(a["yhat32"])
    print("y      rel err", rel(y, o["y"]))
    print("z      rel err", rel(z, o["z"]))
    print("z_sym  match  ", (a["z_sym"].cpu() == o["z_sym"]).float().mean().item())
    print("z_lik  rel    ", ((a["z_lik"].cpu() - o["z_lik"]).abs() / o["z_lik"]).max().item())
    sm = (a["y_sym"].cpu() == o["y_sym"])
    im = (a["y_idx"].cpu() == o["y_idx"])
    print("y_sym  match  ", sm.float().mean().item(), "mismatches", (~sm).sum().item(), "of", sm.numel())
    print("y_idx  match  ", im.float().mean().item())
    print("sym mismatches per slice", [(~sm[:, 32 * s:32 * s + 32]).sum().item() for s in range(10)])
    print("idx mismatches per slice", [(~im[:, 32 * s:32 * s + 32]).sum().item() for s in range(10)])
    lr = ((a["y_lik"].cpu() - o["y_lik"]).abs() / o["y_lik"])
    both = sm & im
    print("y_lik  rel max (matching elems)", lr[sm].max().item(), " frac>1e-3:", (lr[sm] > 1e-3).float().mean().item())
    print("yhat   rel err", rel(yhat, o["y_hat"]))
    ybits = eng.bits(a["y_lik"]).item(); zbits = eng.bits(a["z_lik"]).item()
    print("bits y", ybits, o["pred_y_bit"], "z", zbits, o["pred_z_bit"], "rel", abs(ybits + zbits - o["pred_y_bit"] - o["pred_z_bit"]) / (o["pred_y_bit"] + o["pred_z_bit"]))
    # codec round trip through the public API
    r = model.compress(x, q)
    print("stream bytes mine", [len(s) for s in r["string_list"]], "oracle", [len(s) for s in o["string_list"]],
          "equal:", r["string_list"] == o["string_list"])
    img, z_hat, y_hat = model.decompress(r["string_list"], beta=beta)
    print("decoder y_hat == encoder y_hat:", torch.equal(y_hat, r["y_hat"]), " z_hat:", torch.equal(z_hat, r["z_hat"]))
    oi, _, _, _ = orc.decompress(sd, o["string_list"], beta, eb, gc)
    p_m, p_o = orc.psnr_u8(x, img.cpu()), orc.psnr_u8(x, oi)
    print(f"PSNR mine {p_m:.4f} oracle {p_o:.4f} delta {p_m - p_o:+.4f} dB; image max abs diff {(img.cpu() - oi).abs().max():.2e}")


if __name__ == "__main__":
    main()
