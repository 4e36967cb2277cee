"""Three-way precision probe at the benchmarked shape: the CPU oracle in fp32 (the reference's arithmetic), the same
restatement evaluated in fp64 ("truth"), and the CUDA path.  Reports, for symbols and CDF-table indexes, the mismatch
rate of each fp32-class implementation against the fp64 truth and against each other, and the error of mu / sigma.
If the CUDA path is as close to the truth as the reference's own fp32 arithmetic is, the remaining oracle-vs-CUDA
differences are the reordering noise of fp32 itself, not a precision deficit of the 3-term fp16 split.

    python tools/parity_probe.py [H W] [q ...]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import fixtures  # noqa: E402
import crdr_oracle as orc  # noqa: E402  (test infrastructure: this tool is a checker, not a product path)


@torch.no_grad()
def truth64(sd, x, q, z_hat32):
    """The oracle's encoder-side arithmetic in float64 (same functions, double tensors); z_hat is taken from the fp32
    run (it is quantised, so both runs share it unless a z symbol flips)."""
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    y = orc.g_a(sd64, x.double(), q)
    hyper = orc.h_s(sd64, z_hat32.double())
    cm = orc.sub(sd64, "context_model")
    S, K = orc.CFG["num_slices"], orc.CFG["max_support"]
    hm, hs = torch.chunk(hyper, 2, dim=1)
    ys = torch.chunk(y, S, dim=1)
    hats, mus, sigmas = [], [], []
    for s in range(S):
        sup = hats[:K]
        ms, ss = torch.cat([hm] + sup, dim=1), torch.cat([hs] + sup, dim=1)
        mu = orc.slice_net(orc.sub(cm, f"mean_slice_transforms.{s}"), ms)
        sg = orc.slice_net(orc.sub(cm, f"scale_slice_transforms.{s}"), ss)
        yh = torch.round(ys[s] - mu) + mu
        lrp = orc.slice_net(orc.sub(cm, f"lrp_slice_transforms.{s}"), torch.cat([ms, yh], dim=1))
        hats.append(yh + 0.5 * torch.tanh(lrp))
        mus.append(mu)
        sigmas.append(sg)
    mu, sg = torch.cat(mus, 1), torch.cat(sigmas, 1)
    table = orc.get_scale_table().double()
    sgb = torch.clamp(sg, min=float(torch.tensor(orc.CFG["scale_bound"], dtype=torch.float32)))  # the fp32 bound, exactly
    idx = torch.full_like(sgb, len(table) - 1).int()
    for t in table[:-1]:
        idx -= (sgb <= t).int()
    return dict(y=y, mu=mu, sigma=sg, y_idx=idx, y_sym=torch.round(y - mu).int())


@torch.no_grad()
def teacher_forced(eng, o):
    """CUDA decoder arithmetic on the oracle's symbols (no flip can cascade): index mismatch rate, max |y_hat diff|."""
    z_sym, y_sym = o["z_sym"].int().cuda(), o["y_sym"].int().cuda()
    T, _ = eng.hyper_from_symbols(z_sym)
    seen = {}

    def source(s0, cnt, idx):
        seen["idx"] = idx
        return y_sym
    yhat32 = eng.charm.decode(T, eng.gp, source)
    y_hat = eng.to_nchw(yhat32).cpu()
    return rate(seen["idx"].cpu().int(), o["y_idx"]), (y_hat - o["y_hat"]).abs().max().item()


def rate(a, b):
    return 1.0 - (a == b).double().mean().item()


def main():
    args = [float(a) for a in sys.argv[1:]]
    h, w = (int(args[0]), int(args[1])) if len(args) >= 2 else (512, 768)
    qs = args[2:] or [0.0, 2.0, 4.0]
    model, sd = fixtures.build_model(seed=0, calibrated=True)
    eng = model.engine()
    eb, gc = orc.entropy_models(sd)
    x = fixtures.image(4, h, w, seed=100)
    print(f"{h}x{w}, calibrated weights; mismatch rates (fraction of elements)")
    print(f"{'q':>5} {'img':>3} | {'sym o32-t64':>11} {'sym gpu-t64':>11} {'sym gpu-o32':>11} | {'idx o32-t64':>11} {'idx gpu-t64':>11} "
          f"{'idx gpu-o32':>11} | {'sigma relerr o32':>16} {'mu abserr o32':>13} | teacher-forced CUDA decode: idx mismatch, max|y_hat diff|")
    for k, q in enumerate(qs):
        i = (k + 3) % 4
        o = orc.analysis(sd, x[i:i + 1], q, eb, gc)
        t = truth64(sd, x[i:i + 1], q, o["z_hat"])
        a = eng.analysis(x[i:i + 1].cuda(), q)
        g_sym, g_idx = a["y_sym"].cpu(), a["y_idx"].cpu()
        tf = teacher_forced(eng, o)
        # mu / sigma of the CUDA path are internal to the engine: they are judged through the symbols and indexes
        big = t["sigma"].abs() > 0.11
        rel_o = ((o["sigma"].double() - t["sigma"]).abs() / t["sigma"].abs())[big]
        mu_o = (o["mu"].double() - t["mu"]).abs()
        print(f"{q:5.2f} {i:3d} | {rate(o['y_sym'], t['y_sym']):11.2e} {rate(g_sym, t['y_sym']):11.2e} {rate(g_sym, o['y_sym']):11.2e} | "
              f"{rate(o['y_idx'], t['y_idx']):11.2e} {rate(g_idx, t['y_idx']):11.2e} {rate(g_idx, o['y_idx']):11.2e} | "
              f"{rel_o.mean().item():8.1e}/{rel_o.max().item():7.1e} {mu_o.mean().item():13.2e} | {tf[0]:.2e} {tf[1]:.2e}")


if __name__ == "__main__":
    main()
