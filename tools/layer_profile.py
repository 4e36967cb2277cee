"""Per-launch device time of one encode+decode step (CUDA events around every conv launch).
    python tools/layer_profile.py [batch] [H W]   -> table sorted by time + per-shape aggregate
"""
import sys
import collections

import torch

ROOT = __file__.rsplit("/", 2)[0]
sys.path.insert(0, ROOT)
sys.path.insert(0, ROOT + "/tests")
import fixtures  # noqa: E402
from crdr_b200 import engine as E  # noqa: E402


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (512, 768)
    model, _ = fixtures.build_model(seed=0, calibrated=True)
    eng = model.engine()
    x = fixtures.image(b, h, w).cuda()

    def step():
        a = eng.analysis(x, 1.5)
        eng.decode_device(a["z_sym"], a["y_sym"], 1.5, 3.84, (h, w))

    step(); step()
    torch.cuda.synchronize()
    E.PROFILE.clear(); E.PROFILE_ON[0] = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record()
    torch.cuda.synchronize()
    E.PROFILE_ON[0] = False
    total = e0.elapsed_time(e1)
    agg = collections.OrderedDict()
    for fl, a, c, info in E.PROFILE:
        ms = a.elapsed_time(c)
        key = (info["m"], info["n"], info["k"], info["taps"], info["tile_n"], info["prec"], info["transposed"], info["stride"])
        t = agg.setdefault(key, [0, 0.0, 0.0])
        t[0] += 1; t[1] += ms; t[2] += fl
    conv_total = sum(v[1] for v in agg.values())
    print(f"step {total:.2f} ms, conv launches {len(E.PROFILE)} sum {conv_total:.2f} ms")
    print(f"{'M':>9} {'N':>4} {'K':>6} taps tile prec tr s | cnt {'ms':>8} {'%':>5} {'TFLOP/s':>8} {'eff.MMA TF/s':>12}")
    for key, (cnt, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        m, n, k, taps, tile, prec, tr, s = key
        mult = 3 if prec == 0 else 1
        print(f"{m:9d} {n:4d} {k:6d} {taps:4d} {tile:4d} {'x3' if prec == 0 else 'x1':>4} {int(tr):2d} {s} | {cnt:3d} {ms:8.3f} "
              f"{100 * ms / conv_total:5.1f} {fl / ms / 1e9:8.1f} {mult * fl / ms / 1e9:12.1f}")


if __name__ == "__main__":
    main()
