"""fp16 range report: max |activation| of every convolution output of the path, per sub-network, for
default-init and calibrated weights, and max |loss-scaled gradient| of the training step -- against fp16's 65504.
No trained checkpoint is available offline; the calibrated fixture (latent gain 30, scales up to 24) is the closest stand-in.
    python tools/activation_range.py > gpurun_out/activation_range.txt"""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import fixtures  # noqa: E402
from crdr_b200.train import CodecTrainer  # noqa: E402

DEV = "cuda:0"


def group_of(conv, tr):
    for name, eng in (("g_a", tr.ga), ("h_a", tr.ha), ("h_s", tr.hs), ("ChARM", tr.charm), ("g_s", tr.gs)):
        if conv.cfg is getattr(eng, "cfg", None):
            return name
    for name, eng in (("h_a", tr.ha), ("h_s", tr.hs)):
        for v in vars(eng).values():
            items = v if isinstance(v, list) else [v]
            if any(getattr(c, "cfg", None) is conv.cfg for c in items):
                return name
    return "?"


for config in ("crdr_stage_2.yaml", "crdr.yaml"):
    for calibrated in (False, True):
        model, _ = fixtures.build_model(seed=0, calibrated=calibrated, device=DEV, config=config)
        tr = CodecTrainer(model, device=DEV)
        x = fixtures.image(4, 256, 256, seed=5).to(DEV)
        g = torch.Generator(device=DEV).manual_seed(1)
        mk = lambda c, a, b: torch.rand((4, c, a, b), dtype=torch.float32, device=DEV, generator=g) - 0.5
        noise = {"z": mk(192, 4, 4), "y": mk(320, 16, 16)}
        stats = collections.OrderedDict()
        for q in (0.0, 2.0, 4.0):
            out = tr.forward(x, q, noise, beta=2.56 if config == "crdr.yaml" else None)
            for rec in tr.ctx.tape:
                if rec[0] != "conv":
                    continue
                conv, _, kw, o = rec[1], rec[2], rec[3], rec[4]
                t = o.hi if o is not None else kw.get("out_f32")
                if t is None:
                    continue
                m = float(t.float().abs().max())
                key = group_of(conv, tr)
                a = stats.setdefault(key, [0.0, ""])
                if m > a[0]:
                    a[0], a[1] = m, f"{conv.name} (q={q})"
            ld = tr.losses(x, out, q)
            tr.backward(x, out)
            gmax = max(float(G.hi.float().abs().max()) for G in tr._grads.values())
            a = stats.setdefault("gradients (x loss scale %g)" % tr.loss_scale, [0.0, ""])
            a[0] = max(a[0], gmax)
        print(f"{config}, {'calibrated' if calibrated else 'default-init'} weights, 4 crops 256x256, q in {{0, 2, 4}}:")
        for k, (m, where) in stats.items():
            print(f"    {k:38s} max |x| = {m:10.3f}   = 2^{torch.log2(torch.tensor(max(m, 1e-30))).item():5.1f}   headroom to 65504: {65504.0 / max(m, 1e-30):9.1f}x   {where}")
        del tr, model
        torch.cuda.empty_cache()
