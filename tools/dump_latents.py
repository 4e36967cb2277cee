"""Dump the coder inputs of the bench workload's first images (int16 symbols, uint8 CDF indexes, int32 z symbols) so that
the host coder can be profiled on real data without a GPU (tools/coder_bench.py --latents)."""
import sys
import numpy as np
import torch
ROOT = __file__.rsplit("/", 2)[0]
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import fixtures

model, _ = fixtures.build_model(seed=0, calibrated=True)
x = fixtures.image(4, 512, 768, seed=100)
out = {}
for q in (0.0, 1.5, 4.0):
    a = model.engine().analysis(model._to_device(x), q, compact=True)
    torch.cuda.synchronize()
    out[f"y_sym_q{q}"] = a["y_sym16"].cpu().numpy()
    out[f"y_idx_q{q}"] = a["y_idx8"].cpu().numpy()
    out[f"z_sym_q{q}"] = a["z_sym"].cpu().numpy()
    strs = model.compress_batch(x, q)
    print(q, [len(s["string_list"][2]) for s in strs], [len(s["string_list"][1]) for s in strs])
np.savez_compressed(ROOT + "/gpurun_out/r02_latents.npz", **out)
