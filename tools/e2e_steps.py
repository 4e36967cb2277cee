"""Wall-clock of bench.py's e2e step, split per call (host-side diagnosis; never a bench number)."""
import sys, time
import torch
ROOT = __file__.rsplit("/", 2)[0]
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import fixtures
SWEEP = [0.0, 1.0, 2.0, 3.0, 4.0]
BETAS = [0.0, 3.84]
model, _ = fixtures.build_model(seed=0, calibrated=True)
import os
model.pipeline_chunks = int(os.environ.get('CHUNKS', '2'))
if os.environ.get('WEIGHTS'):
    model.pipeline_weights = tuple(float(v) for v in os.environ['WEIGHTS'].split(','))
x = (fixtures.image(24, 512, 768, seed=100) * 127.5 + 127.5).clamp(0, 255).to(torch.uint8).pin_memory()   # uint8 RGB like bench.py
host = None
def T(): torch.cuda.synchronize(); return time.perf_counter()
for i in range(8):
    q, beta = SWEEP[i % 5], BETAS[i % 2]
    t0 = T(); outs = model.compress_batch(x, q); t1 = T()
    img, _, _ = model.decompress_batch([o["string_list"] for o in outs], beta=beta, out_uint8=True); t2 = T()
    if host is None:
        host = torch.empty(img.shape, dtype=img.dtype, pin_memory=True)
    t3 = T(); host.copy_(img, non_blocking=True); t4 = T()
    print(f"step {i} q={q} beta={beta}: compress {1e3*(t1-t0):.1f} decompress {1e3*(t2-t1):.1f} pinned alloc {1e3*(t3-t2):.1f} d2h image {1e3*(t4-t3):.1f} ms")
