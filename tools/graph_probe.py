"""Probe: how much of the step is launch gaps?  Eager launches vs one CUDA graph of the same launch sequence."""
import sys, time
import torch
ROOT = __file__.rsplit("/", 2)[0]
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import fixtures
model, _ = fixtures.build_model(seed=0, calibrated=True)
eng = model.engine()
x = fixtures.image(24, 512, 768).cuda()
def step():
    a = eng.analysis(x, 1.5)
    img, _, _ = eng.decode_device(a["z_sym"], a["y_sym"], 1.5, 3.84, (512, 768))
    return img
for _ in range(3):
    ref = step()
torch.cuda.synchronize()
def timed(fn, n=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print(f"eager {timed(step):.2f} ms per step")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    out = step()
torch.cuda.synchronize()
print(f"graph {timed(g.replay):.2f} ms per step; equal to eager: {torch.equal(out, ref)}")
