"""Device-timed encode+decode MPix/s at BASELINE's other image shapes (not bench lines; bench.py covers configs[1])."""
import sys
import torch
ROOT = __file__.rsplit("/", 2)[0]
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import fixtures
model, _ = fixtures.build_model(seed=0, calibrated=True)
eng = model.engine()
for n, h, w in ((24, 512, 768), (4, 1365, 2048), (1, 1365, 2048), (1, 2160, 3840), (8, 256, 256)):
    x = fixtures.image(n, h, w, seed=h).cuda()
    def step():
        a = eng.analysis(x, 2.0)
        return eng.decode_device(a["z_sym"], a["y_sym"], 2.0, 3.84, (h, w))[0]
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{n:3d} x {h}x{w}: {ms:8.2f} ms per step  {n * h * w / ms / 1e3:7.1f} MPix/s (device-timed encode+decode)")
