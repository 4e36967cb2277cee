"""Host-side enqueue cost of the launch sequences (no sync inside the timed span)."""
import sys, time
import torch
ROOT = __file__.rsplit("/", 2)[0]
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import fixtures
from crdr_b200 import engine as eng_mod
model, _ = fixtures.build_model(seed=0, calibrated=True)
eng = model.engine()
for b in [int(a) for a in sys.argv[1:]] or (24, 12):
    x = fixtures.image(b, 512, 768).cuda()
    for _ in range(2):
        a = eng.analysis(x, 1.5); eng.decode_device(a["z_sym"], a["y_sym"], 1.5, 3.84, (512, 768))
    torch.cuda.synchronize()
    l0 = eng_mod.LAUNCH_COUNT[0]
    t0 = time.perf_counter(); a = eng.analysis(x, 1.5); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    l1 = eng_mod.LAUNCH_COUNT[0]
    eng.decode_device(a["z_sym"], a["y_sym"], 1.5, 3.84, (512, 768)); t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
    l2 = eng_mod.LAUNCH_COUNT[0]
    print(f"batch {b}: analysis enqueue {1e3*(t1-t0):.1f} ms ({l1-l0} launches) total {1e3*(t2-t0):.1f}; decode enqueue {1e3*(t3-t2):.1f} ms ({l2-l1} launches) total {1e3*(t4-t2):.1f}")
