"""One warm-up + one encode+decode step (for ncu launch lists / captures; never a timing source)."""
import sys
import torch
ROOT = __file__.rsplit("/", 2)[0]
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import fixtures  # noqa: E402
b = int(sys.argv[1]) if len(sys.argv) > 1 else 24
model, _ = fixtures.build_model(seed=0, calibrated=True)
eng = model.engine()
x = fixtures.image(b, 512, 768).cuda()
for _ in range(2):
    a = eng.analysis(x, 1.5)
    eng.decode_device(a["z_sym"], a["y_sym"], 1.5, 3.84, (512, 768))
    torch.cuda.synchronize()
print("done")
