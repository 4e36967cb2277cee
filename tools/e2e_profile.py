"""Wall-clock split of compress_batch / decompress_batch (host coder vs device vs copies)."""
import sys, time
import numpy as np
import torch
ROOT = __file__.rsplit("/", 2)[0]
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import fixtures
from crdr_b200 import rans

model, _ = fixtures.build_model(seed=0, calibrated=True)
eng = model.engine()
x = fixtures.image(24, 512, 768).pin_memory()
for _ in range(2):
    outs = model.compress_batch(x, 1.5)
    model.decompress_batch([o["string_list"] for o in outs], beta=3.84)
torch.cuda.synchronize()
def T(): torch.cuda.synchronize(); return time.perf_counter()
t0 = T(); xd = x.to("cuda:0"); t1 = T()
a = eng.analysis(xd, 1.5); t2 = T()
zs, ys, yi = a["z_sym"].cpu().numpy(), a["y_sym"].cpu().numpy(), a["y_idx"].cpu().numpy(); t3 = T()
from crdr_b200.model import _channel_indexes
zi = _channel_indexes(192, 8, 12)
zt, yt = model.entropy_model_z.coder_tables(), model.entropy_model_y.coder_tables()
z_strs = rans.encode_batch([zs[i] for i in range(24)], [zi] * 24, zt); t4 = T()
y_strs = rans.encode_batch([ys[i] for i in range(24)], [yi[i] for i in range(24)], yt); t5 = T()
print(f"H2D image {1e3*(t1-t0):.1f} ms | analysis {1e3*(t2-t1):.1f} | D2H sym/idx {1e3*(t3-t2):.1f} | z enc {1e3*(t4-t3):.1f} | y enc {1e3*(t5-t4):.1f}")
decs = [rans.Decoder(s) for s in y_strs]
t6 = T(); out = rans.decode_batch(decs, [yi[i] for i in range(24)], yt); t7 = T()
print(f"y decode all (host only, 24 images) {1e3*(t7-t6):.1f} ms; equal {all(np.array_equal(o, ys[i].reshape(-1)) for i, o in enumerate(out))}")
t8 = T(); outs = model.compress_batch(x, 1.5); t9 = T()
img, _, _ = model.decompress_batch([o["string_list"] for o in outs], beta=3.84); t10 = T()
print(f"compress_batch {1e3*(t9-t8):.1f} ms, decompress_batch {1e3*(t10-t9):.1f} ms; bytes/img {sum(len(s) for s in outs[0]['string_list'])}")
