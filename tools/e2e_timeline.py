"""Timeline of one pipelined compress_batch / decompress_batch call (host timestamps per chunk thread)."""
import sys, time, threading
import numpy as np
import torch
ROOT = __file__.rsplit("/", 2)[0]
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import fixtures
from crdr_b200 import rans, native as nv
import crdr_b200.model as M

model, _ = fixtures.build_model(seed=0, calibrated=True)
x = fixtures.image(24, 512, 768, seed=100).pin_memory()
for _ in range(3):
    outs = model.compress_batch(x, 1.5)
    model.decompress_batch([o["string_list"] for o in outs], beta=3.84)
torch.cuda.synchronize()

EV = []
T0 = [0.0]
def mark(tag):
    EV.append((time.perf_counter() - T0[0], threading.current_thread().name, tag))

# wrap the host-side stages
_enc, _dec, _chk = rans.encode_batch, rans.decode_batch, nv.status_check
def enc(*a, **k):
    mark("encode>"); r = _enc(*a, **k); mark("encode<"); return r
def dec(*a, **k):
    mark("decode>"); r = _dec(*a, **k); mark("decode<"); return r
def chk(*a, **k):
    mark("sync>"); r = _chk(*a, **k); mark("sync<"); return r
_plan_run = rans.DecodePlan.run
def plan_run(self, *a, **k):
    mark("decode>"); r = _plan_run(self, *a, **k); mark("decode<"); return r
rans.DecodePlan.run = plan_run
rans.encode_batch, rans.decode_batch, nv.status_check = enc, dec, chk
M.rans.encode_batch, M.rans.decode_batch, M.nv.status_check = enc, dec, chk
# the scheduler's waits for a chunk's device -> host copy (model.py _drive): host idle = device (or copy) on the critical path
class _Ev:
    def __init__(self, ev): self.ev = ev
    def synchronize(self):
        mark("wait>"); self.ev.synchronize(); mark("wait<")
_event = M._CharmModelCore._event
M._CharmModelCore._event = staticmethod(lambda: _Ev(_event()))
_an = model.engine().analysis
def an(*a, **k):
    mark("analysis-enqueue>"); r = _an(*a, **k); mark("analysis-enqueue<"); return r
model.engine().analysis = an

torch.cuda.synchronize(); T0[0] = time.perf_counter(); mark("compress>")
outs = model.compress_batch(x, 1.5); torch.cuda.synchronize(); mark("compress<")
t1 = time.perf_counter(); mark("decompress>")
img, _, _ = model.decompress_batch([o["string_list"] for o in outs], beta=3.84); torch.cuda.synchronize(); mark("decompress<")
for t, th, tag in EV:
    print(f"{1e3*t:8.2f} ms  {th:14s} {tag}")
