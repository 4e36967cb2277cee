"""Pipeline shape sweep of the codec API in one process (host-side diagnosis; never a bench number):
wall-clock of compress_batch and decompress_batch of the default workload for several chunk counts / weights."""
import sys, time
import numpy as np
import torch
ROOT = __file__.rsplit("/", 2)[0]
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import fixtures

model, _ = fixtures.build_model(seed=0, calibrated=True)
x = (fixtures.image(24, 512, 768, seed=100) * 127.5 + 127.5).clamp(0, 255).to(torch.uint8).pin_memory()
def T():
    torch.cuda.synchronize(); return time.perf_counter()

def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        r = fn()
    ts = []
    for _ in range(reps):
        t0 = T(); r = fn(); ts.append(1e3 * (T() - t0))
    return r, float(np.median(ts)), float(np.min(ts))

outs = model.compress_batch(x, 1.5)
streams = [o["string_list"] for o in outs]
print("decompress_batch (24 x 512x768, uint8 out): chunks / weights -> median, min ms")
for weights in [(2, 1), (1, 1), (3, 2), (3, 2, 1), (2, 2, 1), (1, 1, 1), (4, 2, 1), (1, 1, 1, 1)]:
    model.pipeline_chunks, model.pipeline_weights = len(weights), tuple(float(w) for w in weights)
    _, med, mn = timed(lambda: model.decompress_batch(streams, beta=3.84, out_uint8=True))
    print(f"  {str(weights):16s} {med:6.1f} {mn:6.1f}", flush=True)
    model.engine().decode_graph_sets.clear()
print("compress_batch: chunks / weights -> median, min ms")
for weights in [(1, 1), (2, 1), (3, 1), (4, 1), (1, 1, 1), (2, 2, 1), (3, 2, 1), (1, 1, 1, 1), (3, 3, 2, 1), (6, 1, 1)]:
    model.pipeline_chunks_compress, model.pipeline_weights_compress = len(weights), tuple(float(w) for w in weights)
    _, med, mn = timed(lambda: model.compress_batch(x, 1.5))
    print(f"  {str(weights):16s} {med:6.1f} {mn:6.1f}", flush=True)
