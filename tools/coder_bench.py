"""Host range coder micro-benchmark (CPU only): ns per symbol of encode_batch / decode_batch on Gaussian tables with a\nlatent-like index distribution.  usage: coder_bench.py [streams] [threads]; CRDR_CODER_INTERLEAVE forces the bundle size."""
import sys, time, numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from crdr_b200 import rans
from crdr_b200.entropy import GaussianMeanScaleConditional, get_scale_table
import torch
gc = GaussianMeanScaleConditional(scale_bound=0.11)
gc.update_scale_table(get_scale_table(), force=True)
tabs = gc.coder_tables()
nc = tabs.cdfs.shape[0]
print("n_cdf", nc, "stride", tabs.cdfs.shape[1])
rng = np.random.default_rng(0)
N = 491520
cnt = int(sys.argv[1]) if len(sys.argv) > 1 else 8
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 0
table = get_scale_table().numpy()
syms, idxs = [], []
for k in range(cnt):
    u = rng.random(N)
    ix = np.where(u < 0.7, rng.integers(0, 6, N), np.where(u < 0.95, rng.integers(5, 30, N), rng.integers(30, nc, N))).astype(np.uint8)
    s = np.rint(rng.standard_normal(N) * table[ix]).astype(np.int16)
    syms.append(s); idxs.append(ix)
te=1e9
for rep in range(25):
    t0 = time.perf_counter(); strs = rans.encode_batch(syms, idxs, tabs, thr); t1 = time.perf_counter(); te=min(te,t1-t0)
print("encode: %.2f ms for %d x %d  -> %.2f ns/sym (per thread-time), bytes/img %d" % (te*1e3, cnt, N, te*1e9/N/cnt, len(strs[0])))
td=1e9
for rep in range(25):
    decs = [rans.Decoder(s) for s in strs]
    outs = [np.zeros(N, np.int32) for _ in range(cnt)]
    for o in outs: o.fill(1)
    t0 = time.perf_counter(); rans.decode_batch(decs, idxs, tabs, thr, outs=outs); t1 = time.perf_counter(); td=min(td,t1-t0)
print("decode: %.2f ms -> %.2f ns/sym (per thread-time)" % (td*1e3, td*1e9/N/cnt))
assert all((o == s).all() for o, s in zip(outs, syms))
print(rans.pool_info())
