"""Host range coder micro-benchmark (CPU only): time per stream and ns per symbol of encode_batch / decode_batch on the
Gaussian tables.  usage: coder_bench.py [streams] [threads] [--latents FILE.npz [--q Q]]
Without --latents the symbols are synthetic (a latent-like index distribution, about one bit per symbol); with it they are
real coder inputs dumped on a GPU box by tools/dump_latents.py.  CRDR_CODER_INTERLEAVE forces the bundle size."""
import argparse, sys, time
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from crdr_b200 import rans
from crdr_b200.entropy import GaussianMeanScaleConditional, get_scale_table

ap = argparse.ArgumentParser()
ap.add_argument("streams", type=int, nargs="?", default=8)
ap.add_argument("threads", type=int, nargs="?", default=0)
ap.add_argument("--latents")
ap.add_argument("--q", default="1.5")
ap.add_argument("--reps", type=int, default=15)
args = ap.parse_args()
gc = GaussianMeanScaleConditional(scale_bound=0.11)
gc.update_scale_table(get_scale_table(), force=True)
tabs = gc.coder_tables()
nc = tabs.cdfs.shape[0]
rng = np.random.default_rng(0)
cnt, thr = args.streams, args.threads
syms, idxs = [], []
if args.latents:
    d = np.load(args.latents)
    ys, yi = d[f"y_sym_q{args.q}"], d[f"y_idx_q{args.q}"]
    for k in range(cnt):
        syms.append(np.ascontiguousarray(ys[k % ys.shape[0]]).reshape(-1))
        idxs.append(np.ascontiguousarray(yi[k % yi.shape[0]]).reshape(-1))
    N = syms[0].size
    sizes, offs = tabs.sizes[idxs[0]], tabs.offsets[idxs[0]]
    v = syms[0].astype(np.int64) - offs
    print("real latents: %d symbols per stream, escapes %.4f, mean index %.1f" % (N, float(((v < 0) | (v >= sizes - 2)).mean()), idxs[0].mean()))
else:
    N = 491520
    table = get_scale_table().numpy()
    for k in range(cnt):
        u = rng.random(N)
        ix = np.where(u < 0.7, rng.integers(0, 6, N), np.where(u < 0.95, rng.integers(5, 30, N), rng.integers(30, nc, N))).astype(np.uint8)
        syms.append(np.rint(rng.standard_normal(N) * table[ix]).astype(np.int16))
        idxs.append(ix)
te = 1e9
for rep in range(args.reps):
    t0 = time.perf_counter(); strs = rans.encode_batch(syms, idxs, tabs, thr); te = min(te, time.perf_counter() - t0)
print("encode: %.2f ms for %d x %d symbols -> %.2f ns/sym of thread time (%d threads), %.1f bits/sym" %
      (te * 1e3, cnt, N, te * 1e9 * min(cnt, rans.pool_info()[0] if thr == 0 else thr) / N / cnt, min(cnt, rans.pool_info()[0] if thr == 0 else thr), 8.0 * len(strs[0]) / N))
td = 1e9
for rep in range(args.reps):
    decs = [rans.Decoder(s) for s in strs]
    outs = [np.zeros(N, np.int32) for _ in range(cnt)]
    for o in outs:
        o.fill(1)
    t0 = time.perf_counter(); rans.decode_batch(decs, idxs, tabs, thr, outs=outs); td = min(td, time.perf_counter() - t0)
print("decode: %.2f ms -> %.2f ns/sym of thread time" % (td * 1e3, td * 1e9 * min(cnt, rans.pool_info()[0] if thr == 0 else thr) / N / cnt))
assert all((o == s).all() for o, s in zip(outs, syms))
