"""In-kernel event timeline of CTA 0: needs a library built with CRDR_BUILD_TRACE=1 python -m crdr_b200.build.

    python tools/conv_events.py <conv_bench args>     (EV_FROM / EV_TO select the printed records)"""
import ctypes, os, sys
os.environ["CRDR_CONV_TRACE"] = "2"
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0]); sys.path.insert(0, __file__.rsplit("/", 1)[0])
from crdr_b200 import native as nv
import conv_bench
TAGS = {1: "epi  tile start", 2: "epi  acc_full seen", 3: "epi  chunks done", 4: "epi  copy-out done", 8: "mma  acc buffer free",
        9: "mma  tile issued", 16: "epi   c0 tmem ld issued", 17: "epi   c0 res read+refill", 18: "epi   c0 acc in regs", 19: "epi   c0 finish done", 10: "mma  patch landed", 12: "patch TMA issued"}
a = [int(v) for v in sys.argv[1:]]
conv_bench.bench(*a, iters=1)
buf = (ctypes.c_uint32 * 16384)()
nv.lib().crdr_debug_events(buf, 16384)
w = np.frombuffer(buf, dtype=np.uint32)
rec = w[:2 * 3 * 2600].reshape(-1, 2)
rec = rec[rec[:, 0] != 0]
order = np.argsort(rec[:, 1].astype(np.int64), kind="stable")
rec = rec[order]
t0 = int(rec[0, 1])
lo, hi = (int(os.environ.get("EV_FROM", 200)), int(os.environ.get("EV_TO", 290)))
for tagw, clk in rec[lo:hi]:
    tag, pay = int(tagw) >> 24, int(tagw) & 0xFFFFFF
    print(f"{(int(clk) - t0) & 0xFFFFFFFF:9d} cyc  {TAGS.get(tag, tag):22s} {pay}")
