"""Multi-process (gloo, world_size 2) test of the image sharding / result merge used by scripts/compress.py and
bench.py -- the only multi-GPU logic of the inference path (no data-path collective)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from crdr_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names = [f"img{i:02d}.png" for i in range(7)]
    mine = sharding.shard(names, *sharding.rank_world())
    rows = [{"img_name": n, "real_bpp": float(int(n[3:5]))} for n in mine]
    merged = sharding.gather_rows(rows, rank, world)
    # the max-over-ranks timing reduction bench.py uses
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        assert sorted(r["img_name"] for r in merged) == names
        assert abs(sum(r["real_bpp"] for r in merged) / 7 - 3.0) < 1e-12
        assert t.item() == 10.0 + world - 1
        open(os.path.join(tmp, "ok"), "w").write("ok")
    else:
        assert merged == []
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def _train_worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from crdr_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # flat gradient buffer: several buckets, the last one partial
    n = 100_003
    g = torch.arange(n, dtype=torch.float32) * (rank + 1)
    buckets = sharding.allreduce_mean_flat(g, bucket_bytes=64 * 1024)
    want = torch.arange(n, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    assert buckets == -(-n * 4 // (64 * 1024)) and torch.allclose(g, want, rtol=1e-6)
    # the two-phase form the training step uses: some ranges start early (overlapping the rest of the backward), the rest later
    g2 = torch.arange(n, dtype=torch.float32) * (rank + 1)
    early = sharding.allreduce_sum_async(g2, [(1000, 40_000), (60_000, n)], bucket_bytes=64 * 1024)
    late = sharding.allreduce_sum_async(g2, [(0, 1000), (40_000, 60_000)], bucket_bytes=64 * 1024)
    sharding.allreduce_finish_mean(early + late, g2)
    assert torch.allclose(g2, want, rtol=1e-6)
    q = torch.tensor([3.0 if rank == 0 else -1.0])
    assert sharding.broadcast_from_rank0(q).item() == 3.0
    qbpp = torch.tensor([0.1 * (rank + 1)])
    assert abs(sharding.allreduce_mean_scalar(qbpp).item() - 0.1 * (world + 1) / 2) < 1e-7
    if rank == 0:
        open(os.path.join(tmp, "ok_train"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_training_exchanges_world2(tmp_path):
    """The data-parallel exchanges of the training step over gloo: bucketed gradient mean, the rank-0
    quality level, the qbpp mean of the rate switch."""
    port = 31500 + os.getpid() % 2000
    mp.spawn(_train_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok_train").exists()


def test_training_exchanges_single_process_are_noops():
    from crdr_b200 import sharding
    g = torch.ones(10)
    assert sharding.allreduce_mean_flat(g) == 0 and torch.equal(g, torch.ones(10))
    assert sharding.broadcast_from_rank0(torch.tensor([2.0])).item() == 2.0


def test_shard_partition_properties():
    from crdr_b200 import sharding
    items = list(range(23))
    for world in (1, 2, 4, 8):
        parts = [sharding.shard(items, r, world) for r in range(world)]
        assert sorted(x for p in parts for x in p) == items
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    a, b = torch.zeros(1, 3, 8, 8), torch.zeros(1, 3, 8, 16)
    groups = list(sharding.same_shape_batches([("a", a), ("b", a), ("c", a), ("d", b), ("e", a)], 2))
    assert [[n for n, _ in g] for g in groups] == [["a", "b"], ["c"], ["d"], ["e"]]


def test_shard_balanced_by_padded_area():
    from crdr_b200 import sharding
    names = [f"i{k:02d}.png" for k in range(12)]
    sizes = [(2160, 3840)] * 2 + [(1365, 2048)] * 4 + [(512, 768)] * 6
    import random
    perm = list(range(12))
    random.Random(3).shuffle(perm)
    names_p, sizes_p = [names[i] for i in perm], [sizes[i] for i in perm]
    pad = lambda v: -(-v // 64) * 64
    area = {n: pad(s[0]) * pad(s[1]) for n, s in zip(names, sizes)}
    for world in (1, 2, 4):
        parts = [sharding.shard_balanced(names_p, sizes_p, r, world) for r in range(world)]
        assert sorted(x for p in parts for x in p) == names                     # a partition
        loads = [sum(area[n] for n in p) for p in parts]
        assert max(loads) - min(loads) <= max(area.values())                    # balanced to within one image
        for p in parts:                                                         # equal shapes adjacent on every rank
            shapes = [sizes[names.index(n)] for n in p]
            assert shapes == sorted(shapes, key=lambda s: -pad(s[0]) * pad(s[1]))


def test_gather_rows_file_fallback(tmp_path):
    from crdr_b200 import sharding
    import threading
    out = {}
    th = threading.Thread(target=lambda: out.setdefault(1, sharding.gather_rows([{"k": 1}], 1, 2, str(tmp_path))))
    th.start()
    out[0] = sharding.gather_rows([{"k": 0}], 0, 2, str(tmp_path))
    th.join()
    assert sorted(r["k"] for r in out[0]) == [0, 1] and out[1] == []


def test_stale_rank_files_are_not_merged(tmp_path):
    """a _rows_rank*.json left by an earlier crashed run must not end up in _bitrates.csv."""
    import json
    import os
    import time
    import pytest
    from crdr_b200 import sharding
    stale = tmp_path / "_rows_rank1.json"
    stale.write_text(json.dumps([{"img_name": "old.png"}]))
    old = time.time() - 3600
    os.utime(stale, (old, old))
    with pytest.raises(TimeoutError):
        sharding.gather_rows([{"img_name": "a.png"}], 0, 2, str(tmp_path), timeout_s=0.3)
    stale.write_text(json.dumps([{"img_name": "b.png"}]))          # a fresh delivery of rank 1
    rows = sharding.gather_rows([{"img_name": "a.png"}], 0, 2, str(tmp_path), timeout_s=5.0)
    assert sorted(r["img_name"] for r in rows) == ["a.png", "b.png"]


def test_coder_pool_takes_the_ranks_slice_of_the_host():
    """One process per GPU: the coder pool of rank r lives on the r-th of LOCAL_WORLD_SIZE equal slices of the CPUs the
    process may use (rans.cpp Pool), interleaves streams when they outnumber its threads, and still returns the bytes of a
    single-threaded run."""
    import subprocess
    import sys
    import pytest
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if ncpu < 4:
        pytest.skip("needs at least 4 CPUs")
    code = r"""
import sys, hashlib, numpy as np
sys.path.insert(0, %r)
from crdr_b200 import rans
pmf = np.exp(-0.5 * ((np.arange(41) - 20) / 4.0) ** 2).astype(np.float32); pmf /= pmf.sum()
cdf = rans.pmf_to_quantized_cdf(pmf)
t = rans.Tables(np.stack([cdf] * 3), [cdf.size] * 3, [-20] * 3)
rng = np.random.default_rng(5)
syms = [np.rint(rng.normal(0, 5, 4000)).astype(np.int32) for _ in range(11)]
idxs = [rng.integers(0, 3, 4000).astype(np.int32) for _ in range(11)]
out = rans.encode_batch(syms, idxs, t)
back = rans.decode_batch([rans.Decoder(s) for s in out], idxs, t)
assert all(np.array_equal(a, b) for a, b in zip(back, syms))
print(rans.pool_info()[0], rans.pool_info()[1], hashlib.sha1(b"".join(out)).hexdigest())
""" % ROOT
    cpus = sorted(os.sched_getaffinity(0))
    lines = []
    for env in ({}, {"LOCAL_WORLD_SIZE": "4", "LOCAL_RANK": "1"}, {"LOCAL_WORLD_SIZE": "2", "LOCAL_RANK": "1", "CRDR_CODER_THREADS": "1"}):
        e = {k: v for k, v in os.environ.items() if k not in ("LOCAL_WORLD_SIZE", "LOCAL_RANK", "WORLD_SIZE", "RANK", "CRDR_CODER_THREADS")}
        e.update(env)
        r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        lines.append(r.stdout.split())
    (t0, c0, h0), (t1, c1, h1), (t2, c2, h2) = lines
    assert int(t0) == min(ncpu, 32) and int(c0) == cpus[0]
    assert int(t1) == ncpu // 4 and int(c1) == cpus[ncpu // 4]            # second slice of four
    assert int(t2) == 1 and int(c2) == cpus[ncpu // 2]
    assert h0 == h1 == h2                                                  # same bytes whatever the pool
