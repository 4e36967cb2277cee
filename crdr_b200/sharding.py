"""Image-level sharding across GPUs: images are independent units, so rank r of W simply takes items
r, r+W, ... of the sorted list; there is no collective on the data path.  Per-image result rows are merged on
rank 0 -- through torch.distributed when a process group exists, else through per-rank JSON files.

Training (third row) is data parallel: identical replicas, rank-local crops, and three small exchanges per
step -- the quality level drawn by rank 0 (the reference draws one per batch), the quantised-bpp mean behind the HiFiC
rate switch, and the bucketed mean all-reduce of the flat gradient buffer (NCCL over NVLink on the GPUs; the same code
runs over gloo in the CPU tests)."""
import json
import os
import time

_PROCESS_START = time.time()


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard(items, rank, world):
    """Round-robin shard of an (already sorted) list; every item lands on exactly one rank."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(items[rank::world])


def shard_balanced(items, sizes, rank, world, stride=64):
    """"sort by padded area for balance".  items with their (height, width): sorted by padded area (descending),
    then shape, then name; rank r takes positions r, r+W, ... of that order, so every rank gets the same mix of large and
    small images and same-shaped images stay adjacent (they are batched into one launch).  Deterministic on every rank."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    pad = lambda v: -(-v // stride) * stride
    order = sorted(range(len(items)), key=lambda i: (-pad(sizes[i][0]) * pad(sizes[i][1]), sizes[i], str(items[i])))
    return [items[i] for i in order[rank::world]]


def same_shape_batches(named_images, batch):
    """Group consecutive (name, tensor[1,3,H,W]) pairs of identical shape into lists of at most `batch`."""
    group = []
    for item in named_images:
        if group and (len(group) >= max(1, batch) or item[1].shape != group[0][1].shape):
            yield group
            group = []
        group.append(item)
    if group:
        yield group


def gather_rows(rows, rank, world, scratch_dir=None, timeout_s=600.0):
    """All ranks' row lists concatenated on rank 0 (other ranks get [])."""
    if world == 1:
        return list(rows)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            bucket = [None] * world if rank == 0 else None
            dist.gather_object(rows, bucket, dst=0)
            return [r for part in bucket for r in part] if rank == 0 else []
    except ImportError:
        pass
    if scratch_dir is None:
        raise RuntimeError("gather_rows needs a process group or a scratch directory")
    mine = os.path.join(scratch_dir, f"_rows_rank{rank}.json")
    with open(mine + ".tmp", "w") as f:
        json.dump(rows, f)
    os.replace(mine + ".tmp", mine)
    if rank != 0:
        return []
    # file fallback (no process group): a file left behind by an earlier, crashed run must not be merged -- only files
    # written after this job started count (all ranks of a job start together on one host)
    fresh = lambda p: os.path.exists(p) and os.path.getmtime(p) >= _PROCESS_START - 5.0
    out, deadline = [], time.time() + timeout_s
    for r in range(world):
        path = os.path.join(scratch_dir, f"_rows_rank{r}.json")
        while not fresh(path):
            if time.time() > deadline:
                raise TimeoutError(f"rank {r} never delivered its rows")
            time.sleep(0.05)
        with open(path) as f:
            out.extend(json.load(f))
        os.remove(path)
    return out


def _group_world(group=None):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None, 1
    return dist, dist.get_world_size(group)


def allreduce_mean_flat(flat, bucket_bytes=64 << 20, group=None):
    """In-place mean over the ranks of a flat (1-D, contiguous) gradient buffer, in fixed-size buckets so that the
    first buckets are on the wire while the later ones are still being enqueued.  Returns the number of buckets
    (0: single process, nothing done)."""
    dist, world = _group_world(group)
    if world == 1:
        return 0
    assert flat.dim() == 1 and flat.is_contiguous()
    per = max(1, bucket_bytes // flat.element_size())
    handles = [dist.all_reduce(flat[o:o + per], op=dist.ReduceOp.SUM, group=group, async_op=True)
               for o in range(0, flat.numel(), per)]
    for h in handles:
        h.wait()
    flat.mul_(1.0 / world)
    return len(handles)


def allreduce_sum_async(flat, ranges, bucket_bytes=64 << 20, group=None):
    """Start the bucketed sum over the ranks of the element ranges [(lo, hi), ...] of a flat buffer and return the work
    handles: the collective runs on the backend's own stream, ordered after what the current stream has enqueued so far,
    so kernels launched afterwards (the rest of the backward pass) overlap it.  [] for a single process."""
    dist, world = _group_world(group)
    if world == 1:
        return []
    per = max(1, bucket_bytes // flat.element_size())
    return [dist.all_reduce(flat[o:min(o + per, hi)], op=dist.ReduceOp.SUM, group=group, async_op=True)
            for lo, hi in ranges for o in range(lo, hi, per)]


def allreduce_finish_mean(handles, flat, group=None):
    """Wait for the handles of allreduce_sum_async (together covering the whole buffer) and turn the sums into means."""
    dist, world = _group_world(group)
    if world == 1:
        return
    for h in handles:
        h.wait()
    flat.mul_(1.0 / world)


def allreduce_mean_scalar(t, group=None):
    """Mean over the ranks of a small tensor, in place and without a host round trip (the HiFiC rate switch's qbpp)."""
    dist, world = _group_world(group)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.mul_(1.0 / world)
    return t


def broadcast_from_rank0(t, group=None):
    """Rank 0's tensor on every rank (the per-step quality level / beta the reference draws once per batch)."""
    dist, world = _group_world(group)
    if world > 1:
        dist.broadcast(t, src=0, group=group)
    return t
