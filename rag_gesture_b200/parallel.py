"""Multi-GPU partitioning of the path: one process per GPU, torch.distributed for the
plumbing (NCCL over NVLink on the box, gloo in the CPU tests).

  * guided DDIM / inversion: clips are independent units -> contiguous clip shards per rank, NO
    collective inside the loop; an optional all-gather reassembles the output latents;
  * exemplar kNN: database rows sharded, queries replicated, each rank computes its local top-k with
    global indices, ONE all-gather of [Q,k] (index, score) pairs, then a local merge kernel (K11).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced [lo, hi) of n units for `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def gather_clips(local, n_total, group=None):
    """All-gather per-rank clip shards [n_local, ...] back into [n_total, ...] in clip order."""
    rank, world = _world(group)
    if world == 1:
        return local
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)


def _cuda_local_topk(db_shard, queries, k, idx_base):
    from . import _lib
    lib = _lib.load()
    _lib.require_cuda(db_shard, queries)
    q = queries.shape[0]
    idx = torch.empty(q, k, dtype=torch.int64, device=queries.device)
    sc = torch.empty(q, k, device=queries.device)
    with torch.cuda.device(queries.device):
        _lib.check(lib.rg_knn_topk(_lib.ptr(db_shard.contiguous()), db_shard.shape[0], db_shard.shape[1],
                                   _lib.ptr(queries.contiguous()), q, k, idx_base, _lib.ptr(idx),
                                   _lib.ptr(sc), _lib.stream_ptr()))
    return idx, sc


def _cuda_merge(idx_parts, score_parts, k):
    from . import _lib
    lib = _lib.load()
    parts, q, _ = idx_parts.shape
    idx = torch.empty(q, k, dtype=torch.int64, device=idx_parts.device)
    sc = torch.empty(q, k, device=idx_parts.device)
    with torch.cuda.device(idx_parts.device):
        _lib.check(lib.rg_knn_merge(_lib.ptr(idx_parts.contiguous()), _lib.ptr(score_parts.contiguous()),
                                    parts, q, k, _lib.ptr(idx), _lib.ptr(sc), _lib.stream_ptr()))
    return idx, sc


def knn_topk(db, queries, k, idx_base=0):
    """Single-GPU exact fp32 top-k (K10): (idx int64 [Q,k], score fp32 [Q,k]), (score desc, idx asc)."""
    return _cuda_local_topk(db, queries, k, idx_base)


def sharded_knn(db_shard, queries, k, n_total, group=None, local_topk=_cuda_local_topk, merge=_cuda_merge):
    """Top-k over a row-sharded database.  `db_shard` holds rows shard_range(n_total, rank, world);
    `queries` are replicated.  One all-gather of 12*Q*k bytes per rank, then the merge kernel.
    `local_topk` / `merge` are injectable so the CPU (gloo) tests can exercise the exchange."""
    rank, world = _world(group)
    lo, _ = shard_range(n_total, rank, world)
    idx, sc = local_topk(db_shard, queries, k, lo)
    if world == 1:
        return idx, sc
    idx_all = [torch.empty_like(idx) for _ in range(world)]
    sc_all = [torch.empty_like(sc) for _ in range(world)]
    dist.all_gather(idx_all, idx, group=group)
    dist.all_gather(sc_all, sc, group=group)
    return merge(torch.stack(idx_all, 0), torch.stack(sc_all, 0), k)
