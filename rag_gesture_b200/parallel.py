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


class KnnIndex:
    """Tensor-core index of one database shard (rg_knn_index_create): a bf16 copy of the rows as the TMA
    operand of the similarity GEMM plus the largest row norm for the error-bound certificate.  Keeps a
    reference to the fp32 shard: the exact re-score reads it, so results stay bit-identical to the scan.
    Queries at or below `min_queries` go through the exact scan (one HBM pass serves 8 queries)."""
    min_queries = 8

    def __init__(self, db_shard):
        import ctypes
        from . import _lib
        self._lib = _lib
        lib = _lib.load()
        _lib.require_cuda(db_shard)
        self.db = db_shard.contiguous()
        if self.db.dtype != torch.float32 or self.db.dim() != 2:
            raise ValueError("KnnIndex: the shard must be a [n, dim] float32 tensor")
        self.handle = ctypes.c_void_p()
        self.last_uncertified = 0
        with torch.cuda.device(self.db.device):
            _lib.check(lib.rg_knn_index_create(_lib.ptr(self.db), self.db.shape[0], self.db.shape[1],
                                               ctypes.byref(self.handle), _lib.stream_ptr()))

    def topk(self, queries, k, idx_base=0):
        """(idx int64 [Q,k], score fp32 [Q,k]) == knn_topk(self.db, queries, k, idx_base), bit for bit."""
        import ctypes
        _lib = self._lib
        _lib.require_cuda(queries)
        if self.handle is None:
            raise RuntimeError("KnnIndex: already closed")
        if queries.dim() != 2 or queries.shape[1] != self.db.shape[1]:
            raise ValueError(f"KnnIndex.topk: queries must be [Q, {self.db.shape[1]}]")
        queries = queries.to(torch.float32).contiguous()
        q = queries.shape[0]
        if q <= self.min_queries:
            return _cuda_local_topk(self.db, queries, k, idx_base)
        idx = torch.empty(q, k, dtype=torch.int64, device=queries.device)
        sc = torch.empty(q, k, device=queries.device)
        nf = ctypes.c_int32(0)
        with torch.cuda.device(queries.device):
            _lib.check(_lib.load().rg_knn_topk_tc(self.handle, _lib.ptr(self.db), _lib.ptr(queries), q, k, idx_base,
                                                  _lib.ptr(idx), _lib.ptr(sc), ctypes.byref(nf), _lib.stream_ptr()))
        self.last_uncertified = nf.value
        return idx, sc

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self._lib.load().rg_knn_index_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def knn_topk(db, queries, k, idx_base=0, index=None):
    """Single-GPU exact fp32 top-k (K10): (idx int64 [Q,k], score fp32 [Q,k]), (score desc, idx asc).
    With `index` (a KnnIndex of `db`) large query batches take the tensor-core path; same results."""
    if index is not None:
        return index.topk(queries, k, idx_base)
    return _cuda_local_topk(db, queries, k, idx_base)


def sharded_knn(db_shard, queries, k, n_total, group=None, local_topk=_cuda_local_topk, merge=_cuda_merge,
                index=None):
    """Top-k over a row-sharded database.  `db_shard` holds rows shard_range(n_total, rank, world);
    `queries` are replicated.  One all-gather of 12*Q*k bytes per rank, then the merge kernel.
    `local_topk` / `merge` are injectable so the CPU (gloo) tests can exercise the exchange; `index`
    (a KnnIndex of this rank's shard) selects the tensor-core local top-k for large query batches."""
    rank, world = _world(group)
    lo, _ = shard_range(n_total, rank, world)
    idx, sc = index.topk(queries, k, lo) if index is not None else local_topk(db_shard, queries, k, lo)
    if world == 1:
        return idx, sc
    idx_all = [torch.empty_like(idx) for _ in range(world)]
    sc_all = [torch.empty_like(sc) for _ in range(world)]
    dist.all_gather(idx_all, idx, group=group)
    dist.all_gather(sc_all, sc, group=group)
    return merge(torch.stack(idx_all, 0), torch.stack(sc_all, 0), k)
