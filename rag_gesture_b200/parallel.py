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


def _check_knn_inputs(db_shard, queries):
    """The kernels read raw fp32 rows: refuse anything else instead of reading out of bounds."""
    if db_shard.dtype != torch.float32 or queries.dtype != torch.float32:
        raise ValueError(f"kNN operands must be float32 (got {db_shard.dtype}, {queries.dtype})")
    if db_shard.dim() != 2 or queries.dim() != 2 or db_shard.shape[1] != queries.shape[1]:
        raise ValueError(f"kNN operands must be [n, dim] and [Q, dim] (got {tuple(db_shard.shape)}, {tuple(queries.shape)})")


def _out_pair(q, k, device, out):
    if out is None:
        return torch.empty(q, k, dtype=torch.int64, device=device), torch.empty(q, k, device=device)
    idx, sc = out
    assert idx.dtype == torch.int64 and sc.dtype == torch.float32 and idx.is_contiguous() and sc.is_contiguous()
    assert tuple(idx.shape) == (q, k) == tuple(sc.shape)
    return idx, sc


def _cuda_local_topk(db_shard, queries, k, idx_base, out=None):
    from . import _lib
    lib = _lib.load()
    _lib.require_cuda(db_shard, queries)
    _check_knn_inputs(db_shard, queries)
    q = queries.shape[0]
    idx, sc = _out_pair(q, k, queries.device, out)
    with torch.cuda.device(queries.device):
        _lib.check(lib.rg_knn_topk(_lib.ptr(db_shard.contiguous()), db_shard.shape[0], db_shard.shape[1],
                                   _lib.ptr(queries.contiguous()), q, k, idx_base, _lib.ptr(idx),
                                   _lib.ptr(sc), _lib.stream_ptr()))
    return idx, sc


def packed_stride(q, k):
    """Bytes per rank in the packed exchange buffer: [idx int64 q*k | score fp32 q*k], rounded up to 16."""
    return (12 * q * k + 15) // 16 * 16


def packed_views(buf, q, k):
    """(idx int64 [q,k], score fp32 [q,k]) views into one rank's block of the packed buffer (uint8)."""
    n = q * k
    return buf[:8 * n].view(torch.int64).view(q, k), buf[8 * n:12 * n].view(torch.float32).view(q, k)


def _cuda_merge_packed(recv, world, q, k):
    from . import _lib
    lib = _lib.load()
    idx = torch.empty(q, k, dtype=torch.int64, device=recv.device)
    sc = torch.empty(q, k, device=recv.device)
    with torch.cuda.device(recv.device):
        _lib.check(lib.rg_knn_merge_packed(_lib.ptr(recv), packed_stride(q, k), world, q, k, _lib.ptr(idx),
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

    def topk(self, queries, k, idx_base=0, out=None):
        """(idx int64 [Q,k], score fp32 [Q,k]) == knn_topk(self.db, queries, k, idx_base), bit for bit.
        `out` = (idx, score) tensors to write into (the packed exchange buffer of sharded_knn)."""
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
            return _cuda_local_topk(self.db, queries, k, idx_base, out=out)
        idx, sc = _out_pair(q, k, queries.device, out)
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


def sharded_knn(db_shard, queries, k, n_total, group=None, local_topk=None, merge=None, index=None):
    """Top-k over a row-sharded database.  `db_shard` holds rows shard_range(n_total, rank, world);
    `queries` are replicated.  Each rank writes its local top-k (global indices) straight into its block of a
    packed buffer [idx int64 Q*k | score fp32 Q*k]; ONE all_gather_into_tensor (12*Q*k bytes per rank) moves
    all of them, and the merge kernel (rg_knn_merge_packed) reads the gathered blocks in place.
    `local_topk(db_shard, queries, k, idx_base)` / `merge(idx_parts, score_parts, k)` are injectable so the CPU
    (gloo) tests can exercise the same exchange; `index` (a KnnIndex of this rank's shard) selects the
    tensor-core local top-k for large query batches."""
    rank, world = _world(group)
    lo, _ = shard_range(n_total, rank, world)
    q = queries.shape[0]
    if world == 1:
        if index is not None:
            return index.topk(queries, k, lo)
        return (local_topk or _cuda_local_topk)(db_shard, queries, k, lo)
    stride = packed_stride(q, k)
    send = torch.zeros(stride, dtype=torch.uint8, device=queries.device)
    idx, sc = packed_views(send, q, k)
    if index is not None and local_topk is None and hasattr(index, "handle"):
        index.topk(queries, k, lo, out=(idx, sc))
    elif local_topk is None and index is None:
        _cuda_local_topk(db_shard, queries, k, lo, out=(idx, sc))
    else:                                       # injected stand-ins return fresh tensors
        i, s = index.topk(queries, k, lo) if index is not None else local_topk(db_shard, queries, k, lo)
        idx.copy_(i)
        sc.copy_(s)
    recv = torch.empty(world * stride, dtype=torch.uint8, device=queries.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    if merge is None and recv.is_cuda:
        return _cuda_merge_packed(recv, world, q, k)
    parts = [packed_views(recv[r * stride:(r + 1) * stride], q, k) for r in range(world)]
    idx_parts = torch.stack([p[0] for p in parts], 0)
    score_parts = torch.stack([p[1] for p in parts], 0)
    return (merge or _cuda_merge)(idx_parts, score_parts, k)
