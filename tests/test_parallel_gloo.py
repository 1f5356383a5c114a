"""world_size-2 gloo tests (CPU) of the N>1 plumbing: clip sharding/gather and the kNN exchange.
The CUDA kernels are replaced by oracle stand-ins here; tests/test_gpu_retrieval.py checks that the
CUDA merge equals the single-shard answer."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rag_gesture_b200.parallel import gather_clips, shard_range, sharded_knn


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 513):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(h - l for l, h in parts) - min(h - l for l, h in parts) <= 1


def _oracle_topk(db, q, k, base):
    s = q @ db.T
    kk = min(k, db.shape[0])
    order = torch.argsort(-s, dim=1, stable=True)[:, :kk]
    idx = torch.full((q.shape[0], k), -1, dtype=torch.int64)
    sc = torch.full((q.shape[0], k), float("-inf"))
    idx[:, :kk] = order + base
    sc[:, :kk] = torch.gather(s, 1, order)
    return idx, sc


def _oracle_merge(idx_parts, score_parts, k):
    P, Q, K = idx_parts.shape
    idx = idx_parts.permute(1, 0, 2).reshape(Q, P * K)
    sc = score_parts.permute(1, 0, 2).reshape(Q, P * K)
    key = torch.where(idx < 0, torch.full_like(sc, float("-inf")), sc)
    out_i, out_s = [], []
    for q in range(Q):
        order = sorted(range(P * K), key=lambda j: (-key[q, j].item(), idx[q, j].item() if idx[q, j] >= 0 else 1 << 62))[:k]
        out_i.append(idx[q, order])
        out_s.append(sc[q, order])
    return torch.stack(out_i), torch.stack(out_s)


def _worker(rank, world, port, n_total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        db = torch.randn(n_total, 64, generator=g)
        q = torch.randn(5, 64, generator=g)
        lo, hi = shard_range(n_total, rank, world)
        idx, sc = sharded_knn(db[lo:hi], q, 4, n_total, local_topk=_oracle_topk, merge=_oracle_merge)
        ref_i, ref_s = _oracle_topk(db, q, 4, 0)
        assert torch.equal(idx, ref_i) and torch.allclose(sc, ref_s)
        # a per-shard index object (KnnIndex on the GPU box) takes the place of the local scan
        class ShardIndex:
            def topk(self, queries, k, idx_base=0):
                return _oracle_topk(db[lo:hi], queries, k, idx_base)
        idx2, sc2 = sharded_knn(db[lo:hi], q, 4, n_total, merge=_oracle_merge, index=ShardIndex())
        assert torch.equal(idx2, ref_i) and torch.allclose(sc2, ref_s)
        # clip shards: each rank "samples" its own clips, the gather restores clip order
        n_clips = 7
        clo, chi = shard_range(n_clips, rank, world)
        local = torch.arange(clo, chi, dtype=torch.float32).view(-1, 1, 1).expand(-1, 3, 2).contiguous()
        full = gather_clips(local, n_clips)
        assert torch.equal(full[:, 0, 0], torch.arange(n_clips, dtype=torch.float32))
    finally:
        dist.destroy_process_group()


def test_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 101), nprocs=2, join=True)
