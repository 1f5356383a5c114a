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


class _FakeArch:
    """CPU stand-in with the MotionDiffusion methods LongformSynthesizer.run_sharded drives; every result is a
    deterministic function of the window index and the previous latent, so any mistake in sharding, payload
    order or shipping changes the output."""

    class _Diff:
        num_timesteps = 4

    diffusion_test = _Diff()

    def prepare(self, **batch):
        from rag_gesture_b200.architecture import GuidedBatch
        c = float(batch["cidx"])
        gb = GuidedBatch()
        gb.use_outpaint, gb.use_inversion, gb.inversion_start_time, gb.use_guidance = False, True, -1, True
        gb.guidance_iters, gb.guidance_lr, gb.shape, gb.device, gb.extra = [0] * 4, 0.1, (1, 7, 8), torch.device("cpu"), {}
        gb.jobs, gb.c = ((0, "w"),), c
        xf = {k: torch.full((1, 3, 8), c + i) for i, k in enumerate(("xf_text", "xf_audio", "xf_spk"))}
        gb.model_kwargs = dict(xf_out=xf, query_mask={"q": torch.ones(1, 7)}, motion_mask=torch.full((1, 7), c + 1))
        return gb

    def invert_many(self, gbs):
        for gb in gbs:
            gb.inv = torch.full((4, 1, 7, 8), gb.c)

    def encode_clip_conditions(self, gb):       # prepare() already encoded them (defer_conditions=False)
        return gb

    def insertion_targets(self, gb):
        rows = gb.inv[-1].clone()
        mask = torch.zeros(1, 7, 8, dtype=torch.bool)
        mask[:, 2:4] = True
        return rows, mask, gb.inv.clone()

    def mask_prev_latent(self, prev):
        return None if prev is None else prev * 0.5

    def run_prepared(self, gb):
        rows, mask, inv = gb.targets if gb.targets is not None else self.insertion_targets(gb)
        x = torch.where(mask, rows, torch.zeros_like(rows)) + inv.sum(0) + gb.model_kwargs["xf_out"]["xf_audio"].mean()
        x = x + gb.model_kwargs["motion_mask"].unsqueeze(-1)
        return x if gb.prev_latent is None else x + gb.prev_latent

    def finish(self, gb, out):
        res = {"prev_latentout": out}
        res.update({k: out[:, :, :6].repeat(1, 22, 1)[:, :150] for k in
                    ("pred_upper", "pred_lower", "pred_hands", "pred_facepose", "pred_exps", "pred_transl")})
        return res


def _longform_worker(rank, world, port, n_frames, ref_path):
    from rag_gesture_b200 import longform as LF
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = LF.LongformSynthesizer(_FakeArch()).run_sharded(n_frames, lambda c, f0, f1: {"cidx": c}, {}, chain_rank=1)
        if rank == 1:
            ref = torch.load(ref_path)
            assert out["window_starts"] == ref["window_starts"]
            for k in ("latents", "pred_upper", "pred_exps"):
                assert torch.equal(out[k], ref[k]), k
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_longform_sharded_world2_gloo(tmp_path):
    """LongformSynthesizer.run_sharded at world_size 2 (windows 0-2 prepared by rank 0, 3-5 by rank 1, chain on rank
    1) == the single-process run: shard ranges, payload packing, the all-gather and the chain order."""
    from rag_gesture_b200 import longform as LF
    n_frames = 150 + 4 * 135
    ref = LF.LongformSynthesizer(_FakeArch()).run_sharded(n_frames, lambda c, f0, f1: {"cidx": c}, {})
    assert len(ref["window_starts"]) == 6          # [0] + range(135, 690, 135): the reference formula (:263)
    path = str(tmp_path / "ref.pt")
    torch.save(ref, path)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_longform_worker, args=(2, port, n_frames, path), nprocs=2, join=True)


class _CpuTextIndex:
    """CPU stand-in for TextSimilarityIndex (the kernels need a GPU): same score definition
    (mean(diag(Q D^T)) over min(Tq, Td) tokens, rag/utils.py:107-118) and the same stable ranking."""

    def __init__(self, db):
        self.names = list(db.idx_2_text.keys())
        self.row = {n: i for i, n in enumerate(self.names)}
        self.feats = [db.idx_2_text[n][0] for n in self.names]

    def score(self, q, name):
        f = self.feats[self.row[name]]
        m = min(q.shape[0], f.shape[0])
        return float((q[:m] * f[:m]).sum(1).mean())

    def rank(self, query, rows, k):
        sc = [self.score(query, self.names[r]) for r in rows]
        return [rows[j] for j in sorted(range(len(rows)), key=lambda j: -sc[j])[:k]]


def _retrieval_worker(rank, world, port, n_db, n_q):
    from rag_gesture_b200 import config as C
    from rag_gesture_b200 import synthetic as S
    from rag_gesture_b200.retrieval import RetrievalDatabase, ShardedDiscourseRetriever, discourse_retrieval
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        db = RetrievalDatabase(dataset=S.SyntheticGestureDataset(n_db, seed=7), **C.retrieval_cfg()).eval()
        cpu_index = _CpuTextIndex(db)
        qs = S.SyntheticGestureDataset(n_q, seed=8)
        mine = list(range(rank, n_q, world))                    # each rank asks for different clips
        queries = []
        for i in mine:
            spk, disc, prom, _, _ = qs.annotations(i)
            queries.append(dict(discourse=disc, prominence=prom, speaker_id=spk, encoded_text=qs.text_feature(i)))
        sharded = ShardedDiscourseRetriever(db, sim_fn=lambda q, names: [cpu_index.score(q, n) for n in names])
        assert sharded.hi - sharded.lo in (n_db // world, n_db // world + 1)
        got = sharded.retrieve(queries)
        n_points = 0
        for q, (si, bounds, qb) in zip(queries, got):
            ref_si, ref_b, ref_qb = discourse_retrieval(
                text="", discourse=q["discourse"], prominence=q["prominence"], speaker_id=q["speaker_id"],
                db_idx_2_sense=db.idx_2_sense, db_idx_2_discbounds=db.idx_2_discbounds,
                db_idx_2_prominence=db.idx_2_prominence, encoded_text=q["encoded_text"], text_feat_cache=db.idx_2_text,
                index=cpu_index, sense_index=db._sense_index, sense_tables=db._sense_tables)
            assert si == ref_si and bounds == ref_b and qb == ref_qb
            n_points += len(si)
        assert n_points > 0
    finally:
        dist.destroy_process_group()


def test_sharded_rule_scoring_world2_gloo():
    """rule scores + text-similarity tie-breaks over a ROW-SHARDED database (each rank scores
    every rank's query points against its half of the rows, one all-gather of 10 candidates per point, merge under
    (score desc, similarity desc, DB order)) == the unsharded discourse_retrieval: names, order and bounds."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_retrieval_worker, args=(2, port, 600, 12), nprocs=2, join=True)


def test_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, 101), nprocs=2, join=True)
