"""GPU parity of the retrieval kernels (K10/K11) and of the whole guided batch
(MotionDiffusion.forward) against the reference's golden outputs.  Needs a GPU."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle import retrieval as ORT
from rag_gesture_b200 import config as C
from rag_gesture_b200 import synthetic as S

pytestmark = pytest.mark.gpu
N_DB, N_QUERY = 1200, 48


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def db(dev):
    from rag_gesture_b200.retrieval import RetrievalDatabase
    return RetrievalDatabase(dataset=S.SyntheticGestureDataset(N_DB, seed=7), device=dev, **C.retrieval_cfg()).eval()


def test_text_similarity_kernel(db, dev):
    """rg_text_similarity == mean(diag(Q D^T)) over min(Tq,Td) tokens (rag/utils.py:107-118)."""
    index = db.text_index(dev)
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    for qi in (0, 3):
        q = qs.text_feature(qi)
        got = index.scores(q.to(dev)).cpu()
        ref = torch.stack([torch.diagonal(torch.mm(q.double(), db.idx_2_text[n][0].double().T)).mean()
                           for n in index.names]).float()
        assert torch.allclose(got, ref, rtol=2e-5, atol=2e-5)
        rows = [5, 900, 17, 17, 1199]
        assert torch.equal(index.scores(q.to(dev), rows).cpu(), got[rows])
    names = index.names[:200]
    with open(os.path.join(GOLDEN, "retrieval.json")) as f:
        gold = json.load(f)
    order = [index.names[r] for r in index.rank(qs.text_feature(0).to(dev), [index.row[n] for n in names], 32)]
    assert order == gold["sim_order"][:32]


def test_discourse_retrieval_vs_reference(db, dev):
    with open(os.path.join(GOLDEN, "retrieval.json")) as f:
        gold = json.load(f)
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    for i in range(N_QUERY):
        spk, disc, prom, _, _ = qs.annotations(i)
        idx, bounds, qb = db.retrieval_method["discourse"](
            text="", discourse=disc, prominence=prom, speaker_id=spk, db_idx_2_sense=db.idx_2_sense,
            db_idx_2_discbounds=db.idx_2_discbounds, db_idx_2_prominence=db.idx_2_prominence,
            encoded_text=qs.text_feature(i).to(dev), text_feat_cache=db.idx_2_text)
        g = gold["queries"][i]
        assert json.loads(json.dumps({str(k): v for k, v in idx.items()})) == g["idx"], i
        assert json.loads(json.dumps({str(k): v for k, v in bounds.items()})) == g["bounds"], i


def test_gesture_type_retrieval_vs_reference(db, dev):
    """retrieval_method="gesture_type" with the CUDA tie-break ranking: per-clip triples and the batched forward
    (retrieve_many -> gesture_type_retrieval_many, window placement) against the reference's golden vectors."""
    from rag_gesture_b200.codec import SyntheticGestureCodec
    with open(os.path.join(GOLDEN, "gesture_type.json")) as f:
        gold = json.load(f)
    norm = lambda o: json.loads(json.dumps(o))
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    for i in range(N_QUERY):
        spk, _, _, gest, _ = qs.annotations(i)
        idx, bounds, qb = db.retrieval_method["gesture_type"](
            text="", gesture_labels=gest, speaker_id=spk, db_idx_2_gesture_labels=db.idx_2_gesture_labels,
            encoded_text=qs.text_feature(i).to(dev), text_feat_cache=db.idx_2_text)
        g = gold["queries"][i]
        assert norm({str(k): v for k, v in idx.items()}) == g["idx"], i
        assert norm({str(k): v for k, v in bounds.items()}) == g["bounds"], i
    batch = S.collate([qs[i] for i in range(N_QUERY)])
    cond = dict(text=batch["raw_word"], audio=batch["raw_audio"], text_enc=batch["word"].to(dev),
                text_features=[t.to(dev) for t in batch["text_features"]], audio_enc=batch["audio"].to(dev),
                discourse=batch["discourse"], prominence=batch["prominence"], speaker_ids=batch["speaker_ids"].to(dev),
                gesture_labels=batch["gesture_labels"], text_times=batch["text_segments"])
    for d in (db.test_indexes, db.test_dbounds, db.test_qbounds):
        d.clear()
    torch.manual_seed(5)
    re = db(cond, batch["motion_length"], dev, idx=batch["sample_name"], retrieval_method="gesture_type",
            gesture_rep_encoder=SyntheticGestureCodec(C.denoiser_cfg()["vae_cfg"]).to(dev))
    for d in (db.test_indexes, db.test_dbounds, db.test_qbounds):
        d.clear()
    assert norm([{str(k): list(v) for k, v in d.items()} for d in re["retr_startends"]]) == gold["retr_startends"]
    assert norm([{str(k): list(v) for k, v in d.items()} for d in re["query_startends"]]) == gold["query_startends"]
    assert norm(re["raw_sample_names"]) == gold["raw_sample_names"]
    assert re["re_mask"].sum(1).tolist() == gold["re_mask_sum"]
    ref = torch.from_numpy(np.load(os.path.join(GOLDEN, "gesture_type_latents.npz"))["raw_motion_latents"])
    assert torch.allclose(re["raw_motion_latents"][:4].cpu(), ref, atol=1e-5)


def test_sharded_retriever_vs_reference(db, dev):
    """ShardedDiscourseRetriever on the device (its own block of the text features through rg_text_similarity, the
    merge under (score, similarity, DB order)) against the reference's golden lists and bounds for all 48 queries."""
    from rag_gesture_b200.retrieval import ShardedDiscourseRetriever
    with open(os.path.join(GOLDEN, "retrieval.json")) as f:
        gold = json.load(f)
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    queries = []
    for i in range(N_QUERY):
        spk, disc, prom, _, _ = qs.annotations(i)
        queries.append(dict(discourse=disc, prominence=prom, speaker_id=spk, encoded_text=qs.text_feature(i).to(dev)))
    got = ShardedDiscourseRetriever(db).retrieve(queries)
    for i, (idx, bounds, qb) in enumerate(got):
        g = gold["queries"][i]
        assert json.loads(json.dumps({str(k): v for k, v in idx.items()})) == g["idx"], i
        assert json.loads(json.dumps({str(k): v for k, v in bounds.items()})) == g["bounds"], i


@pytest.mark.parametrize("N,Q,k", [(20000, 1, 8), (20000, 5, 8), (4099, 64, 8), (777, 9, 32), (40, 3, 8), (5, 2, 8)])
def test_knn_topk_exact(dev, N, Q, k):
    """Retrieved indices bit-exact against the float64 oracle unless the oracle itself reports an
    fp32-level near-tie at the cut (north_star: exact indices, deterministic tie-break)."""
    from rag_gesture_b200.parallel import knn_topk
    g = torch.Generator().manual_seed(N + Q)
    db = torch.nn.functional.normalize(torch.randn(N, 768, generator=g), dim=1)
    qs = torch.nn.functional.normalize(torch.randn(Q, 768, generator=g), dim=1)
    if N == 777:                       # exact duplicates: ties must resolve to the lower index
        db[500:520] = db[100:120]
    idx, sc = knn_topk(db.to(dev), qs.to(dev), k, idx_base=1000)
    idx, sc = idx.cpu(), sc.cpu()
    kk = min(k, N)
    ref_idx, ref_sc, gaps = ORT.knn_topk_f64(db, qs, kk)
    for q in range(Q):
        if gaps[q] > 1e-6:
            assert torch.equal(idx[q, :kk] - 1000, ref_idx[q]), q
        assert torch.allclose(sc[q, :kk].double(), ref_sc[q], atol=2e-6)
    if k > N:
        assert bool((idx[:, N:] == -1).all())


def test_knn_topk_generic_dim(dev):
    """dim != 768 takes the generic scan kernel."""
    from rag_gesture_b200.parallel import knn_topk
    g = torch.Generator().manual_seed(5)
    db, qs = torch.randn(3001, 256, generator=g), torch.randn(11, 256, generator=g)
    idx, sc = knn_topk(db.to(dev), qs.to(dev), 8)
    ref_idx, ref_sc, gaps = ORT.knn_topk_f64(db, qs, 8)
    for q in range(11):
        if gaps[q] > 1e-5:
            assert torch.equal(idx[q].cpu(), ref_idx[q]), q
    assert torch.allclose(sc.cpu().double(), ref_sc, atol=1e-4)


def test_knn_merge_matches_single_shard(dev):
    from rag_gesture_b200.parallel import _cuda_merge, knn_topk, shard_range
    g = torch.Generator().manual_seed(3)
    db = torch.randn(10007, 768, generator=g).to(dev)
    qs = torch.randn(16, 768, generator=g).to(dev)
    full_i, full_s = knn_topk(db, qs, 8)
    parts_i, parts_s = [], []
    for r in range(8):
        lo, hi = shard_range(db.shape[0], r, 8)
        i, s = knn_topk(db[lo:hi].contiguous(), qs, 8, idx_base=lo)
        parts_i.append(i)
        parts_s.append(s)
    mi, ms = _cuda_merge(torch.stack(parts_i), torch.stack(parts_s), 8)
    assert torch.equal(mi, full_i) and torch.equal(ms, full_s)


@pytest.mark.parametrize("N,Q,k,dim", [(20000, 300, 8, 768), (4099, 64, 8, 768), (100000, 1000, 8, 768),
                                         (777, 130, 16, 768), (40, 9, 8, 768), (5, 12, 8, 768),
                                         (3001, 257, 8, 256), (9000, 33, 32, 768)])
def test_knn_topk_tc_bit_identical(dev, N, Q, k, dim):
    """Tensor-core path (bf16 similarity GEMM + over-selection + certified exact re-score) == the exact
    fp32 scan, indices AND scores bit for bit, including duplicate rows (tie -> lower index), shards
    smaller than k, ragged last tiles (N % 256, Q % 256 != 0) and a generic dim."""
    from rag_gesture_b200.parallel import KnnIndex, knn_topk
    g = torch.Generator().manual_seed(N + Q)
    db = torch.nn.functional.normalize(torch.randn(N, dim, generator=g), dim=1)
    qs = torch.nn.functional.normalize(torch.randn(Q, dim, generator=g), dim=1)
    if N == 777:
        db[500:520] = db[100:120]
    db, qs = db.to(dev), qs.to(dev)
    index = KnnIndex(db)
    idx, sc = knn_topk(db, qs, k, idx_base=1000, index=index)
    ref_idx, ref_sc = knn_topk(db, qs, k, idx_base=1000)
    assert torch.equal(idx, ref_idx)
    assert torch.equal(sc, ref_sc)
    if k <= 16 and N != 777:
        assert index.last_uncertified == 0            # random unit-norm data: every query is certified
    index.close()


def test_knn_topk_tc_unnormalised_and_fallback(dev):
    """Un-normalised features with a wide norm spread (the reference ranks raw BERT sums, rag/utils.py:104-118)
    and a clustered database where 40 near-duplicates of each query sit in one chunk: the certificate must
    refuse (k-th exact score is not separated from the bound) and the exact fallback must give the answer."""
    from rag_gesture_b200.parallel import KnnIndex, knn_topk
    g = torch.Generator().manual_seed(11)
    db = torch.randn(30000, 768, generator=g) * (0.2 + 3 * torch.rand(30000, 1, generator=g))
    qs = torch.randn(40, 768, generator=g) * 2.5
    # cluster: rows 256..295 are tiny perturbations of query 0 -> 40 candidates of one chunk above everything
    db[256:296] = qs[0] + 1e-3 * torch.randn(40, 768, generator=g)
    db, qs = db.to(dev), qs.to(dev)
    index = KnnIndex(db)
    idx, sc = knn_topk(db, qs, 32, index=index)
    ref_idx, ref_sc = knn_topk(db, qs, 32)
    assert torch.equal(idx, ref_idx) and torch.equal(sc, ref_sc)
    assert index.last_uncertified >= 1
    index.close()


def test_knn_full_size_configs3(dev):
    """BASELINE configs[3] at full size on one GPU: 1M x 768 unit-norm embeddings, Q = 4096, k = 8.  All but a
    handful of queries are certified (a query whose 8th best score is unusually low can miss the margin against a
    chunk's 16th best + eps and is then answered by the exact scan -- still exact, just slower); a spread of 96
    queries is checked bit for bit against the exact fp32 scan of the same shard;
    size-independent properties hold for all 4096: scores sorted (score desc, index asc), indices in range and
    distinct, and each reported score equals the exact dot product of its row recomputed in float64."""
    from rag_gesture_b200.parallel import KnnIndex, knn_topk
    N, Q, k = 1_000_000, 4096, 8
    g = torch.Generator(device=dev).manual_seed(42)
    db = torch.nn.functional.normalize(torch.randn(N, 768, device=dev, generator=g), dim=1)
    qs = torch.nn.functional.normalize(torch.randn(Q, 768, device=dev, generator=g), dim=1)
    index = KnnIndex(db)
    idx, sc = knn_topk(db, qs, k, index=index)
    assert index.last_uncertified <= 8
    sample = torch.arange(0, Q, 43, device=dev)[:96]
    ref_i, ref_s = knn_topk(db, qs[sample].contiguous(), k)
    assert torch.equal(idx[sample], ref_i) and torch.equal(sc[sample], ref_s)
    assert bool((idx >= 0).all()) and bool((idx < N).all())
    assert bool((sc[:, :-1] >= sc[:, 1:]).all())
    ties = sc[:, :-1] == sc[:, 1:]
    assert bool((idx[:, :-1][ties] < idx[:, 1:][ties]).all())
    assert bool((idx.sort(dim=1).values[:, 1:] != idx.sort(dim=1).values[:, :-1]).all())
    exact = (db[idx.reshape(-1)].double() * qs.repeat_interleave(k, 0).double()).sum(1).reshape(Q, k)
    assert float((sc.double() - exact).abs().max()) < 2e-6
    index.close()


def test_knn_tc_sharded_merge(dev):
    """8 tensor-core shard indexes + the merge kernel == the unsharded exact scan."""
    from rag_gesture_b200.parallel import KnnIndex, _cuda_merge, knn_topk, shard_range
    g = torch.Generator().manual_seed(4)
    db = torch.nn.functional.normalize(torch.randn(40003, 768, generator=g), dim=1).to(dev)
    qs = torch.nn.functional.normalize(torch.randn(200, 768, generator=g), dim=1).to(dev)
    full_i, full_s = knn_topk(db, qs, 8)
    parts_i, parts_s = [], []
    for r in range(8):
        lo, hi = shard_range(db.shape[0], r, 8)
        shard = db[lo:hi].contiguous()
        i, s = knn_topk(shard, qs, 8, idx_base=lo, index=KnnIndex(shard))
        parts_i.append(i)
        parts_s.append(s)
    mi, ms = _cuda_merge(torch.stack(parts_i), torch.stack(parts_s), 8)
    assert torch.equal(mi, full_i) and torch.equal(ms, full_s)


def test_knn_packed_merge_equals_unsharded(dev):
    """The packed exchange layout of sharded_knn (per rank [idx int64 Q*k | score fp32 Q*k], local top-k written in
    place, rg_knn_merge_packed reading the gathered blocks) == the unsharded exact scan; odd Q*k (stride padding)."""
    from rag_gesture_b200.parallel import (KnnIndex, _cuda_local_topk, _cuda_merge_packed, knn_topk, packed_stride,
                                           packed_views, shard_range)
    g = torch.Generator().manual_seed(14)
    db = torch.nn.functional.normalize(torch.randn(30011, 768, generator=g), dim=1).to(dev)
    for Q, k, world in ((5, 3, 8), (200, 8, 4), (1, 8, 2)):
        qs = torch.nn.functional.normalize(torch.randn(Q, 768, generator=g), dim=1).to(dev)
        full_i, full_s = knn_topk(db, qs, k)
        stride = packed_stride(Q, k)
        assert stride % 16 == 0 and stride >= 12 * Q * k
        recv = torch.zeros(world * stride, dtype=torch.uint8, device=dev)
        for r in range(world):
            lo, hi = shard_range(db.shape[0], r, world)
            shard = db[lo:hi].contiguous()
            out = packed_views(recv[r * stride:(r + 1) * stride], Q, k)
            if Q > 8:
                index = KnnIndex(shard)
                index.topk(qs, k, lo, out=out)
                index.close()
            else:
                _cuda_local_topk(shard, qs, k, lo, out=out)
        mi, ms = _cuda_merge_packed(recv, world, Q, k)
        assert torch.equal(mi, full_i) and torch.equal(ms, full_s), (Q, k, world)
    with pytest.raises(ValueError):
        knn_topk(db.double(), qs.double(), 8)             # raw fp32 kernels refuse other dtypes


@pytest.fixture(scope="module")
def arch(dev):
    import rag_gesture_b200 as R
    cfg = C.model_cfg()
    cfg["use_retrieval_for_test"] = True
    m = R.build_architecture(cfg, database=S.SyntheticGestureDataset(N_DB, seed=7))
    missing, unexpected = m.model.load_state_dict(S.synthetic_state_dict(0), strict=False)
    assert not unexpected
    return m.to(dev).eval()


def _cpu_noise(shape, device):
    return torch.randn(*shape).to(device)     # global CPU generator, like the reference run


def _full_guided_batch(arch, dev, tol, tier):
    g = np.load(os.path.join(GOLDEN, "pipeline_b3.npz"))
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    batch = S.collate([qs[i] for i in [1, 2, 4]])
    batch["retrieval_method"] = "discourse"
    batch["inference_kwargs"] = dict(use_inversion=True, outpaint=False, inversion_start_time=-1,
                                     insertion_guidance=True, guidance_iters=[0] * 25 + list(range(25)),
                                     guidance_lr=0.1)
    db = arch.model.database
    for d in (db.test_indexes, db.test_dbounds, db.test_qbounds):
        d.clear()
    arch.diffusion_test.noise_fn = _cpu_noise
    torch.manual_seed(2024)
    res = arch(**batch)
    n_ex = sum(len(x) for x in res["retrieval_dict"]["retr_startends"])
    assert n_ex == int(g["n_exemplars"]) and n_ex > 0
    errs = (rel_l2(res["prev_latentout"].cpu(), torch.from_numpy(g["prev_latentout"])),
            rel_l2(res["pred_upper"][:, ::10].cpu(), torch.from_numpy(g["pred_upper"])),
            rel_l2(res["pred_hands"][:, ::10].cpu(), torch.from_numpy(g["pred_hands"])))
    print("full guided batch (%s tier) rel-L2 vs reference: latents %.3g, upper %.3g, hands %.3g" % ((tier,) + errs))
    assert max(errs) < tol
    assert tuple(res["pred_lower"].shape) == (3, 150, 27) and tuple(res["pred_exps"].shape) == (3, 150, 100)


def _prev_latent_chain(arch, dev, tol, tier):
    g = np.load(os.path.join(GOLDEN, "pipeline_b3.npz"))
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    db = arch.model.database
    for d in (db.test_indexes, db.test_dbounds, db.test_qbounds):
        d.clear()
    arch.diffusion_test.noise_fn = _cpu_noise
    prev = None
    torch.manual_seed(77)
    outs = []
    for w in range(2):
        bw = S.collate([qs[5 + w]])
        bw["retrieval_method"] = "discourse"
        bw["inference_kwargs"] = dict(use_inversion=True, insertion_guidance=True,
                                      guidance_iters=[0] * 25 + list(range(25)), guidance_lr=0.1,
                                      use_prev_latent=True, prev_latent=prev)
        prev = arch(**bw)["prev_latentout"]
        outs.append(prev.cpu())
    errs = (rel_l2(outs[0], torch.from_numpy(g["chain_w0"])), rel_l2(outs[1], torch.from_numpy(g["chain_w1"])))
    print("prev-latent chain (%s tier) rel-L2 vs reference: window 0 %.3g, window 1 %.3g" % ((tier,) + errs))
    assert max(errs) < tol


def _gesture_type_batch(arch, dev, tol, tier):
    g = np.load(os.path.join(GOLDEN, "pipeline_gesture_type_b2.npz"))
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    batch = S.collate([qs[i] for i in [0, 2]])
    batch["retrieval_method"] = "gesture_type"
    batch["inference_kwargs"] = dict(use_inversion=True, outpaint=False, inversion_start_time=-1,
                                     insertion_guidance=True, guidance_iters=[0] * 25 + list(range(25)),
                                     guidance_lr=0.1)
    db = arch.model.database
    for d in (db.test_indexes, db.test_dbounds, db.test_qbounds):
        d.clear()
    arch.diffusion_test.noise_fn = _cpu_noise
    torch.manual_seed(515)
    res = arch(**batch)
    n_ex = sum(len(x) for x in res["retrieval_dict"]["retr_startends"])
    assert n_ex == int(g["n_exemplars"]) == 3
    errs = (rel_l2(res["prev_latentout"].cpu(), torch.from_numpy(g["prev_latentout"])),
            rel_l2(res["pred_upper"][:, ::10].cpu(), torch.from_numpy(g["pred_upper"])),
            rel_l2(res["pred_hands"][:, ::10].cpu(), torch.from_numpy(g["pred_hands"])))
    print("gesture_type guided batch (%s tier) rel-L2 vs reference: latents %.3g, upper %.3g, hands %.3g" % ((tier,) + errs))
    assert max(errs) < tol


def test_gesture_type_batch_vs_reference(arch, dev):
    """The whole guided batch with exemplars chosen by the semantic-gesture rules (retrieval_method="gesture_type")."""
    _gesture_type_batch(arch, dev, 1e-3, "fp32")


def test_full_guided_batch_vs_reference(arch, dev):
    """configs[1] in miniature: B=3, discourse retrieval, batched inversion, insertion guidance."""
    _full_guided_batch(arch, dev, 1e-3, "fp32")          # north_star fp32 tier


def test_longform_prev_latent_chain(arch, dev):
    _prev_latent_chain(arch, dev, 1e-3, "fp32")


@pytest.fixture(scope="module", params=[("bf16x3", 1e-3), ("bf16", 2e-2)], ids=["bf16x3", "bf16"])
def arch_tc(request, dev):
    """The same architecture on the tensor-core tiers (tcgen05 GEMMs): bf16x3 is held to the fp32 tier's
    tolerance, bf16 (what bench.py runs) to north_star's 2e-2."""
    import rag_gesture_b200 as R
    from rag_gesture_b200 import _lib as L
    tier, tol = request.param
    cfg = C.model_cfg()
    cfg["use_retrieval_for_test"] = True
    cfg["model"]["precision"] = {"bf16x3": L.PREC_BF16X3, "bf16": L.PREC_BF16}[tier]
    m = R.build_architecture(cfg, database=S.SyntheticGestureDataset(N_DB, seed=7))
    missing, unexpected = m.model.load_state_dict(S.synthetic_state_dict(0), strict=False)
    assert not unexpected
    return m.to(dev).eval(), tol, tier


def test_tc_full_guided_batch_vs_reference(arch_tc, dev):
    """The whole guided batch (retrieval, batched inversion, insertion guidance, decode) on the tensor-core
    tiers against the unmodified reference's golden outputs."""
    m, tol, tier = arch_tc
    _full_guided_batch(m, dev, tol, tier)


def test_tc_longform_prev_latent_chain(arch_tc, dev):
    m, tol, tier = arch_tc
    _prev_latent_chain(m, dev, tol, tier)


def test_tc_gesture_type_batch_vs_reference(arch_tc, dev):
    a, tol, tier = arch_tc
    _gesture_type_batch(a, dev, tol, tier)


def _outpaint_batch(arch, dev, tol, tier):
    g = np.load(os.path.join(GOLDEN, "pipeline_outpaint_b2.npz"))
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    batch = S.collate([qs[i] for i in [9, 14]])
    batch["retrieval_method"] = "discourse"
    batch["inference_kwargs"] = dict(outpaint=True, use_inversion=False, insertion_guidance=False)
    db = arch.model.database
    for d in (db.test_indexes, db.test_dbounds, db.test_qbounds):
        d.clear()
    arch.diffusion_test.noise_fn = _cpu_noise
    torch.manual_seed(4242)
    res = arch(**batch)
    seq = res["retrieval_dict"]["raw_motion_latents"]
    assert int((seq != 0).any(-1).sum()) == int(g["n_rows"]) > 0
    errs = (rel_l2(res["prev_latentout"].cpu(), torch.from_numpy(g["prev_latentout"])),
            rel_l2(res["pred_upper"][:, ::10].cpu(), torch.from_numpy(g["pred_upper"])))
    print("outpaint batch (%s tier) rel-L2 vs reference: latents %.3g, upper %.3g" % ((tier,) + errs))
    assert max(errs) < tol


def test_outpaint_batch_vs_reference(arch, dev):
    """The outpaint branch (diffusion_architecture.py:279-283,472): exemplar latents blended on every step of the
    plain loop, against the unmodified reference's output (tests/golden/make_golden.py outpaint)."""
    _outpaint_batch(arch, dev, 1e-3, "fp32")


def test_tc_outpaint_batch_vs_reference(arch_tc, dev):
    m, tol, tier = arch_tc
    _outpaint_batch(m, dev, tol, tier)


def test_resident_corpus_matches_host_fetch(arch, dev):
    """The HBM-resident exemplar corpus (one gather per field) yields the same prepared batch as the
    per-batch host fetch + stack + H2D path it replaces."""
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    db = arch.model.database
    gbs = []
    for budget in (0, 64 << 30):
        db.corpus_budget_bytes, db._corpus = budget, None
        for d in (db.test_indexes, db.test_dbounds, db.test_qbounds):
            d.clear()
        batch = S.collate([qs[i] for i in [1, 2, 4, 7]])
        batch["retrieval_method"] = "discourse"
        batch["inference_kwargs"] = dict(use_inversion=True, insertion_guidance=True,
                                         guidance_iters=[0] * 25 + list(range(25)), guidance_lr=0.1)
        torch.manual_seed(5)
        gbs.append(arch.prepare(**batch))
        assert (db._corpus is None) == (budget == 0)
    a, b = gbs
    assert a.jobs == b.jobs and len(a.jobs) > 0 and a.windows == b.windows
    for k in a.ex:
        assert torch.equal(a.ex[k], b.ex[k]), k
    ra, rb = a.model_kwargs["re_dict"], b.model_kwargs["re_dict"]
    for k in ("raw_motion_latents", "raw_motion", "raw_trans", "raw_facial", "re_mask"):
        assert torch.equal(ra[k], rb[k]), k


def test_guided_pipeline_equals_sequential_forward(arch, dev):
    """GuidedPipeline (stage 1 of batch i+1 on a worker thread + side stream while batch i's loops run) gives
    bit-identical results to sequential MotionDiffusion.forward calls: same host batches, same seeds."""
    from rag_gesture_b200.architecture import GuidedPipeline
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    arch.diffusion_test.noise_fn = None          # sampler noise from the CUDA generator (main thread)

    def batches():
        for ids in ([1, 2, 4], [7, 8], [10, 11, 12, 13], [20]):
            b = S.collate([qs[i] for i in ids])
            for k, v in b.items():
                if torch.is_tensor(v):
                    b[k] = v.pin_memory()
            b["retrieval_method"] = "discourse"
            b["inference_kwargs"] = dict(use_inversion=True, outpaint=False, inversion_start_time=-1,
                                         insertion_guidance=True, guidance_iters=[0] * 25 + list(range(25)),
                                         guidance_lr=0.1)
            yield b

    def reset():
        db = arch.model.database
        for d in (db.test_indexes, db.test_dbounds, db.test_qbounds):
            d.clear()
        torch.manual_seed(5)
        torch.cuda.manual_seed(6)

    def grab(r):
        return {k: r[k].cpu() for k in ("prev_latentout", "pred_upper", "pred_hands")}

    reset()
    seq = [grab(arch(**b)) for b in batches()]
    reset()
    pipe = [grab(r) for r in GuidedPipeline(arch).run(batches())]
    assert len(pipe) == len(seq) == 4
    for a, b in zip(seq, pipe):
        for k in a:
            assert torch.equal(a[k], b[k]), k
