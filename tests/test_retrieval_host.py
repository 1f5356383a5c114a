"""CPU tests: the retrieval oracle and the product's host-side retrieval logic (rule scores, tiers,
window placement) against vectors produced by the reference's own functions (retrieval.json)."""
import json
import os

import pytest
import torch

from conftest import GOLDEN
from oracle import retrieval as ORT
from rag_gesture_b200 import config as C
from rag_gesture_b200 import synthetic as S
from rag_gesture_b200.codec import SyntheticGestureCodec
from rag_gesture_b200.retrieval import RetrievalDatabase

N_DB, N_QUERY = 1200, 48


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLDEN, "retrieval.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def db():
    return RetrievalDatabase(dataset=S.SyntheticGestureDataset(N_DB, seed=7), device="cpu", **C.retrieval_cfg()).eval()


class OracleIndex:
    """Ranks with the oracle's similarity on the CPU (tests only; the product index is CUDA)."""

    def __init__(self, db):
        self.names = list(db.idx_2_text.keys())
        self.row = {n: i for i, n in enumerate(self.names)}
        self.cache, self.device = db.idx_2_text, torch.device("cpu")

    def rank(self, query, rows, k):
        order = ORT.sort_by_text_similarity([self.names[r] for r in rows], query, self.cache)
        return [self.row[n] for n in order[:k]]


def _norm(o):
    return json.loads(json.dumps(o))


def test_oracle_retrieval_vs_reference(gold, db):
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    for i in range(N_QUERY):
        spk, disc, prom, _, _ = qs.annotations(i)
        idx, bounds, qb = ORT.discourse_retrieval(disc, prom, spk, db.idx_2_sense, db.idx_2_discbounds,
                                                  db.idx_2_prominence, qs.text_feature(i), db.idx_2_text)
        g = gold["queries"][i]
        assert _norm({str(k): v for k, v in idx.items()}) == g["idx"], i
        assert _norm({str(k): v for k, v in bounds.items()}) == g["bounds"], i
        assert _norm({str(k): list(v) for k, v in qb.items()}) == g["qbounds"], i
    names = list(db.idx_2_text.keys())[:200]
    assert ORT.sort_by_text_similarity(names, qs.text_feature(0), db.idx_2_text) == gold["sim_order"]


def test_db_order_is_lmdb_cursor_order(db):
    names = list(db.idx_2_text.keys())
    assert names == sorted(names, key=lambda s: s.encode("ascii")) and len(names) == N_DB


def test_product_host_logic_vs_reference(gold, db):
    """discourse scoring via the inverted index + tiers + window placement == reference, with the
    similarity ranking delegated to the oracle (the CUDA ranking is tested in test_gpu_retrieval)."""
    db._index = OracleIndex(db)
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    for i in range(N_QUERY):
        spk, disc, prom, gest, _ = qs.annotations(i)
        idx, bounds, qb = db.retrieval_method["discourse"](
            text="", discourse=disc, prominence=prom, speaker_id=spk, db_idx_2_sense=db.idx_2_sense,
            db_idx_2_discbounds=db.idx_2_discbounds, db_idx_2_prominence=db.idx_2_prominence,
            encoded_text=qs.text_feature(i), text_feat_cache=db.idx_2_text)
        g = gold["queries"][i]
        assert _norm({str(k): v for k, v in idx.items()}) == g["idx"], i
        assert _norm({str(k): v for k, v in bounds.items()}) == g["bounds"], i
    batch = S.collate([qs[i] for i in range(N_QUERY)])
    cond = dict(text=batch["raw_word"], audio=batch["raw_audio"], text_enc=batch["word"],
                text_features=batch["text_features"], audio_enc=batch["audio"], discourse=batch["discourse"],
                prominence=batch["prominence"], speaker_ids=batch["speaker_ids"],
                gesture_labels=batch["gesture_labels"], text_times=batch["text_segments"])
    torch.manual_seed(5)
    re = db(cond, batch["motion_length"], "cpu", idx=batch["sample_name"], retrieval_method="discourse",
            gesture_rep_encoder=SyntheticGestureCodec(C.denoiser_cfg()["vae_cfg"]))
    assert _norm([{str(k): list(v) for k, v in d.items()} for d in re["retr_startends"]]) == gold["retr_startends"]
    assert _norm([{str(k): list(v) for k, v in d.items()} for d in re["query_startends"]]) == gold["query_startends"]
    assert _norm(re["raw_sample_names"]) == gold["raw_sample_names"]
    assert re["re_mask"].sum(1).tolist() == gold["re_mask_sum"]
    import numpy as np
    ref = torch.from_numpy(np.load(os.path.join(GOLDEN, "retrieval_latents.npz"))["raw_motion_latents"])
    assert torch.allclose(re["raw_motion_latents"][:4], ref, atol=1e-6)
    # second call with the same names: served from the retrieval cache (the reference crashes here)
    re2 = db(cond, batch["motion_length"], "cpu", idx=batch["sample_name"], retrieval_method="discourse",
             gesture_rep_encoder=SyntheticGestureCodec(C.denoiser_cfg()["vae_cfg"]))
    assert re2["retr_startends"] == re["retr_startends"]


@pytest.fixture(scope="module")
def gold_gt():
    with open(os.path.join(GOLDEN, "gesture_type.json")) as f:
        return json.load(f)


def test_partial_ratio_restatements():
    """fuzzywuzzy 0.18.0 is not in the image: the product's and the oracle's restatements of fuzz.partial_ratio
    agree with each other and with the answers the package documents."""
    import random
    from oracle import fuzz_ratio
    from rag_gesture_b200.wordsim import partial_ratio, word_similarity
    for fn in (partial_ratio, fuzz_ratio.partial_ratio):
        assert fn("YANKEES", "NEW YORK YANKEES") == 100
        assert fn("NEW YORK METS", "NEW YORK YANKEES") == 69
        assert fn("this is a test", "this is a test!") == 100
        assert fn("", "abc") == 0 and fn(None, "abc") == 0 and fn("abc", "abc") == 100
    r = random.Random(3)
    words = S.GESTURE_WORDS + S.CONNECTIVES + S.FILLERS
    for _ in range(2000):
        a, b = r.choice(words), r.choice(words)
        if r.random() < 0.3:
            a = "".join(r.choice("abcdeh ") for _ in range(r.randint(1, 12)))
        assert partial_ratio(a, b) == fuzz_ratio.partial_ratio(a, b), (a, b)
    # the hook for a real embedding model: the reference's multi-word averaging, and its catch-all fall-back
    const = lambda w1, w2: 0.25 if w1 != "open" else 0.75
    assert word_similarity("open hand", "big", const) == (0.75 + 0.25) / 2
    assert word_similarity("big", "first of all", const) == pytest.approx(0.25)
    assert word_similarity("open hand", "both hands", const) == (0.75 + 0.75 + 0.25 + 0.25) / 4

    def missing(w1, w2):
        raise KeyError(w1)
    assert word_similarity("around", "round", missing) == partial_ratio("around", "round") / 100 == 1.0


def _cond(batch):
    return dict(text=batch["raw_word"], audio=batch["raw_audio"], text_enc=batch["word"],
                text_features=batch["text_features"], audio_enc=batch["audio"], discourse=batch["discourse"],
                prominence=batch["prominence"], speaker_ids=batch["speaker_ids"],
                gesture_labels=batch["gesture_labels"], text_times=batch["text_segments"])


def test_gesture_type_host_logic_vs_reference(gold_gt, db):
    """gesture-type rules through the type index + tiers + window placement (the -0.2/+0.1 s padding of long
    exemplars) == the reference's gesture_type_retrieval and forward(retrieval_method="gesture_type")."""
    import numpy as np
    db._index = OracleIndex(db)
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    n_labels = 0
    for i in range(N_QUERY):
        spk, _, _, gest, _ = qs.annotations(i)
        idx, bounds, qb = db.retrieval_method["gesture_type"](
            text="", gesture_labels=gest, speaker_id=spk, db_idx_2_gesture_labels=db.idx_2_gesture_labels,
            encoded_text=qs.text_feature(i), text_feat_cache=db.idx_2_text)
        g = gold_gt["queries"][i]
        assert _norm({str(k): v for k, v in idx.items()}) == g["idx"], i
        assert _norm({str(k): v for k, v in bounds.items()}) == g["bounds"], i
        assert _norm({str(k): list(v) for k, v in qb.items()}) == g["qbounds"], i
        n_labels += len(idx)
    assert n_labels > 60
    batch = S.collate([qs[i] for i in range(N_QUERY)])
    for d in (db.test_indexes, db.test_dbounds, db.test_qbounds):      # keyed by clip name, ONE method per clip (as the
        d.clear()                                                      # reference): drop the discourse test's entries
    torch.manual_seed(5)
    re = db(_cond(batch), batch["motion_length"], "cpu", idx=batch["sample_name"], retrieval_method="gesture_type",
            gesture_rep_encoder=SyntheticGestureCodec(C.denoiser_cfg()["vae_cfg"]))
    assert _norm([{str(k): list(v) for k, v in d.items()} for d in re["retr_startends"]]) == gold_gt["retr_startends"]
    assert _norm([{str(k): list(v) for k, v in d.items()} for d in re["query_startends"]]) == gold_gt["query_startends"]
    assert _norm(re["raw_sample_names"]) == gold_gt["raw_sample_names"]
    assert _norm([{str(k): list(v) for k, v in d.items()} for d in re["raw_type2words"]]) == gold_gt["raw_type2words"]
    assert re["re_mask"].sum(1).tolist() == gold_gt["re_mask_sum"]
    ref = torch.from_numpy(np.load(os.path.join(GOLDEN, "gesture_type_latents.npz"))["raw_motion_latents"])
    assert torch.allclose(re["raw_motion_latents"][:4], ref, atol=1e-6)
    # a clip without semantic labels retrieves nothing
    assert db.retrieval_method["gesture_type"](
        text="", gesture_labels=[{"name": "beat", "word": "and", "start": 0.0, "end": 1.0}], speaker_id=3,
        db_idx_2_gesture_labels=db.idx_2_gesture_labels, encoded_text=qs.text_feature(0),
        text_feat_cache=db.idx_2_text) == ({}, {}, {})


def test_window_placement_edge_cases(db):
    # exemplar at the very end of the clip, query at the start: clamps + parity rules (App. D)
    w = db.place_window(("and", "s", 0.0, 0.2), ("and", "s", 9.8, 10.0), "discourse", -1)
    assert w == ((9, 10), (0, 1))
    # collision with the previous placement pushes the window right and truncates at 10 chunks
    w = db.place_window(("and", "s", 9.0, 9.9), ("and", "s", 2.0, 5.0), "discourse", 8)
    (r0, r1), (s, e) = w
    assert (s, e) == (8, 10) and r1 - r0 == 2
    assert db.place_window(("and", "s", 9.0, 9.9), ("and", "s", 2.0, 5.0), "discourse", 10) is None
    # exemplar bounds past the clip end clamp to a zero-length window and are skipped (:641-642)
    assert db.place_window(("and", "s", 1.0, 2.0), ("and", "s", 10.7, 10.8), "discourse", -1) is None
    # inverted exemplar bounds: the reference stops in breakpoint() (:652-653); we raise
    with pytest.raises(AssertionError):
        db.place_window(("and", "s", 1.0, 2.0), ("and", "s", 0.0, -0.4), "discourse", -1)


def test_unsupported_methods_raise(db):
    with pytest.raises(NotImplementedError):
        db.retrieval_method["llm"](text="")
    with pytest.raises(AssertionError):
        db.retrieve("prosody", None, None, None, [], [], [], [], 0)


def test_vectorised_scores_equal_loop(db):
    """SenseTable.score (numpy, all samples of a sense at once) == the per-sample loop, bit for bit
    (float64 scores and chosen entry), for every query connective of the synthetic set."""
    from rag_gesture_b200 import retrieval as RT
    qs = S.SyntheticGestureDataset(N_QUERY, seed=8)
    tabs, conn_ids = {}, {}
    n_checked = 0
    for i in range(N_QUERY):
        spk, disc, prom, _, _ = qs.annotations(i)
        if not disc:
            continue
        conns = [d[0] for d in disc]
        q_prom = RT.map_conns_to_prominence(conns, prom)
        for k, cv in q_prom.items():
            if cv is not None:
                q_prom[k] = (disc[k][1], cv[1])
        for qi, d in enumerate(disc):
            sense, conn = d[1], d[0]
            ref_scored, ref_bounds = RT._score_loop(sense, conn, qi, q_prom, spk, db._sense_index, db.idx_2_sense,
                                                    db.idx_2_discbounds, db.idx_2_prominence)
            if sense not in tabs:
                tabs[sense] = RT.SenseTable(sense, db._sense_index.get(sense, []), db.idx_2_sense,
                                            db.idx_2_prominence, conn_ids)
            sc, top = tabs[sense].score(conn_ids.get(conn), spk, None if q_prom[qi] is None else float(q_prom[qi][1]))
            assert [n for n, _ in ref_scored] == tabs[sense].names
            assert [s_ for _, s_ in ref_scored] == sc.tolist()
            assert [ref_bounds[n] for n in tabs[sense].names] == [db.idx_2_discbounds[n][t] for n, t in zip(tabs[sense].names, top.tolist())]
            n_checked += len(ref_scored)
    assert n_checked > 5000
