"""ORACLE (test infrastructure, not product): CPU restatement of the reference's exemplar ranking.

  sort_by_text_similarity  <- rag/utils.py:86-132   (torch.mm per candidate, diagonal mean, stable sort)
  discourse_retrieval      <- rag/discourse_retrieval.py:8-316 (score EVERY db sample per connective)
  knn_topk_f64             <- the flat-embedding kNN of BASELINE.json configs[3], float64 scores
Loops exactly like the reference does (so timing it is timing the reference's algorithm).
Parity pin: tests/golden/retrieval.json, produced by the unmodified reference functions.
"""
import numpy as np
import torch

from rag_gesture_b200.retrieval import _clean, map_conns_to_prominence  # host logic pinned separately


def sort_by_text_similarity(names, query, cache):
    if len(names) == 0:
        return names
    score = {}
    for n in names:
        d = cache[n][0]
        score[n] = torch.diagonal(torch.mm(query, d.T)).mean()
    return sorted(score, key=score.get, reverse=True)


def discourse_retrieval(discourse, prominence, speaker_id, db_sense, db_bounds, db_prom, query_feat, cache):
    out_idx, out_bounds = {}, {}
    if len(discourse) == 0:
        return out_idx, out_bounds, {}
    senses, conns = [d[1] for d in discourse], [d[0] for d in discourse]
    qb = {i: (d[0].lower(), d[1], d[6], d[7]) for i, d in enumerate(discourse)}
    qp = map_conns_to_prominence(conns, prominence)
    for i, cv in qp.items():
        if cv is not None:
            qp[i] = (senses[i], cv[1])
    for qi, (sense, conn) in enumerate(zip(senses, conns)):
        score, rb = {}, {}
        for name, entry in db_sense.items():
            score[name] = 0
            spk, disco = entry[0], entry[1:]
            if len(disco) == 0:
                continue
            ss, cc = [d[0] for d in disco], [d[1] for d in disco]
            sp = {j: (None if v is None else (ss[j], v[1])) for j, v in db_prom[name].items()}
            if sense in ss:
                score[name] += 2
                rel = [j for j, s in enumerate(ss) if s == sense]
                top, chosen = rel[0], False
                rc = [cc[j] for j in rel]
                if conn in rc:
                    score[name] += 4
                    top, chosen = rel[rc.index(conn)], True
                if spk == speaker_id:
                    score[name] += 3
                tot, cnt, dif = 0, 0, {}
                for j in rel:
                    if sp[j] is None or qp[qi] is None:
                        continue
                    d = abs(sp[j][1] - qp[qi][1])
                    dif[j] = d
                    tot += 4 / (1 + 2 * d)
                    cnt += 1
                if cnt > 0:
                    score[name] += tot / cnt
                    best = sorted(dif, key=dif.get)
                    if top != best[0] and not chosen:
                        top = best[0]
                rb[name] = db_bounds[name][top]
        order = sorted(score, key=score.get, reverse=True)
        tiers = {}
        for name in order:
            tiers.setdefault(score[name], [])
            if score[name] > 0:
                tiers[score[name]].append(name)
        ranked = []
        for s in sorted(tiers.keys(), reverse=True):
            tier = tiers[s]
            if len(tier) > 1:
                tier = sort_by_text_similarity(tier, query_feat, cache)
            ranked += tier
            if len(ranked) >= 10:
                break
        out_idx[qi] = ranked[:10]
        out_bounds[qi] = {n: (rb[n][1], rb[n][0], round(rb[n][4], 3), round(rb[n][5], 3)) for n in ranked[:10]}
    return out_idx, out_bounds, qb


def knn_topk_f64(db, queries, k):
    """Exact top-k by dot product in float64, order (score desc, index asc); also returns the gap
    between the k-th and (k+1)-th score so tests can tell a real mismatch from an fp32 near-tie."""
    s = queries.double() @ db.double().T
    order = np.lexsort((np.arange(s.shape[1])[None, :].repeat(s.shape[0], 0), -s.numpy()), axis=1)
    idx = torch.from_numpy(order[:, :k].copy())
    top = torch.gather(s, 1, torch.from_numpy(order[:, :k + 1].copy()))
    gaps = (top[:, :-1] - top[:, 1:]).min(dim=1).values
    return idx, torch.gather(s, 1, idx), gaps
