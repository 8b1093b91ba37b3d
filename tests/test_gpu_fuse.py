"""K3/K5 parity: CUDA fusion vs the reference's own outputs (golden) and vs the oracle — bit-exact."""
import itertools
import random

import pytest
import torch

from oracle import fusion as of
from tests.util import csr

pytestmark = pytest.mark.gpu
fh = float.fromhex


def _run(engine, variant, cases, weights, raws=None, tie_mode=0, **kw):
    dev = engine.device
    B = len(cases)
    lists = []
    for c in range(3):
        col = [case[c] for case in cases]
        if all(x is None for x in col):
            lists.append(None)
            continue
        ids, off = csr(col, torch.int64, dev)
        sc = None
        if raws is not None:
            sc, _ = csr([r[c] if case[c] is not None else None for r, case in zip(raws, cases)], torch.float64, dev)
        lists.append((ids, off, sc))
    w = torch.tensor(weights, dtype=torch.float64, device=dev)
    o_ids, o_rrf, o_rk, o_raw, o_cnt = engine.fuse(variant, B, lists, w, want_raw=raws is not None,
                                                  tie_mode=tie_mode, max_out=768, **kw)
    engine.sync()
    o_ids, o_rrf, o_rk, o_cnt = o_ids.cpu(), o_rrf.cpu(), o_rk.cpu(), o_cnt.cpu()
    o_raw = None if o_raw is None else o_raw.cpu()
    out = []
    for b in range(B):
        n = int(o_cnt[b])
        out.append([(int(o_ids[b, i]), float(o_rrf[b, i]).hex(), [int(x) for x in o_rk[b, i]],
                     None if o_raw is None else [float(x).hex() for x in o_raw[b, i]]) for i in range(n)])
    return out


def test_rag2_fuse_golden(engine, fusion_golden):
    cases = fusion_golden["rag2"]
    got = _run(engine, 0, [c["lists"] for c in cases], [c["weights"] for c in cases])
    for g, c in zip(got, cases):
        assert [(i, r, k) for i, r, k, _ in g] == [(o["id"], o["rrf"], o["ranks"]) for o in c["out"]]


def test_lib_fuse_golden(engine, fusion_golden):
    cases = fusion_golden["lib"]
    key = lambda c: (c["thr"], c["alpha"], c["denoise"], c["top_k"] or 0)
    for params, grp in itertools.groupby(sorted(cases, key=key), key=key):
        grp = list(grp)
        raws = [[[fh(x) for x in r] for r in c["raw"]] for c in grp]
        got = _run(engine, 1, [c["lists"] for c in grp], [c["weights"] or [0.7, 0.8, 1.0] for c in grp], raws=raws,
                   safety_thr=params[0], alpha=params[1], denoise=params[2], top_k=params[3])
        for g, c in zip(got, grp):
            assert [(i, r, raw) for i, r, _, raw in g] == [(o["id"], o["rrf"], o["raw"]) for o in c["out"]], params


def test_rag1_fuse_golden(engine, fusion_golden):
    cases = fusion_golden["rag1"]
    for k, grp in itertools.groupby(sorted(cases, key=lambda c: c["rrf_k"]), key=lambda c: c["rrf_k"]):
        grp = list(grp)
        ls = [([l for l in c["lists"] if l is not None] + [None] * 3)[:3] for c in grp]
        got = _run(engine, 2, ls, [[1.0, 1.0, 1.0]] * len(grp), rrf_k=k)
        for g, c in zip(got, grp):
            assert [(i, r) for i, r, _, _ in g] == [(o["id"], o["rrf"]) for o in c["out"]]


def test_fuse_vs_oracle_random_and_ties(engine):
    rng = random.Random(5)
    cases, weights = [], []
    for _ in range(300):
        pool = rng.choice([60, 300, 100000])
        ls = [rng.sample(range(pool), rng.randint(0, min(pool, 100))),
              rng.sample(range(pool), rng.randint(0, min(pool, 100))),
              rng.sample(range(pool), rng.randint(0, min(pool, 50)))]
        cases.append(ls)
        weights.append([0.7, 0.8, 1.0])
    # disjoint lists: L10 / S20 / G40 tie exactly at 0.01 (SURVEY App. A golden B)
    cases.append([list(range(1001, 1011)), list(range(2001, 2021)), list(range(3001, 3041))])
    weights.append([0.7, 0.8, 1.0])
    cases.append([[], [], []])
    weights.append([0.7, 0.8, 1.0])
    for tie in (0, 1):
        got = _run(engine, 0, cases, weights, tie_mode=tie)
        for g, c, w in zip(got, cases, weights):
            want = of.fuse(of.RAG2, c, w, tie_mode=tie)
            assert [(i, r, k) for i, r, k, _ in g] == [(x["id"], x["rrf"].hex(), list(x["ranks"])) for x in want]


def test_fuse_max_lengths(engine):
    ls = [[list(range(0, 256)), list(range(128, 384)), list(range(300, 556))]]
    got = _run(engine, 0, ls, [[0.7, 0.8, 1.0]])
    want = of.fuse(of.RAG2, ls[0])
    assert [(i, r) for i, r, _, _ in got[0]] == [(x["id"], x["rrf"].hex()) for x in want]


def test_safety_golden(engine, fusion_golden):
    dev = engine.device
    cases = fusion_golden["safety"]
    key = lambda c: (c["thr"], c["alpha"], c["top_k"])
    for params, grp in itertools.groupby(sorted(cases, key=key), key=key):
        grp = list(grp)
        rrf, off = csr([[fh(x) for x in c["rrf"]] for c in grp], torch.float64, dev)
        rer, _ = csr([[0.0 if x is None else fh(x) for x in c["rerank"]] for c in grp], torch.float64, dev)
        has, _ = csr([[0 if x is None else 1 for x in c["rerank"]] for c in grp], torch.uint8, dev)
        keep, refused, mx = engine.safety(off, rrf, rer, has, *params)
        engine.sync()
        keep, refused, mx, off = keep.cpu(), refused.cpu(), mx.cpu(), off.cpu()
        for b, c in enumerate(grp):
            lo, hi = int(off[b]), int(off[b + 1])
            kept = [i - lo for i in range(lo, hi) if keep[i]]
            o = c["out"]
            assert kept == o["kept"] and bool(refused[b]) == o["refused"] and float(mx[b]).hex() == o["max"]


def test_merge_topk(engine):
    g = torch.Generator().manual_seed(3)
    G, B, k = 8, 37, 100
    sc = torch.rand((G, B, k), generator=g, dtype=torch.float64)
    sc[:, :, ::7] = 0.5  # exact ties across shards
    sc, _ = torch.sort(sc, dim=2, descending=True)
    ids = torch.randperm(G * B * k, generator=g).view(G, B, k).to(torch.int64)
    cnt = torch.randint(0, k + 1, (G, B), generator=g, dtype=torch.int32)
    cnt[0, 0] = 0
    o_sc, o_ids, o_cnt = engine.merge_topk(sc.to(engine.device), ids.to(engine.device), cnt.to(engine.device), 100)
    engine.sync()
    o_sc, o_ids, o_cnt = o_sc.cpu(), o_ids.cpu(), o_cnt.cpu()
    for b in range(B):
        rows = [(-float(sc[gg, b, j]), int(ids[gg, b, j])) for gg in range(G) for j in range(int(cnt[gg, b]))]
        rows.sort()
        rows = rows[:100]
        assert int(o_cnt[b]) == len(rows)
        assert [(-float(o_sc[b, i]), int(o_ids[b, i])) for i in range(len(rows))] == rows
