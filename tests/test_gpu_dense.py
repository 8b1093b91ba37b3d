"""K1 parity: tcgen05 dense top-k vs the fp64 oracle — ids bit-exact, scores within 1e-3 relative
(north_star's tolerance; measured error is ~1e-15 because survivors are re-scored in fp64)."""
import os

import numpy as np
import pytest
import torch

from oracle import dense as od
from triple_hybrid_rag_b200 import synth
from triple_hybrid_rag_b200.engine import Engine

pytestmark = pytest.mark.gpu
REL_TOL = 1e-3


def _check(engine, X, Q, k, margin=28, id_base=0):
    engine.dense_index_set(X.to(engine.device), id_base=id_base)
    ids, sc, cnt, gap = engine.dense_topk(Q.to(engine.device), k, margin)
    engine.sync()
    want_i, want_s = od.dense_topk(Q.float().numpy(), X.float().numpy(), k, id_base=id_base)
    ids, sc, cnt, gap = ids.cpu().numpy(), sc.cpu().numpy(), cnt.cpu().numpy(), gap.cpu().numpy()
    n = min(k, X.shape[0])
    assert (cnt == n).all()
    assert np.array_equal(ids, want_i), f"{(ids != want_i).sum()} of {ids.size} ids differ"
    assert np.allclose(sc[:, :n], want_s[:, :n], rtol=REL_TOL, atol=1e-6)
    return gap


@pytest.mark.parametrize("N,D,B,k", [(10_000, 1536, 256, 50), (5_000, 128, 7, 10), (777, 64, 1, 100),
                                     (70_000, 1024, 300, 100)])
def test_dense_topk_matches_oracle(engine, N, D, B, k):
    X = synth.dense_block(0, N, D)
    Q = synth.dense_queries(B, D, X)
    gap = _check(engine, X, Q, k)
    # certificate: the k-th exact score clears everything that was not re-scored by far more than
    # the fp32 accumulation error of a unit-vector dot product
    assert (gap > 2e-6).all() or N <= k + 28


def test_dense_small_corpus_and_id_base(engine):
    X = synth.dense_block(1, 40, 64)
    Q = synth.dense_queries(3, 64, X)
    _check(engine, X, Q, 50, id_base=1_000_000)


def test_dense_duplicate_rows_tie_by_id(engine):
    X = synth.dense_block(2, 3000, 128)
    X[100:400] = X[7]          # 301 identical chunks: exact score ties, must come out in id order
    Q = synth.dense_queries(9, 128, X)
    Q[0] = X[7]
    _check(engine, X, Q, 100)


def test_dense_seeded_path_large_corpus(engine):
    """N >= ~0.9M switches on the seed pass (a prefix of the corpus is scored first and its K'-th best
    score, found by a histogram select, becomes every cluster's initial threshold).  Exact duplicates of a
    top chunk placed far outside the seed prefix must still come out, in id order."""
    N, D, B, k = 1_000_000, 64, 12, 100
    X = synth.dense_rows(0, N, D)
    X[500_000:500_150] = X[7]
    X[999_990:] = X[7]
    Q = synth.dense_queries(B, D, X, n_plant=50_000)
    Q[0] = X[7]
    Q[1] = X[123_456]
    gap = _check(engine, X, Q, k)
    assert (gap > 0).all()


@pytest.mark.parametrize("case", ["ties", "skewed_subsample"])
def test_dense_seed_select_fallbacks(engine, case):
    """The seed select reads the seed lists once, guided by a cut from a strided subsample (the first entries of
    every list), and falls back to a two-level histogram when the cut is useless.  Both fallbacks, on corpora
    large enough for the seed pass: "ties" — four distinct vectors, so tens of thousands of sample scores tie at the
    cut (more than the kernel keeps); "skewed_subsample" — the rows the subsample sees (the head of every 512-row
    block of the prefix) all score near 1 for query 0 while the rest of the sample does not, so fewer than K' scores
    clear the cut.  The answer must not depend on any of this."""
    N, D, B, k = 700_000, 64, 5, 100
    if case == "ties":
        base = synth.dense_block(5, 4, D)
        X = base[torch.arange(N) % 4].contiguous()
        Q = synth.dense_queries(B, D, X)
        Q[0] = base[1]
    else:
        X = synth.dense_rows(0, N, D)
        v = X[N - 1].clone()
        pos = (torch.arange(0, 80_000, 512)[:, None] + torch.arange(55)[None, :]).reshape(-1)
        scale = 0.9 + 0.1 * torch.arange(pos.numel(), dtype=torch.float32) / pos.numel()
        X[pos] = (v.float()[None, :] * scale[:, None]).to(torch.bfloat16)
        Q = synth.dense_queries(B, D, X)
        Q[0] = v
    gap = _check(engine, X, Q, k)
    assert (gap >= 0).all()


def test_dense_single_cta_path_agrees():
    """THR_DENSE_CTA_GROUP=1 (M=128 single-CTA MMA) must give the same answer as the CTA-pair path."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    os.environ["THR_DENSE_CTA_GROUP"] = "1"
    try:
        eng = Engine(0)
        X = synth.dense_block(0, 20_000, 256)
        Q = synth.dense_queries(130, 256, X)
        _check(eng, X, Q, 100)
        eng.close()
    finally:
        del os.environ["THR_DENSE_CTA_GROUP"]


def test_dense_errors(engine):
    from triple_hybrid_rag_b200._lib import ThrError
    with pytest.raises(ThrError):
        engine.dense_index_set(torch.zeros((10, 100), dtype=torch.bfloat16, device=engine.device))  # D % 64
    X = synth.dense_block(0, 1000, 64).to(engine.device)
    engine.dense_index_set(X)
    with pytest.raises(ThrError):
        engine.dense_topk(torch.zeros((2, 64), dtype=torch.bfloat16, device=engine.device), 250, 28)  # k+margin
