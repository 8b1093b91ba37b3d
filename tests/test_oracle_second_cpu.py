"""The two independent CPU restatements (scipy.sparse BM25, torch.einsum MaxSim) against the primary oracles on
small seeded inputs: two statements of each definition that share no code must agree (fp64 vs fp32 within 1e-3
relative — north_star's tolerance — and on the ranking wherever the scores are not within rounding of each other)."""
import numpy as np
import torch

from oracle import bm25 as ob
from oracle import bm25_sparse as obs
from oracle import maxsim as om
from oracle import maxsim_einsum as ome
from triple_hybrid_rag_b200 import synth


def test_sparse_bm25_agrees_with_the_posting_list_oracle():
    n, V, k = 6000, 1500, 50
    doc, term, tf, L = (x.numpy() for x in synth.bm25_block_coo(0, n, V=V))
    orc = ob.CsrIndex.from_coo(doc, term, tf, L, V)
    spm = obs.SparseBM25(doc, term, tf, L, V)
    qs = synth.bm25_queries(24, V=V, min_rank=20) + [[3, 3, 7], [V + 5, 2], []]
    wi, ws, wc = ob.bm25_topk(orc, qs, k)
    si, ss, sc = spm.topk(qs, k)
    assert np.array_equal(wc, sc)
    for q in range(len(qs)):
        c = wc[q]
        assert np.allclose(ws[q, :c], ss[q, :c], rtol=1e-3)
        S = spm.scores([qs[q]])[0]
        # same ids, except where the fp64 scores of the swapped docs are within fp32 rounding of each other
        for a, b in zip(wi[q, :c], si[q, :c]):
            assert a == b or abs(S[a] - S[b]) <= 1e-5 * abs(S[b])


def test_einsum_maxsim_agrees_with_the_loop_oracle():
    Qt, Dt, cand = synth.maxsim_tokens(3, 17, Tq=32, Td=64)
    a = om.maxsim(Qt.float().numpy(), Dt.float().numpy(), cand.numpy())
    b = ome.maxsim_einsum(Qt.float(), Dt.float(), cand).numpy()
    assert np.allclose(a, b, rtol=1e-12, atol=1e-12)
