"""The pruned BM25 restatement (oracle/bm25_pruned.py — the round-2 algorithm stated on the CPU) returns exactly what
the exhaustive oracle returns: ids, fp32 score bits and counts, on corpora with heavy terms, many exact ties and
queries whose terms are all frequent, for several k and block sizes; and it does prune."""
import numpy as np
import pytest

from oracle import bm25 as ob
from oracle.bm25_pruned import bm25_topk_pruned
from triple_hybrid_rag_b200 import synth


def _index(n_docs, V, seed_block=0):
    doc, term, tf, L = synth.bm25_block_coo(seed_block, n_docs, V=V)
    return ob.CsrIndex.from_coo(doc.numpy(), term.numpy(), tf.numpy(), L.numpy(), V)


@pytest.mark.parametrize("n_docs,V,k,block", [(30_000, 3_000, 100, 4096), (30_000, 3_000, 7, 32768),
                                              (12_000, 300, 50, 1024), (5_000, 40, 256, 2048)])
def test_pruned_equals_exhaustive(n_docs, V, k, block):
    idx = _index(n_docs, V)
    qs = synth.bm25_queries(24, V=V, min_rank=min(30, V // 4)) + [[0, 1, 2, 3, 4, 5], [V - 1], [], [1, 1, 2], [V + 5, -1]]
    wi, ws, wc = ob.bm25_topk(idx, qs, k, id_base=77)
    stats = {}
    gi, gs, gc = bm25_topk_pruned(idx, qs, k, id_base=77, block=block, stats=stats)
    assert np.array_equal(gc, wc)
    assert np.array_equal(gi, wi)
    assert np.array_equal(gs.view(np.uint32), ws.view(np.uint32))
    assert stats["streamed"] <= stats["postings"]


def test_pruning_actually_prunes_on_the_benchmark_distribution():
    idx = _index(120_000, 100_000)
    qs = synth.bm25_queries(32, V=100_000)
    stats = {}
    gi, gs, gc = bm25_topk_pruned(idx, qs, 100, stats=stats)
    wi, ws, wc = ob.bm25_topk(idx, qs, 100)
    assert np.array_equal(gi, wi) and np.array_equal(gs.view(np.uint32), ws.view(np.uint32))
    assert stats["streamed"] < 0.8 * stats["postings"]      # even at 120k docs a good part of the lists is skipped
