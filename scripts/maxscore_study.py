"""How much would MaxScore-style pruning save on the benchmark's BM25 workload?  CPU study on the oracle
(numpy; no GPU): python scripts/maxscore_study.py [docs] [queries]
For each query: tau = exact k-th best score; terms sorted by upper bound ub_t = idf_t * max impact_t; the longest
prefix with sum(ub) <= tau is NON-ESSENTIAL: a doc that contains only those terms cannot reach the top-k, so their
postings need not be streamed — only looked up for the docs the essential terms touch."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import bm25 as ob
from triple_hybrid_rag_b200 import synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
V, k = 100_000, 100
doc, term, tf, L = synth.bm25_block_coo(0, N, V=V)
idx = ob.CsrIndex.from_coo(doc.numpy(), term.numpy(), tf.numpy(), L.numpy(), V)
qs = synth.bm25_queries(B, V=V)
ids, sc, cnt = ob.bm25_topk(idx, qs, k)
tot_post = ess_post = touched_ess = 0
fr = []
for q, terms in enumerate(qs):
    tau = float(sc[q, cnt[q] - 1]) if cnt[q] == k else 0.0
    ub, df = [], []
    for t in terms:
        lo, hi = idx.indptr[t], idx.indptr[t + 1]
        df.append(int(hi - lo))
        ub.append(float(idx.idf[t]) * float(idx.imp[lo:hi].max()) if hi > lo else 0.0)
    order = np.argsort(ub)
    acc, non_ess = 0.0, set()
    for j in order:
        if acc + ub[j] <= tau * (1 - 1e-6):
            acc += ub[j]; non_ess.add(j)
        else:
            break
    p_all = sum(df)
    p_ess = sum(d for j, d in enumerate(df) if j not in non_ess)
    tot_post += p_all; ess_post += p_ess
    fr.append(p_ess / max(p_all, 1))
    docs_ess = set()
    for j, t in enumerate(terms):
        if j not in non_ess:
            docs_ess.update(idx.doc[idx.indptr[t]:idx.indptr[t + 1]].tolist())
    touched_ess += len(docs_ess) * len(non_ess)
print(f"docs={N} queries={B} k={k}: postings of all terms {tot_post}, of essential terms {ess_post} "
      f"({100 * ess_post / tot_post:.1f} %), look-ups into non-essential lists {touched_ess} "
      f"({100 * touched_ess / tot_post:.1f} % of the postings)")
print("per-query essential fraction: p10 %.2f  p50 %.2f  p90 %.2f" % tuple(np.percentile(fr, [10, 50, 90])))

# the same question with a RUNNING tau, as a kernel would have it (oracle/bm25_pruned.py, exactness pinned by
# tests/test_oracle_bm25_pruned.py): blocks of 32768 docs in ascending id, tau = k-th best so far
from oracle.bm25_pruned import bm25_topk_pruned
for budget in (1.0, 0.7, 0.5, 0.3):
    st = {}
    gi, gs, gc = bm25_topk_pruned(idx, qs, k, stats=st, ne_budget=budget)
    assert np.array_equal(gi, ids) and np.array_equal(gs.view(np.uint32), sc.view(np.uint32))
    print(f"running tau, non-essential budget {budget:.1f} * tau: streamed {100 * st['streamed'] / st['postings']:.1f} % of the "
          f"postings, survivors re-scored {100 * st['survivors'] / max(st['touched'], 1):.1f} % of the touched docs, "
          f"look-ups {100 * st['lookups'] / st['postings']:.1f} % of the postings; results identical")
