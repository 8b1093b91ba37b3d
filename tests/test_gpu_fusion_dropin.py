"""Drop-in surface of the two fusion entry points (triple_hybrid_rag_b200/fusion.py) replayed against the
reference's own outputs: the golden cases were produced by RRFFusion.fuse and HybridSearcher._rrf_fusion
(tests/golden/make_golden.py); here the same inputs go in as result OBJECTS, the way a caller of the
reference passes them, and the returned objects must carry the reference's ids, order and fp64 bits."""
from types import SimpleNamespace

import pytest

from triple_hybrid_rag_b200.fusion import GpuRRFFusion, rag1_rrf_fusion

pytestmark = pytest.mark.gpu
fh = float.fromhex
FIELDS = ("lexical_score", "semantic_score", "graph_score")


def _lib_objects(case):
    lists = []
    for c in range(3):
        objs = []
        for i, cid in enumerate(case["lists"][c] or []):
            o = SimpleNamespace(chunk_id=f"00000000-0000-0000-0000-{cid:012d}", lexical_score=0.0, semantic_score=0.0,
                                graph_score=0.0, rrf_score=0.0, final_score=0.0, metadata={}, tag=(c, i))
            setattr(o, FIELDS[c], fh(case["raw"][c][i]))
            objs.append(o)
        lists.append(objs)
    return lists


def test_gpu_rrf_fusion_replays_the_library_golden(engine, fusion_golden):
    cases = fusion_golden["lib"]
    for case in cases:
        cfg = SimpleNamespace(rag_lexical_weight=0.7, rag_semantic_weight=0.8, rag_graph_weight=1.0,
                              rag_safety_threshold=case["thr"], rag_denoise_enabled=case["denoise"],
                              rag_denoise_alpha=case["alpha"])
        plan = None
        if case["weights"]:
            plan = SimpleNamespace(weights=dict(zip(("lexical", "semantic", "graph"), case["weights"])))
        lists = _lib_objects(case)
        first = {}
        for lst in lists:
            for o in lst:
                first.setdefault(o.chunk_id, o)
        got = GpuRRFFusion(cfg, engine=engine).fuse(*lists, query_plan=plan, top_k=case["top_k"])
        assert [(int(o.chunk_id[-12:]), o.rrf_score.hex(), [getattr(o, f).hex() for f in FIELDS]) for o in got] == \
               [(w["id"], w["rrf"], w["raw"]) for w in case["out"]]
        for o in got:
            assert o is first[o.chunk_id]                      # the first-seen object is the one mutated and returned
            assert o.final_score == o.rrf_score
            present = [n for c, n in enumerate(("lexical", "semantic", "graph"))
                       if int(o.chunk_id[-12:]) in (case["lists"][c] or [])]
            assert o.metadata["source_channels"] == present


def test_gpu_rrf_fusion_batch_equals_single_calls(engine, fusion_golden):
    cases = [c for c in fusion_golden["lib"] if (c["thr"], c["alpha"], c["denoise"], c["top_k"], c["weights"]) ==
             (0.6, 0.6, True, None, None)][:16]
    assert cases
    fus = GpuRRFFusion(engine=engine)
    got = fus.fuse_batch([tuple(_lib_objects(c)) for c in cases])
    for g, c in zip(got, cases):
        assert [(int(o.chunk_id[-12:]), o.rrf_score.hex()) for o in g] == [(w["id"], w["rrf"]) for w in c["out"]]


def test_rag1_rrf_fusion_replays_the_golden(engine, fusion_golden):
    for case in fusion_golden["rag1"]:
        lists = []
        for c, ids in enumerate(l for l in case["lists"] if l is not None):
            lists.append([SimpleNamespace(chunk_id=str(cid), similarity_score=0.1 * (c + 1), bm25_score=float(i),
                                          rrf_score=0.0, retrieval_method="vector") for i, cid in enumerate(ids)])
        got = rag1_rrf_fusion(engine, lists, k=case["rrf_k"])
        assert [(int(o.chunk_id), o.rrf_score.hex()) for o in got] == [(w["id"], w["rrf"]) for w in case["out"]]
        assert all(o.retrieval_method == "hybrid" for o in got)
        # best raw scores are kept on the first-seen object (hybrid_search.py:486-491)
        for o in got:
            seen = [x for lst in lists for x in lst if x.chunk_id == o.chunk_id]
            assert o is seen[0]


def test_fuse_two_channels_and_empty_inputs(engine):
    fus = GpuRRFFusion(engine=engine)
    mk = lambda cid: SimpleNamespace(chunk_id=cid, lexical_score=0.9, semantic_score=0.9, graph_score=0.0, rrf_score=0.0,
                                     final_score=0.0, metadata={})
    a, b = [mk("a"), mk("b"), mk("c")], [mk("b"), mk("d")]
    out = fus.fuse_two_channels(a, b, 1.0, 0.5)
    want = {"a": 1.0 * (1.0 / 61), "b": 1.0 * (1.0 / 62) + 0.5 * (1.0 / 61), "c": 1.0 * (1.0 / 63), "d": 0.5 * (1.0 / 62)}
    assert [o.chunk_id for o in out] == sorted(want, key=lambda k: -want[k])
    assert all(o.rrf_score == want[o.chunk_id] and o.final_score == o.rrf_score for o in out)
    assert fus.fuse([], [], []) == []
    assert fus.fuse_batch([]) == []
