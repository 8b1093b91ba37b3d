"""Host logic of the BM25 index builder (no GPU): layout invariants of include/thr.h, agreement with
the oracle's term-major CSR, and concat(parts) == build(whole)."""
import numpy as np
import pytest
import torch

from oracle import bm25 as ob
from triple_hybrid_rag_b200 import synth
from triple_hybrid_rag_b200.index import BM25Index, pack_queries


def _coo(n_docs, V, block=0):
    return synth.bm25_block_coo(block, n_docs, V=V)


def test_build_matches_oracle_csr_and_skip_invariants():
    n_docs, V, R = 5000, 800, 1024
    doc, term, tf, L = _coo(n_docs, V)
    idx = BM25Index.build(doc, term, tf, L, V, blk_docs=R)
    orc = ob.CsrIndex.from_coo(doc.numpy(), term.numpy(), tf.numpy(), L.numpy(), V)
    n_blk = idx.n_blk
    assert n_blk == 5 and idx.skip.numel() == V * n_blk + 1
    skip = idx.skip.numpy()
    assert np.array_equal(skip[::n_blk][: V + 1], orc.indptr)            # the CSR indptr is skip[::n_blk]
    assert (np.diff(skip) >= 0).all() and skip[-1] == idx.nnz
    post = idx.postings.numpy()
    assert np.array_equal(post[: idx.nnz, 0].astype(np.int64), orc.doc)
    assert np.array_equal(post[: idx.nnz, 1].view(np.float32), orc.imp)
    assert (post[idx.nnz:] == 0).all()                                    # 16 B tail padding
    assert np.array_equal(idx.idf.numpy().view(np.uint32), orc.idf.view(np.uint32))
    # every posting of (term t, range r) lies in skip[t*n_blk+r : t*n_blk+r+1] and docs ascend inside it
    d = post[: idx.nnz, 0].astype(np.int64)
    for t in (0, 1, 17, V - 1):
        for r in range(n_blk):
            seg = d[skip[t * n_blk + r]: skip[t * n_blk + r + 1]]
            assert ((seg // R) == r).all() and (np.diff(seg) > 0).all()


def test_concat_equals_whole_build():
    V, R = 500, 1024
    parts, coo = [], []
    base = 0
    for b, rows in enumerate((2048, 1024, 700)):
        doc, term, tf, L = _coo(rows, V, block=b)
        parts.append(BM25Index.build(doc, term, tf, L, V, blk_docs=R, avgdl=200.0))
        coo.append((doc + base, term, tf, L))
        base += rows
    whole = BM25Index.build(torch.cat([c[0] for c in coo]), torch.cat([c[1] for c in coo]),
                            torch.cat([c[2] for c in coo]), torch.cat([c[3] for c in coo]), V, blk_docs=R,
                            avgdl=200.0)
    cat = BM25Index.concat(parts)
    assert cat.n_docs == whole.n_docs == base and cat.nnz == whole.nnz
    assert torch.equal(cat.skip, whole.skip)
    assert torch.equal(cat.postings, whole.postings)
    assert torch.equal(cat.df, whole.df)
    assert torch.equal(cat.idf.view(torch.int32), whole.idf.view(torch.int32))


def test_pack_queries_and_algorithmic_bytes():
    doc, term, tf, L = _coo(3000, 300)
    idx = BM25Index.build(doc, term, tf, L, 300, blk_docs=1024)
    qs = [[1, 2], [], [299, 5, 7]]
    qt, qo = pack_queries(qs, "cpu")
    assert qo.tolist() == [0, 2, 2, 5] and qt.tolist() == [1, 2, 299, 5, 7]
    df = idx.df
    want = sum(int(df[t]) * 8 + 8 * idx.n_blk for q in qs for t in q)
    assert idx.algorithmic_bytes(qs) == want


def test_save_load_roundtrip(tmp_path):
    doc, term, tf, L = _coo(2500, 200)
    idx = BM25Index.build(doc, term, tf, L, 200, blk_docs=512)
    f = tmp_path / "idx.pt"
    idx.save(f)
    back = BM25Index.load(f)
    for name in ("skip", "postings", "df"):
        assert torch.equal(getattr(idx, name), getattr(back, name))
    assert torch.equal(idx.idf.view(torch.int32), back.idf.view(torch.int32))
    assert (idx.n_docs, idx.blk_docs, idx.V, idx.nnz, idx.k1, idx.b, idx.avgdl) == \
           (back.n_docs, back.blk_docs, back.V, back.nnz, back.k1, back.b, back.avgdl)


def test_tokenizer_hook_portuguese():
    """The linguistic hook in front of the lexical channel: stop words dropped, plurals folded,
    and the same callable serves corpus and queries."""
    from triple_hybrid_rag_b200.retriever import Tokenizer, tokenize
    tk = Tokenizer.portuguese()
    assert tk("Os contratos de pagamento e as condições") == ["contrato", "pagamento", "condição"]
    assert tk("contrato") == tk("CONTRATOS") and tk("de para com") == []
    assert Tokenizer()("Os contratos") == tokenize("Os contratos") == ["os", "contratos"]


def test_pad_dim_keeps_dot_products():
    import torch
    from triple_hybrid_rag_b200.retriever import pad_dim
    x = torch.randn(5, 4000).to(torch.bfloat16)      # the RAG 1.0 halfvec(4000) width
    p = pad_dim(x)
    assert p.shape == (5, 4032) and torch.equal(p[:, :4000], x) and not p[:, 4000:].any()
    assert torch.equal(pad_dim(p), p)


def _table_rows(n=20, D=8, seed=3):
    """Rows as the reference's ingest writes them (src/voice_agent/rag2/ingest.py:434-449) and as PostgREST returns
    them: uuid-like strings, the vector as text."""
    g = np.random.default_rng(seed)
    docs = [{"id": f"doc-{j}", "org_id": "org-1", "collection": ["contracts", "policies", None][j % 3], "title": f"t{j}"}
            for j in range(4)]
    parents = [{"id": f"par-{j}", "document_id": f"doc-{j // 2}", "org_id": "org-1", "index_in_document": j % 2,
                "text": f"parent text {j}", "section_heading": f"h{j}" if j % 2 else None} for j in range(8)]
    children = []
    for i in range(n):
        v = g.standard_normal(D).astype(np.float32)
        children.append({"id": f"ch-{i}", "parent_id": f"par-{i % 8}", "document_id": f"doc-{(i % 8) // 2}",
                         "org_id": "org-1", "index_in_parent": i // 8, "text": f"texto numero {i}", "token_count": 3,
                         "page": 1 + i % 5, "modality": "table" if i % 7 == 0 else "text", "content_hash": f"h{i}",
                         "metadata": {}, "embedding_1024": "[" + ",".join(repr(float(x)) for x in v) + "]",
                         "_vec": v})
    return children, docs, parents


def test_export_from_table_rows():
    """rag_child_chunks / rag_documents / rag_parent_chunks rows -> ResidentIndex inputs, with the
    predicates of the search RPCs (20260114_rag2_schema.sql:366-370): one org, the document's collection."""
    from triple_hybrid_rag_b200.export import from_tables, parse_pgvector
    children, docs, parents = _table_rows()
    children.append({**children[0], "id": "ch-other", "org_id": "org-2"})          # another tenant's row
    children.append({**children[1], "id": "ch-orphan", "document_id": "doc-gone"})  # its document row is missing
    children.append(dict(children[2]))                                               # the same primary key twice
    chunks, X, pmap, st = from_tables(children, docs, parents, org_id="org-1")
    assert st.children_seen == 23 and st.children_kept == 20 and st.other_org == 1 and st.unknown_document == 1
    assert st.duplicate_ids == 1 and st.dim == 8 and X.shape == (20, 8) and X.dtype == torch.float32
    for i, c in enumerate(chunks):
        src = children[i]
        assert c["child_id"] == src["id"] and c["parent_id"] == src["parent_id"] and c["document_id"] == src["document_id"]
        assert c["text"] == src["text"] and c["page"] == src["page"] and c["modality"] == src["modality"]
        assert c["collection"] == docs[(i % 8) // 2]["collection"]          # the DOCUMENT's collection
        assert np.array_equal(X[i].numpy(), src["_vec"])                    # text rendering round-trips float32 exactly
    assert set(pmap) == {f"par-{j}" for j in range(8)} and pmap["par-3"] == {"id": "par-3", "text": "parent text 3",
                                                                             "section_heading": "h3"}
    assert st.collections == {"contracts": 10, "policies": 6, None: 4}
    # vector renderings
    assert parse_pgvector(None) is None and parse_pgvector("null") is None
    assert np.array_equal(parse_pgvector([1, 2.5]), np.array([1, 2.5], dtype=np.float32))
    assert np.array_equal(parse_pgvector(" (1, -2) "), np.array([1, -2], dtype=np.float32))
    with pytest.raises(ValueError):
        parse_pgvector("[1,2,3]", dim=4)
    with pytest.raises(ValueError):
        parse_pgvector([1.0, float("nan")])


def test_export_missing_embeddings_and_jsonl(tmp_path):
    import json
    from triple_hybrid_rag_b200.export import from_tables, read_jsonl
    children, docs, parents = _table_rows(n=6)
    for c in children:
        c.pop("_vec")
    children[4]["embedding_1024"] = None
    with pytest.raises(ValueError, match="ch-4"):
        from_tables(children, docs, parents)
    chunks, X, _, st = from_tables(children, docs, parents, missing_embedding="drop")
    assert [c["child_id"] for c in chunks] == ["ch-0", "ch-1", "ch-2", "ch-3", "ch-5"] and st.missing_embedding == 1
    chunks, X, _, st = from_tables(children, docs, parents, missing_embedding="zero")
    assert len(chunks) == 6 and float(X[4].abs().sum()) == 0.0 and float(X[3].abs().sum()) > 0
    # without document rows the chunk's own `collection` key (if the export denormalised it) is used
    chunks, _, _, _ = from_tables([{**children[0], "collection": "x"}])
    assert chunks[0]["collection"] == "x"
    p = tmp_path / "children.jsonl"
    p.write_text("\n".join(json.dumps(c) for c in children) + "\n\n", encoding="utf-8")
    assert read_jsonl(p) == children
