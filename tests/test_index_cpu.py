"""Host logic of the BM25 index builder (no GPU): layout invariants of include/thr.h, agreement with
the oracle's term-major CSR, and concat(parts) == build(whole)."""
import numpy as np
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
