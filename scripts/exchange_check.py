"""Sharded search, pushed exchange vs NCCL all-gather vs the unsharded index: bit-identical outputs.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/exchange_check.py"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, ".")
from triple_hybrid_rag_b200 import synth
from triple_hybrid_rag_b200.engine import Engine
from triple_hybrid_rag_b200.index import BM25Index, pack_queries
from triple_hybrid_rag_b200.pipeline import TripleHybridSearcher, shard_bounds
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, D, B, k, V = 400_000, 256, 96, 100, 20_000
eng = Engine(local); dev = eng.device
_b = shard_bounds(N, world); lo, hi = _b[rank], _b[rank + 1]
doc, term, tf, L = synth.bm25_block_coo(0, N, V=V)
full = BM25Index.build(doc, term, tf, L, V, blk_docs=2048)
m = (doc >= lo) & (doc < hi)
loc = BM25Index.build(doc[m] - lo, term[m], tf[m], L[lo:hi], V, blk_docs=2048, avgdl=full.avgdl, idf=full.idf)
X = synth.dense_rows(0, N, D)
Q = synth.dense_queries(B, D, X[: N // 8]).to(dev)
qt, qo = pack_queries(synth.bm25_queries(B, V=V, min_rank=50), dev)
graph = torch.randint(0, N, (B, 50), generator=torch.Generator().manual_seed(7)).to(dev)
# rerank stage: a token store over the whole corpus (Td = 64), each rank holding the rows of its own chunks
gt = torch.Generator().manual_seed(5)
store = torch.randn((N // 8, 64, 128), generator=gt)
store = (store / store.norm(dim=-1, keepdim=True)).to(torch.bfloat16)      # chunk id -> row id % (N // 8), on every rank
Qtok = torch.randn((B, 32, 128), generator=gt)
Qtok = (Qtok / Qtok.norm(dim=-1, keepdim=True)).to(torch.bfloat16).to(dev)
P = N // 8
def tup(o):
    return (o.ids.cpu(), o.rrf.cpu(), o.count.cpu(), o.sem_ids.cpu(), o.lex_ids.cpu(), o.lex_scores.cpu(),
            o.rr_ids.cpu(), o.rr_score.cpu(), o.rr_keep.cpu(), o.refused.cpu(), o.max_score.cpu())
outs = {}
for mode in ("peer", "nccl"):
    s = TripleHybridSearcher(eng, group=dist.group.WORLD, exchange=mode)
    s.set_dense(X[lo:hi].to(dev), id_base=lo); s.set_bm25(loc, id_base=lo)
    s.set_token_store(store.to(dev), lo, hi, period=P, row_off=lo % P)     # row = global id % P on every rank
    for it in range(5):   # several steps: both buffer halves and growing sequence numbers; a smaller batch in between
        if it == 2:
            s.search(Q[:40], *pack_queries(synth.bm25_queries(B, V=V, min_rank=50)[:40], dev), graph[:40], k_sem=k, k_lex=k, top_k=k)
        o = s.search(Q, qt, qo, graph, k_sem=k, k_lex=k, top_k=k)
        o = s.rerank(o, Qtok, 60, 0.5, 0.9, 10)
    eng.sync()
    outs[mode] = tup(o)
    if rank == 0:
        print(mode, "->", s.exchange_mode, flush=True)
same = all(torch.equal(a, b) for a, b in zip(outs["peer"], outs["nccl"]))
ok = torch.tensor([int(same)], device=dev)
if rank == 0:   # unsharded reference on the same GPU
    s1 = TripleHybridSearcher(eng)
    s1.set_dense(X.to(dev)); s1.set_bm25(full)
    s1.set_token_store(store.to(dev), 0, N, period=P)
    o = s1.search(Q, qt, qo, graph, k_sem=k, k_lex=k, top_k=k)
    o = s1.rerank(o, Qtok, 60, 0.5, 0.9, 10)
    eng.sync()
    ref = tup(o)
    ok[0] = int(same and all(torch.equal(a, b) for a, b in zip(outs["peer"], ref)))
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"exchange check (world {world}):", "OK — pushed == NCCL == unsharded, bit for bit, fused lists and reranked lists" if int(ok) else "MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(ok) else 1)
