"""Rows of the reference's tables -> the inputs of ResidentIndex (index builder input).

The reference keeps the corpus in Postgres: `rag_child_chunks` (database/migrations/20260114_rag2_schema.sql:104-152,
written by src/voice_agent/rag2/ingest.py:434-449), `rag_parent_chunks` (:63-90) and `rag_documents` (:17-58, the
`collection` column the search RPCs join on, :366-370 / :402-406).  A deployment exports those tables once (a
PostgREST `select("*")`, a COPY to JSON lines, a pg_dump read back — anything that yields one dict per row) and hands
the records to `from_tables`; what comes back is exactly what `ResidentIndex(engine, chunks, embeddings, parents)`
takes.  Host-side, no GPU needed; no scoring happens here.

Column semantics follow the RPCs:
  * `c.org_id = p_org_id`            -> `org_id=` keeps one tenant's rows (None: every row given);
  * `d.collection = p_collection`    -> a chunk's collection is its DOCUMENT's (`rag_documents.collection`);
  * `c.embedding_1024 IS NOT NULL`   -> only the semantic RPC has this predicate: a chunk without an embedding can
                                        still match lexically.  ResidentIndex keeps one embedding row per chunk, so
                                        `missing_embedding=` decides: "error" (default: the reference's ingest never
                                        stores a chunk whose embedding failed, ingest.py:429-432), "drop" (the chunk
                                        leaves both channels; reported in the returned stats) or "zero" (a zero row:
                                        lexical behaviour kept, semantic score 0 instead of absent).
pgvector values arrive as text ("[0.12,-0.5,...]", PostgREST's rendering), as lists, or as anything numpy can read.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, List, Mapping, Optional, Tuple

import numpy as np
import torch

CHILD_COLUMNS = ("id", "parent_id", "document_id", "text", "page", "modality")   # what the RPCs return per hit


def parse_pgvector(v: Any, dim: Optional[int] = None) -> Optional[np.ndarray]:
    """One `vector(n)` / `halfvec(n)` value -> float32 array; None for SQL NULL."""
    if v is None:
        return None
    if isinstance(v, str):
        s = v.strip()
        if not s or s.lower() == "null":
            return None
        if s[0] in "[({" and s[-1] in "])}":
            s = s[1:-1]
        a = np.array([float(x) for x in s.split(",") if x.strip()], dtype=np.float32)
    elif isinstance(v, torch.Tensor):
        a = v.detach().to(torch.float32).cpu().numpy().reshape(-1)
    else:
        a = np.asarray(v, dtype=np.float32).reshape(-1)
    if dim is not None and a.shape[0] != dim:
        raise ValueError(f"embedding has {a.shape[0]} dimensions, expected {dim}")
    if not np.isfinite(a).all():
        raise ValueError("embedding holds a non-finite value")
    return a


@dataclass
class ExportStats:
    children_seen: int = 0
    children_kept: int = 0
    other_org: int = 0
    missing_embedding: int = 0
    unknown_document: int = 0
    duplicate_ids: int = 0
    dim: int = 0
    collections: Dict[Optional[str], int] = field(default_factory=dict)


def from_tables(child_rows: Iterable[Mapping[str, Any]],
                document_rows: Optional[Iterable[Mapping[str, Any]]] = None,
                parent_rows: Optional[Iterable[Mapping[str, Any]]] = None,
                org_id: Optional[str] = None, embedding_column: str = "embedding_1024",
                missing_embedding: str = "error"
                ) -> Tuple[List[Dict[str, Any]], torch.Tensor, Dict[str, Dict[str, Any]], ExportStats]:
    """(chunks, embeddings, parents, stats) for ResidentIndex from exported table rows.

    chunks[i] carries child_id / parent_id / document_id / text / page / modality (the columns the search RPCs return,
    20260114_rag2_schema.sql:347-355) plus `collection`; embeddings[i] is its row (float32, [n, dim]); parents maps a
    parent id to {"id", "text", "section_heading"} — the columns `_expand_to_parents` selects (retrieval.py:389-392).
    Rows keep the order they were given in (row index = the dense / lexical doc id on the device), so an export ordered
    by (document_id, parent_id, index_in_parent) keeps a document's chunks adjacent."""
    if missing_embedding not in ("error", "drop", "zero"):
        raise ValueError("missing_embedding must be 'error', 'drop' or 'zero'")
    collection_of: Dict[str, Optional[str]] = {}
    have_documents = document_rows is not None
    for d in document_rows or ():
        if org_id is not None and d.get("org_id") is not None and str(d["org_id"]) != str(org_id):
            continue
        collection_of[str(d["id"])] = d.get("collection")
    stats = ExportStats()
    chunks: List[Dict[str, Any]] = []
    vecs: List[Optional[np.ndarray]] = []
    seen = set()
    for c in child_rows:
        stats.children_seen += 1
        if org_id is not None and c.get("org_id") is not None and str(c["org_id"]) != str(org_id):
            stats.other_org += 1
            continue
        cid = str(c["id"])
        if cid in seen:            # `id` is the primary key: a second row with it is an export mistake, not data
            stats.duplicate_ids += 1
            continue
        doc = str(c["document_id"])
        if have_documents and doc not in collection_of:
            stats.unknown_document += 1     # the RPCs JOIN rag_documents: a chunk without its document row is invisible
            continue
        vec = parse_pgvector(c.get(embedding_column), stats.dim or None)
        if vec is None:
            stats.missing_embedding += 1
            if missing_embedding == "error":
                raise ValueError(f"child chunk {cid} has no {embedding_column} (pass missing_embedding='drop' or 'zero')")
            if missing_embedding == "drop":
                continue
        elif not stats.dim:
            stats.dim = int(vec.shape[0])
        seen.add(cid)
        collection = collection_of.get(doc) if have_documents else c.get("collection")
        chunks.append({"child_id": cid, "parent_id": str(c["parent_id"]), "document_id": doc,
                       "text": c.get("text") or "", "page": c.get("page"), "modality": c.get("modality") or "text",
                       "collection": collection})
        vecs.append(vec)
        stats.collections[collection] = stats.collections.get(collection, 0) + 1
    if chunks and not stats.dim:
        raise ValueError(f"no row carries a {embedding_column}")
    X = np.zeros((len(chunks), stats.dim), dtype=np.float32)
    for i, v in enumerate(vecs):
        if v is not None:
            X[i] = v
    stats.children_kept = len(chunks)
    parents: Dict[str, Dict[str, Any]] = {}
    wanted = {c["parent_id"] for c in chunks}
    for p in parent_rows or ():
        pid = str(p["id"])
        if pid in wanted:
            parents[pid] = {"id": pid, "text": p.get("text") or "", "section_heading": p.get("section_heading")}
    return chunks, torch.from_numpy(X), parents, stats


def read_jsonl(path) -> List[Dict[str, Any]]:
    """A table exported as JSON lines (`COPY (SELECT row_to_json(t) FROM rag_child_chunks t) TO ...`)."""
    import json
    out = []
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            line = line.strip()
            if line:
                out.append(json.loads(line))
    return out
