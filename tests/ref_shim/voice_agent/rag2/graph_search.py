from triple_hybrid_rag_b200.retriever import GraphEdge, GraphNode, GraphSearchResult  # noqa: F401

__thr_shim__ = True
