from triple_hybrid_rag_b200.retriever import QueryPlan  # noqa: F401  (field-for-field the reference's, query_planner.py:23-50)

__thr_shim__ = True
