"""Test shim: resolves the reference's import paths (`voice_agent.rag2.retrieval`, `.query_planner`,
`.graph_search`, `voice_agent.config`) to the B200 drop-in, so that the reference's OWN test files
(tests/golden/ref_tests/*.py.gz) run unmodified against it on a box that has no reference checkout.
Test infrastructure only: nothing under triple_hybrid_rag_b200/ imports this package."""
__thr_shim__ = True
