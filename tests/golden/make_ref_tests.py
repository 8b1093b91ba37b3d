#!/usr/bin/env python
"""Carry the reference's own hot-path test files to the GPU box as fixtures.

/root/reference does not exist on the GPU box, and the drop-in claim is "the reference's tests pass against it".
This script packs the two test files that exercise the path (patching _lexical_search / _semantic_search /
_graph_search / _expand_to_parents / _apply_safety by name, asserting RRF known answers, safety and refusal
behaviour) byte for byte into tests/golden/ref_tests/*.py.gz.  They are FIXTURES — outputs of the reference
checkout, like the golden vectors — not product source: tests/test_gpu_ref_tests.py unpacks them into a temporary
directory and runs them, unmodified, against the drop-in through tests/ref_shim (which resolves the reference's
import paths).  Re-run after a reference update:  python tests/golden/make_ref_tests.py
"""
import gzip
import hashlib
import json
from pathlib import Path

REF = Path("/root/reference/tests")
OUT = Path(__file__).resolve().parent / "ref_tests"
FILES = ["test_rag2_triple_hybrid.py", "test_rag2_retrieval.py"]

if __name__ == "__main__":
    OUT.mkdir(exist_ok=True)
    manifest = {}
    for name in FILES:
        raw = (REF / name).read_bytes()
        with gzip.GzipFile(OUT / (name + ".gz"), "wb", mtime=0) as fh:
            fh.write(raw)
        manifest[name] = {"sha256": hashlib.sha256(raw).hexdigest(), "bytes": len(raw),
                          "tests": raw.count(b"def test_")}
    (OUT / "MANIFEST.json").write_text(json.dumps(manifest, indent=1) + "\n")
    print(json.dumps(manifest, indent=1))
