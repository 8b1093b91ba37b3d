"""The reference's own hot-path test files, unmodified, against the drop-in.

tests/golden/ref_tests/*.py.gz are byte-for-byte copies of the reference's tests/test_rag2_triple_hybrid.py and
tests/test_rag2_retrieval.py (tests/golden/make_ref_tests.py; sha256 in MANIFEST.json).  They are unpacked into a
temporary directory and run in a child pytest whose import path resolves `voice_agent.*` to tests/ref_shim, i.e. to
triple_hybrid_rag_b200.retriever on this GPU: every `_fuse_rrf` / `_apply_safety` the reference's tests trigger runs
in thr_fuse_ranked / thr_safety.  RAG2_GRAPH_ENABLED=true as in the survey's run of the same files against the
reference itself (81/81 with the two RAG 1.0 / tool files, 51 of them in these two files)."""
import gzip
import json
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
HERE = Path(__file__).resolve().parent
ROOT = HERE.parent


def test_reference_test_files_pass_against_the_drop_in(tmp_path):
    manifest = json.loads((HERE / "golden" / "ref_tests" / "MANIFEST.json").read_text())
    total = 0
    for name, meta in manifest.items():
        with gzip.open(HERE / "golden" / "ref_tests" / (name + ".gz"), "rb") as fh:
            (tmp_path / name).write_bytes(fh.read())
        total += meta["tests"]
    env = dict(os.environ, RAG2_GRAPH_ENABLED="true", PYTHONDONTWRITEBYTECODE="1",
               PYTHONPATH=os.pathsep.join([str(HERE / "ref_shim"), str(HERE), str(ROOT)]))
    env.pop("PYTEST_CURRENT_TEST", None)
    cmd = [sys.executable, "-m", "pytest", "-p", "asyncio_shim", "-p", "no:cacheprovider", "-c", os.devnull,
           "--rootdir", str(tmp_path), "-q", "-rf"] + [str(tmp_path / n) for n in manifest]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=tmp_path, timeout=900)
    tail = out.stdout[-4000:] + out.stderr[-2000:]
    m = re.search(r"(\d+) passed", out.stdout)
    passed = int(m.group(1)) if m else 0
    print(f"reference tests through the drop-in: {passed}/{total} passed")
    assert out.returncode == 0 and passed == total, tail
