"""The hardware path of the two contraction kernels, read off the built library (no GPU needed): cuobjdump's SASS of
libthr.so must show tcgen05 MMAs (UTCHMMA), TMA tensor loads (UTMALDG), TMEM reads (LDTM) and tcgen05 commits (UTCBAR)
in dense_score_kernel and maxsim_kernel — a rebuild that silently fell back to mma.sync / plain loads would not."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

SO = Path(__file__).resolve().parents[1] / "triple_hybrid_rag_b200" / "lib" / "libthr.so"


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None or not SO.exists():
        pytest.skip("cuobjdump or the built library is not here")
    out = subprocess.run(["cuobjdump", "-sass", str(SO)], capture_output=True, text=True, timeout=300).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            funcs[name].append(line)
    assert "sm_100a" in out or "sm_100" in out
    return funcs


def _count(funcs, kernel, mnemonic):
    return {n: sum(mnemonic in l for l in ls) for n, ls in funcs.items() if kernel in n}


def test_dense_and_maxsim_kernels_are_tcgen05_tma_kernels(sass):
    for kernel in ("dense_score_kernel", "maxsim_kernel"):
        for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR"):
            counts = _count(sass, kernel, mnemonic)
            assert counts and all(c > 0 for c in counts.values()), (kernel, mnemonic, counts)
    # the CTA-pair instantiations issue cta_group::2 MMAs and multicast commits
    pair = [ls for n, ls in sass.items() if "dense_score_kernel" in n and "Li2E" in n]
    assert pair and all(any("UTCHMMA.2CTA" in l for l in ls) for ls in pair)
    # no legacy tensor-core path anywhere in the library
    assert not any(re.search(r"\b(HMMA|IMMA)\b", l) for ls in sass.values() for l in ls)


def test_seed_instantiation_stores_256_bits(sass):
    seed = {n: sum("STG.E.ENL2.256" in l for l in ls) for n, ls in sass.items()
            if "dense_score_kernel" in n and "Lb1E" in n}
    main = {n: sum("STG.E.ENL2.256" in l for l in ls) for n, ls in sass.items()
            if "dense_score_kernel" in n and "Lb0E" in n}
    assert seed and all(c >= 8 for c in seed.values()) and main and all(c == 0 for c in main.values())
