"""K4 parity: tcgen05 MaxSim vs the fp64 oracle, within 1e-3 relative (north_star's tolerance)."""
import numpy as np
import pytest
import torch

from oracle import maxsim as om
from triple_hybrid_rag_b200 import synth

pytestmark = pytest.mark.gpu
REL_TOL = 1e-3


def _check(engine, Q, D, cand, q_len=None, d_len=None):
    dev = engine.device
    out = engine.maxsim(Q.to(dev), D.to(dev), cand.to(dev),
                        None if q_len is None else q_len.to(dev), None if d_len is None else d_len.to(dev))
    engine.sync()
    want = om.maxsim(Q.float().numpy(), D.float().numpy(), cand.numpy(),
                     None if q_len is None else q_len.numpy(), None if d_len is None else d_len.numpy())
    got = out.cpu().numpy().astype(np.float64)
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.allclose(got[fin], want[fin], rtol=REL_TOL, atol=1e-4), np.abs(got[fin] - want[fin]).max()


@pytest.mark.parametrize("B,C,Tq,Td", [(4, 37, 32, 128), (3, 300, 128, 128), (2, 50, 17, 64), (70, 5, 32, 128)])
def test_maxsim_matches_oracle(engine, B, C, Tq, Td):
    Q, D, cand = synth.maxsim_tokens(B, C, Tq=Tq, Td=Td)
    _check(engine, Q, D, cand)


def test_maxsim_ragged_lengths_and_invalid_candidates(engine):
    Q, D, cand = synth.maxsim_tokens(5, 40, Tq=64, Td=128)
    g = torch.Generator().manual_seed(1)
    q_len = torch.randint(1, 65, (5,), generator=g, dtype=torch.int32)
    d_len = torch.randint(0, 129, (D.shape[0],), generator=g, dtype=torch.int32)
    cand = cand[:, torch.randperm(40, generator=g)].contiguous()
    cand[0, 3] = -1
    cand[4, 0] = D.shape[0] + 5
    cand[2] = cand[1]  # candidates shared between queries
    _check(engine, Q, D, cand, q_len, d_len)
