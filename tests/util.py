"""Helpers shared by the GPU parity tests."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


def csr(lists: Sequence[Optional[Sequence]], dtype, device):
    """Ragged per-query lists -> (flat tensor, off int32 [B+1]); None entries count as empty."""
    off = [0]
    flat: List = []
    for l in lists:
        if l is not None:
            flat.extend(l)
        off.append(len(flat))
    if not flat:
        flat = [0]
    return (torch.tensor(flat, dtype=dtype, device=device), torch.tensor(off, dtype=torch.int32, device=device))
