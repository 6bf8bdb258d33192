"""BatchLenInfo -- same surface as /root/reference/e2e/punica-atom/punica/utils/cat_tensor.py:26-66: a step's batch is the
concatenation of the prefill prompts' tokens followed by one token per decoding sequence."""
from typing import Sequence

import torch


class BatchLenInfo:
    def __init__(self, prefills: Sequence[int], decode: int, indptr_device: torch.device, indptr_dtype: torch.dtype = torch.int32):
        self._prefills = list(prefills)
        self._decode = decode
        if len(self._prefills) > 0:
            cum = [0]
            for n in self._prefills:
                cum.append(cum[-1] + n)
            self._indptr = torch.tensor(cum, dtype=indptr_dtype, device=indptr_device)
            self._doff = cum[-1]
        else:
            self._indptr = None
            self._doff = 0

    @property
    def prefills(self) -> list:
        """Length of each prefill request."""
        return self._prefills

    @property
    def decode(self) -> int:
        """Number of decode requests."""
        return self._decode

    @property
    def doff(self) -> int:
        """Index of the first decode token == total length of the prefills."""
        return self._doff

    @property
    def indptr(self):
        """indptr[i] = sum(prefills[:i]); None without prefill."""
        return self._indptr
