"""Paged INT4 KV-cache bookkeeping -- same classes, attributes and layouts as
/root/reference/e2e/punica-atom/punica/utils/kvcache.py:6-127 (KvPoolInt4 / KvCacheInt4 / BatchedKvCacheInt4),
which is what the append / decode kernels consume:

    data  uint8   [capacity, num_layers, 2, num_heads, block_len, head_dim // 2]
    param float16 [capacity, num_layers, 2, num_heads, block_len, 2]            (scale, zero)
"""
from typing import Sequence

import torch


class KvPoolInt4:
    def __init__(self, num_layers: int, num_heads: int, head_dim: int, capacity: int, block_len: int,
                 device: torch.device):
        self._buf = torch.empty((capacity, num_layers, 2, num_heads, block_len, head_dim // 2), dtype=torch.uint8,
                                device=device)
        self._param = torch.empty((capacity, num_layers, 2, num_heads, block_len, 2), dtype=torch.float16,
                                  device=device)
        self._free = set(range(capacity))

    @property
    def buf(self):
        return self._buf

    @property
    def param(self):
        return self._param

    @property
    def num_layers(self):
        return self._buf.shape[1]

    @property
    def block_len(self):
        return self._buf.shape[4]

    @property
    def num_free_blocks(self):
        return len(self._free)

    def alloc_block(self) -> int:
        if not self._free:
            raise RuntimeError("KvPoolInt4: out of pages (capacity %d)" % self._buf.size(0))
        return self._free.pop()

    def free_block(self, idx: int):
        assert 0 <= idx < self._buf.size(0)
        assert idx not in self._free
        self._free.add(idx)


class KvCacheInt4:
    """Key-value cache of one sequence: a list of page ids plus the sequence length."""

    def __init__(self, pool: KvPoolInt4, init_len: int):
        if init_len < 0:
            raise ValueError("init_len must be non-negative")
        self._pool = pool
        blocks = (init_len + pool.block_len - 1) // pool.block_len
        self._indicies = [pool.alloc_block() for _ in range(blocks)]
        self._seqlen = init_len

    @property
    def pool(self) -> KvPoolInt4:
        return self._pool

    @property
    def seqlen(self) -> int:
        return self._seqlen

    @property
    def indicies(self) -> list:
        return self._indicies

    def acquire_one(self):
        """Reserve space for one more token (a new page when the last one is full)."""
        last_page_offset = (self._seqlen - 1) % self._pool.block_len + 1
        if last_page_offset == self._pool.block_len:
            self._indicies.append(self._pool.alloc_block())
        self._seqlen += 1

    def release(self):
        self._seqlen = 0
        for idx in self._indicies:
            self._pool.free_block(idx)
        self._indicies.clear()


class BatchedKvCacheInt4:
    """Page table of a batch in the CSR form the kernels read: indptr / indicies / last_page_offset (int32)."""

    def __init__(self, kv: Sequence[KvCacheInt4]):
        assert len(kv) > 0
        pool = kv[0].pool
        device = pool.buf.device
        indptr, indicies, last_page_offset = [0], [], []
        for c in kv:
            assert c.pool is pool
            indptr.append(indptr[-1] + len(c.indicies))
            indicies.extend(c.indicies)
            last_page_offset.append((c.seqlen - 1) % pool.block_len + 1)
        self.data = pool.buf
        self.param = pool.param
        self.indptr = torch.tensor(indptr, dtype=torch.int32, device=device)
        self.indicies = torch.tensor(indicies, dtype=torch.int32, device=device)
        self.last_page_offset = torch.tensor(last_page_offset, dtype=torch.int32, device=device)

    @property
    def page_size(self):
        return self.data.size(-2)
