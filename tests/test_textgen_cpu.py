"""CPU suite: the continuous-batching harness (atom_b200/textgen.py) -- workload generator pinned to the reference's own
(tests/golden/ref_py_request_set.npz <- benchmarks/bench_textgen.py:31-47), and the scheduler's invariants with a stand-in
model on a CPU page pool (no kernels involved: this is host logic)."""
import os

import numpy as np
import pytest
import torch

from atom_b200 import textgen as tg
from atom_b200.kvcache import KvPoolInt4


def test_request_set_matches_reference_generator(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_py_request_set.npz"))
    a, b = tg.generate_request_set(160, 2048), tg.generate_request_set(48, 512)
    assert np.array_equal(a.prompt_lens, g["p160"]) and np.array_equal(a.output_lens, g["o160"])
    assert np.array_equal(b.prompt_lens, g["p48"]) and np.array_equal(b.output_lens, g["o48"])
    assert a.prompt_lens.dtype == np.int32 and (a.output_lens >= 2).all() and (a.prompt_lens + a.output_lens <= 2048).all()


class _FakeLM:
    """logits such that argmax(row) = (input id * 7 + position-independent salt) % vocab; records what it was fed."""

    def __init__(self, vocab, batch_size, block_len):
        self.vocab, self.batch_size, self.block_len = vocab, batch_size, block_len
        self.calls = []

    def __call__(self, ids, blen, prefill_kv, decode_kv):
        assert ids.dtype in (torch.long, torch.int32) and ids.numel() == blen.doff + blen.decode
        assert len(blen.prefills) + blen.decode <= self.batch_size
        assert (prefill_kv is None) == (len(blen.prefills) == 0) and (decode_kv is None) == (blen.decode == 0)
        if prefill_kv is not None:      # page table of the prompts: enough pages, last page offset consistent with the length
            n_pages = (prefill_kv.indptr[1:] - prefill_kv.indptr[:-1]).tolist()
            for plen, pages, last in zip(blen.prefills, n_pages, prefill_kv.last_page_offset.tolist()):
                assert pages == -(-plen // self.block_len) and last == (plen - 1) % self.block_len + 1
        if decode_kv is not None:
            nnz = int(decode_kv.indptr[-1])          # a static (graph) page table is over-allocated: only the first nnz count
            assert decode_kv.indptr.numel() == blen.decode + 1 and decode_kv.indicies.numel() >= nnz
            assert len(set(decode_kv.indicies[:nnz].tolist())) == nnz                          # no page shared
        self.calls.append((list(blen.prefills), blen.decode, None if decode_kv is None else
                           ((decode_kv.indptr[1:] - decode_kv.indptr[:-1] - 1) * self.block_len + decode_kv.last_page_offset).tolist()))
        nxt = (ids.long() * 7 + 3) % self.vocab
        return torch.nn.functional.one_hot(nxt, self.vocab).float(), None


@pytest.mark.parametrize("batch_size,maxlen", [(4, 96), (7, 64)])
def test_scheduler_invariants(batch_size, maxlen):
    rs = tg.generate_request_set(3 * batch_size + 2, maxlen)
    block = 16
    pool = KvPoolInt4(1, 2, 128, tg.pool_capacity(batch_size, maxlen, block), block, torch.device("cpu"))
    cap = pool.num_free_blocks
    lm = _FakeLM(97, batch_size, block)
    res = tg.run_textgen(lm, rs, tg.TextGenConfig(batch_size), pool, torch.device("cpu"), keep_tokens=True)
    assert pool.num_free_blocks == cap                                        # every page returned
    assert [len(t) for t in res.tokens] == rs.output_lens.tolist()            # every request generated its length
    # greedy chain of the stand-in model: each token follows from the previous one
    for toks in res.tokens:
        assert all(toks[i + 1] == (toks[i] * 7 + 3) % 97 for i in range(len(toks) - 1))
    # token accounting: a request costs prompt + (output - 1) rows; the steps add up to it
    rows = sum(sum(p) + d for p, d, _ in lm.calls)
    assert rows == int(rs.prompt_lens.sum() + rs.output_lens.sum()) - len(rs)
    assert res.steps == len(lm.calls)
    # FCFS: prompts are admitted in request order
    admitted = [p for call in lm.calls for p in call[0]]
    assert admitted == rs.prompt_lens.tolist()
    # a decoding sequence's KV length grows by exactly one per step (the slot for this step's token is already reserved)
    assert all(l >= 2 for _, _, lens in lm.calls if lens for l in lens)
    assert (res.encode_latency > 0).all() and (res.decode_latency >= res.encode_latency).all()
    rep = tg.report(rs, tg.TextGenConfig(batch_size), res)
    assert rep["total_new_tokens"] == int(rs.output_lens.sum()) and rep["throughput_tokens_per_s"] > 0


def test_admission_waits_for_pages_and_reports_impossible_requests():
    rs = tg.RequestSet(np.array([40, 40, 40], np.int32), np.array([5, 5, 5], np.int32))
    block = 16
    pool = KvPoolInt4(1, 1, 128, 9, block, torch.device("cpu"))               # room for two such requests at a time, not three
    lm = _FakeLM(31, 3, block)
    res = tg.run_textgen(lm, rs, tg.TextGenConfig(3), pool, torch.device("cpu"), keep_tokens=True)
    assert [len(t) for t in res.tokens] == [5, 5, 5] and pool.num_free_blocks == 9
    assert lm.calls[0][0] == [40, 40]                                         # the third prompt had to wait
    tiny = KvPoolInt4(1, 1, 128, 2, block, torch.device("cpu"))
    with pytest.raises(RuntimeError, match="KV pool too small"):
        tg.run_textgen(lm, rs, tg.TextGenConfig(3), tiny, torch.device("cpu"))
    with pytest.raises(RuntimeError, match="out of pages"):
        [tiny.alloc_block() for _ in range(3)]


def test_decode_graph_runner_static_buffers_equal_the_eager_page_tables():
    """DecodeGraphRunner with capture=False: same packed int32 buffer, same model calls, no CUDA graph -- the host logic
    of the graphed decode path.  Tokens, step count and page accounting must equal the plain loop's."""
    rs = tg.generate_request_set(11, 80)
    block, bs = 16, 4
    outs = []
    for use_runner in (False, True):
        pool = KvPoolInt4(1, 2, 128, tg.pool_capacity(bs, 80, block), block, torch.device("cpu"))
        lm = _FakeLM(97, bs, block)
        runner = tg.DecodeGraphRunner(lm, pool, "cpu", max_pages_per_seq=80 // block + 1, capture=False) if use_runner else None
        res = tg.run_textgen(lm, rs, tg.TextGenConfig(bs), pool, torch.device("cpu"), keep_tokens=True, decode_runner=runner)
        outs.append((res.tokens, res.steps, [(c[0], c[1], c[2]) for c in lm.calls], res.graphed_steps))
        assert pool.num_free_blocks == tg.pool_capacity(bs, 80, block)
    assert outs[0][:3] == outs[1][:3]                     # identical tokens, steps and per-step (prefills, decode, kv lengths)
    assert outs[0][3] == 0 and 0 < outs[1][3] < outs[1][1]
    with pytest.raises(RuntimeError, match="sized for"):
        pool = KvPoolInt4(1, 2, 128, 64, block, torch.device("cpu"))
        tg.run_textgen(_FakeLM(97, bs, block), rs, tg.TextGenConfig(bs), pool, torch.device("cpu"),
                       decode_runner=tg.DecodeGraphRunner(_FakeLM(97, bs, block), pool, "cpu", max_pages_per_seq=1, capture=False))


def test_prompt_ids_follow_the_models_vocabulary():
    """round-1 GPU failure: prompts were drawn from the default 32000-entry vocabulary for a 128-entry embedding table."""
    class _Head:
        out_features = 23

    class _SmallVocabLM(_FakeLM):
        lm_head = _Head()

        def __call__(self, ids, blen, prefill_kv, decode_kv):
            assert int(ids.max()) < 23, "prompt id outside the model's embedding table"
            return super().__call__(ids, blen, prefill_kv, decode_kv)

    rs = tg.generate_request_set(5, 64)
    pool = KvPoolInt4(1, 2, 128, tg.pool_capacity(2, 64, 16), 16, torch.device("cpu"))
    res = tg.run_textgen(_SmallVocabLM(23, 2, 16), rs, tg.TextGenConfig(2), pool, torch.device("cpu"), keep_tokens=True)
    assert all(0 <= t < 23 for toks in res.tokens for t in toks)
