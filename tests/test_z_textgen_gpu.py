"""GPU suite (-m gpu), sorted last: the continuous-batching harness end to end on the kernels with a small random model --
mixed prefill + decode steps, page allocation / release through KvPoolInt4, greedy tokens read back every step."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(seed_pages, graphs=False):
    from atom_b200 import textgen as tg
    from atom_b200.kvcache import KvPoolInt4
    from atom_b200.llama import LinearInt4, LlamaConfig, LlamaForCausalLM
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    torch.set_default_dtype(torch.float16)
    try:
        with dev:
            model = LlamaForCausalLM(LlamaConfig(hidden_size=512, intermediate_size=1024, num_attention_heads=4,
                                                 num_hidden_layers=2, vocab_size=128))
    finally:
        torch.set_default_dtype(torch.float32)
    for i, m in enumerate(mod for mod in model.modules() if isinstance(mod, LinearInt4)):
        m.init_random(i)
    rs = tg.generate_request_set(7, 72)
    pool = KvPoolInt4(2, 4, 128, tg.pool_capacity(3, 72, 16) + seed_pages, 16, dev)
    for _ in range(seed_pages):            # shift the page ids the requests will get
        pool.alloc_block()
    free0 = pool.num_free_blocks
    runner = tg.DecodeGraphRunner(model.eval(), pool, dev, max_pages_per_seq=72 // 16 + 1) if graphs else None
    res = tg.run_textgen(model.eval(), rs, tg.TextGenConfig(3), pool, dev, sync=torch.cuda.synchronize, keep_tokens=True,
                         decode_runner=runner)
    assert (res.graphed_steps > 0) == graphs
    assert pool.num_free_blocks == free0
    assert [len(t) for t in res.tokens] == rs.output_lens.tolist()
    assert all(0 <= t < 128 for toks in res.tokens for t in toks)
    return res.tokens


@pytest.mark.timeout(300)
def test_textgen_harness_end_to_end_and_page_placement_independent():
    a = _run(0)
    b = _run(5)          # same requests on different physical pages: greedy tokens must not change
    assert a == b


@pytest.mark.timeout(300)
def test_textgen_decode_steps_from_cuda_graphs_give_the_same_tokens():
    assert _run(0, graphs=True) == _run(0)
