"""Multi-GPU correctness of the tensor-parallel path (run under torchrun, N >= 2):
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/tp_check.py
1. the one-shot push all-reduce (csrc/comm_kernels.cuh) against the FP32 sum of the all-gathered inputs and against
   ncclAllReduce, eagerly and replayed from a CUDA graph (the epoch counter lives on the device);
2. a TP decoder layer step with the push all-reduce against the same layer with NCCL (same weights, same KV).
Prints one JSON line per check on rank 0; exit status 1 if any check fails."""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atom_b200.comm import NcclAllReduce, PushAllReduce
from atom_b200.kvcache import BatchedKvCacheInt4, KvCacheInt4, KvPoolInt4
from atom_b200.llama import LlamaConfig
from atom_b200.tp import TPLlamaDecoderLayer

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok_all = True


def report(name, ok, **kw):
    global ok_all
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok_all = ok_all and bool(flag.item())
    if rank == 0:
        print(json.dumps({"check": name, "ok": bool(flag.item()), "world": world, **kw}), flush=True)


# ---------------------------------------------------------------- 1. all-reduce
numel = 32 * 8192
push = PushAllReduce(numel, dev)
nccl = NcclAllReduce()
worst = 0.0
for it, n in enumerate([8, 4096, 32 * 5120, numel, numel, 2048, numel]):
    torch.manual_seed(1000 * it + rank)
    x = (torch.randn(n, device=dev) * 3).half()
    gathered = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(gathered, x)
    want = torch.stack([g.float() for g in gathered]).sum(0)         # rank order, FP32: what the kernel computes
    got = push(x)
    ref = nccl(x.clone())
    err = (got.float() - want).abs().max().item()
    worst = max(worst, err / max(1.0, want.abs().max().item()))
    exact = torch.equal(got, want.half())
    close = torch.allclose(got.float(), ref.float(), rtol=2e-3, atol=2e-2)
    report(f"push_allreduce_eager_n{n}", exact and close, max_abs_err_vs_fp32_sum=err)

# graph replay: 25 all-reduces per replay, input changes between replays
x = torch.zeros(numel, device=dev, dtype=torch.float16)
st = torch.cuda.Stream(dev)
with torch.cuda.stream(st):
    y = push(x)
    st.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        y = x
        for _ in range(25):
            y = push(y) * (1.0 / world)                              # keeps the magnitude: mean over ranks
    for rep in range(4):
        torch.manual_seed(77 * rep + rank)
        x.copy_((torch.randn(numel, device=dev) * 2).half())
        g.replay()
        st.synchronize()
        gathered = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(gathered, x)
        z = None
        # after the first all-reduce every rank holds the same tensor, so the remaining 24 are x -> x (up to FP16 rounding)
        first = (torch.stack([t.float() for t in gathered]).sum(0).half() * (1.0 / world)).half()
        z = first
        for _ in range(24):
            z = ((z.float() * world).half() * (1.0 / world)).half()
        report(f"push_allreduce_graph_replay_{rep}", torch.equal(y, z), max_abs_diff=(y.float() - z.float()).abs().max().item())
g = None

# latency of one all-reduce of the decode step's size (batch 32 x hidden 8192 FP16 = 512 KiB): 50 chained calls in one graph
def time_allreduce(ar, n, chain=50, reps=20):
    xs = torch.randn(n, device=dev).half()
    with torch.cuda.stream(st):
        y = ar(xs.clone())
        st.synchronize()
        dist.barrier()
        gg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gg, stream=st):
            y = xs
            for _ in range(chain):
                y = ar(y.clone() if isinstance(ar, NcclAllReduce) else y)
        gg.replay(); st.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            gg.replay()
        e1.record(st); e1.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * chain)
    t = torch.tensor([us], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gg = None
    return round(t.item(), 2)

for n in (16 * 4096, 32 * 8192):
    tp_us, tn_us = time_allreduce(push, n), time_allreduce(nccl, n)
    if rank == 0:
        print(json.dumps({"allreduce_latency_us": {"numel": n, "bytes": 2 * n, "push": tp_us, "nccl_incl_one_copy": tn_us, "world": world,
                                                   "how": "50 chained calls per CUDA graph, 20 replays, device time, max over ranks"}}), flush=True)

# ---------------------------------------------------------------- 2. TP decoder layer: push vs NCCL
# two chained layers: with the push all-reduce both all-reduces of layer 0 and the first of layer 1 run FUSED (GEMM epilogue pushes,
# the following add+RMSNorm reduces); the last one is the stand-alone kernel
hidden, inter, heads, batch, kvlen, page = 4096, 11008, 32, 16, 300, 16
cfg = LlamaConfig(hidden_size=hidden, intermediate_size=inter, num_attention_heads=heads, num_hidden_layers=1)
lh = heads // world
outs = []
unfused = PushAllReduce(batch * hidden, dev)
unfused.fuse = False                    # same kernel pieces, but GEMM -> stand-alone all-reduce kernel -> add+RMSNorm
for ar in (PushAllReduce(batch * hidden, dev), unfused, NcclAllReduce()):
    torch.manual_seed(31337)            # the layers draw their reorder permutations from the global generator: same layers in every run
    layers = [TPLlamaDecoderLayer(cfg, 0, rank, world, allreduce=ar).to(dev).init_random(5 + i) for i in range(2)]
    kvs = []
    for i in range(2):
        torch.manual_seed(4242 + rank + 17 * i)
        pool = KvPoolInt4(1, lh, 128, capacity=batch * ((kvlen + page) // page + 1), block_len=page, device=dev)
        pool.buf.random_(0, 256); pool.param[..., 0].uniform_(0.01, 0.05); pool.param[..., 1].uniform_(0.0, 0.4)
        caches = [KvCacheInt4(pool, kvlen) for _ in range(batch)]
        for c in caches:
            c.acquire_one()
        kvs.append(BatchedKvCacheInt4(caches))
    torch.manual_seed(99)
    xin = torch.randn(batch, hidden, device=dev, dtype=torch.float16)
    y, pend = xin, None
    for i, (l, kv) in enumerate(zip(layers, kvs)):
        y, pend = l.forward_chain(y, pend, kv, last=(i == 1))
    assert pend is None
    if isinstance(ar, PushAllReduce) and rank == 0:
        print(json.dumps({"fused_allreduce_in_use": bool(layers[0].o_proj.can_push(batch))}), flush=True)
    outs.append(y.float())
    torch.cuda.synchronize()
# fused and stand-alone push all-reduce form the same rank-order FP32 sums: bit-identical layer outputs
report("tp_layers_fused_equals_unfused_push", torch.equal(outs[0], outs[1]), max_abs_diff=(outs[0] - outs[1]).abs().max().item())
# NCCL sums in another order and precision; two quantised layers amplify that (an INT4 code flips here and there)
d = (outs[0] - outs[2]).abs().max().item()
scale = outs[2].abs().max().item()
rel = ((outs[0] - outs[2]).norm() / outs[2].norm()).item()
report("tp_layers_push_vs_nccl", rel <= 2e-2 and bool(torch.isfinite(outs[0]).all()), max_abs_diff=d, out_absmax=scale, rel_l2=rel)
# all ranks must hold the same (replicated) hidden state
ref = outs[0].clone()
dist.broadcast(ref, 0)
report("tp_layer_replicated_across_ranks", torch.equal(ref, outs[0]))

dist.barrier()
torch.cuda.synchronize()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
