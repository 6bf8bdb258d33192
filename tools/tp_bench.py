"""Tensor-parallel decode step of one W4A4 decoder layer (BASELINE config #5: Llama-65B dims over N GPUs).
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/tp_bench.py [--hidden 8192 --inter 22016 --heads 64 --batch 32 --kvlen 1024]
Every rank holds heads/N heads and 1/N of the MLP channels; two NCCL all-reduces per layer (o_proj, down_proj).
Time is measured on the device with CUDA events, max over ranks."""
import argparse, json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atom_b200.kvcache import BatchedKvCacheInt4, KvCacheInt4, KvPoolInt4
from atom_b200.llama import LlamaConfig
from atom_b200.tp import TPLlamaDecoderLayer

ap = argparse.ArgumentParser()
ap.add_argument("--hidden", type=int, default=8192); ap.add_argument("--inter", type=int, default=22016)
ap.add_argument("--heads", type=int, default=64); ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--kvlen", type=int, default=1024); ap.add_argument("--layers", type=int, default=80)
ap.add_argument("--page", type=int, default=32); ap.add_argument("--iters", type=int, default=30)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = LlamaConfig(hidden_size=a.hidden, intermediate_size=a.inter, num_attention_heads=a.heads, num_hidden_layers=1)
copies = 2
layers = [TPLlamaDecoderLayer(cfg, 0, rank, world).to(dev).init_random(i) for i in range(copies)]
lh = a.heads // world
kvs = []
for i in range(copies):
    pool = KvPoolInt4(1, lh, 128, capacity=a.batch * ((a.kvlen + a.page) // a.page + 1), block_len=a.page, device=dev)
    pool.buf.random_(0, 256); pool.param[..., 0].uniform_(0.01, 0.05); pool.param[..., 1].uniform_(0.0, 0.4)
    caches = [KvCacheInt4(pool, a.kvlen) for _ in range(a.batch)]
    for c in caches:
        c.acquire_one()
    kvs.append(BatchedKvCacheInt4(caches))
x = torch.randn(a.batch, a.hidden, device=dev, dtype=torch.float16)

def step():
    y = x
    for l, kv in zip(layers, kvs):
        y = l(y, kv)
    return y

for _ in range(5):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
# one CUDA graph per step (NCCL all-reduces are capturable); falls back to eager launches if capture is refused
graph, launch = None, "eager (python launch overhead included)"
try:
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        step(); st.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            step()
        graph.replay(); st.synchronize()
    launch = "one CUDA graph per step (NCCL all-reduce captured)"
except Exception as e:  # noqa: BLE001
    graph = None
    if rank == 0:
        print("graph capture failed:", str(e)[:200], file=sys.stderr)
    torch.cuda.synchronize()
ts = []
run_stream = st if graph is not None else torch.cuda.current_stream()
with torch.cuda.stream(run_stream):
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(run_stream)
        if graph is not None:
            graph.replay()
        else:
            step()
        e1.record(run_stream); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / copies)
ts.sort()
t = torch.tensor([ts[len(ts) // 2]], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    us = t.item()
    print(json.dumps({"tp_decode_layer": {"hidden": a.hidden, "inter": a.inter, "heads": a.heads, "batch": a.batch, "kv_len": a.kvlen, "tp": world},
                      "us_per_layer": round(us, 1), "tokens_per_s_at_layers": {str(a.layers): round(a.batch / (us * a.layers * 1e-6), 1)},
                      "launch": launch, "allreduces_per_layer": 2 if world > 1 else 0}))
# a captured NCCL graph must be gone before the communicator is torn down (otherwise destroy can block forever)
graph = None
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
sys.stdout.flush()
os._exit(0)
