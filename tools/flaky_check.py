import os, sys, json
import numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from oracle import oracle as O
from atom_b200 import ops
def T(a, sync):
    t = torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")
    if sync: torch.cuda.synchronize()
    return t
m, inter, k = 16, 11008, 4096
t = [O.make_gemm_inputs(m, inter, k, seed=3 * m + inter + k + i) for i in range(2)]
ws = [(x[1], x[3], x[5], x[7]) for x in t]
b = np.concatenate([p[0] for p in ws], 0); bs = np.concatenate([p[1] for p in ws], 1)
bk = np.concatenate([p[2] for p in ws], 0); bks = np.concatenate([p[3] for p in ws], 0)
for sync in (0, 1):
    bad = 0
    for it in range(25):
        act = [T(t[0][i], sync) for i in (0, 2, 4, 6)]
        g = ops.dense_layer_gemm_i4_fp16(act[0], T(ws[0][0], sync), act[1], T(ws[0][1], sync), act[2], T(ws[0][2], sync), act[3], T(ws[0][3], sync), flags=1)
        u = ops.dense_layer_gemm_i4_fp16(act[0], T(ws[1][0], sync), act[1], T(ws[1][1], sync), act[2], T(ws[1][2], sync), act[3], T(ws[1][3], sync), flags=1)
        ref = ops.activate_fp16_i4(g, u)
        got = ops.dense_layer_gemm_i4_gateup_act(act[0], T(b, sync), act[1], T(bs, sync), act[2], T(bk, sync), act[3], T(bks, sync))
        torch.cuda.synchronize()
        ok = torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])
        bad += 0 if ok else 1
    print(json.dumps({"sync_after_upload": sync, "mismatches_of_25": bad, "pdl_env": os.environ.get("ATOM_B200_GEMM_PDL", "default")}), flush=True)
