"""Which side of the gate/up fusion test is non-deterministic?  Replays the pytest sequence (small cases first), then evaluates the
reference path (gate GEMM, up GEMM, activate) and the fused op several times each on the big case and compares them pairwise."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
from atom_b200 import ops

def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")

def cat(ws):
    return (np.concatenate([p[0] for p in ws], 0), np.concatenate([p[1] for p in ws], 1),
            np.concatenate([p[2] for p in ws], 0), np.concatenate([p[3] for p in ws], 0))

def run_case(m, inter, k, reps):
    t = [O.make_gemm_inputs(m, inter, k, seed=3 * m + inter + k + i) for i in range(2)]
    act = [T(t[0][i]) for i in (0, 2, 4, 6)]
    ws = [(x[1], x[3], x[5], x[7]) for x in t]
    refs, gs, us, fused = [], [], [], []
    for _ in range(reps):
        g = ops.dense_layer_gemm_i4_fp16(act[0], T(ws[0][0]), act[1], T(ws[0][1]), act[2], T(ws[0][2]), act[3], T(ws[0][3]), flags=1)
        u = ops.dense_layer_gemm_i4_fp16(act[0], T(ws[1][0]), act[1], T(ws[1][1]), act[2], T(ws[1][2]), act[3], T(ws[1][3]), flags=1)
        refs.append(ops.activate_fp16_i4(g, u)); gs.append(g); us.append(u)
        b, bs, bk, bks = cat(ws)
        fused.append(ops.dense_layer_gemm_i4_gateup_act(act[0], T(b), act[1], T(bs), act[2], T(bk), act[3], T(bks)))
    torch.cuda.synchronize()
    def nd(xs, idx=None):
        return [int((xs[0][idx] != x[idx]).sum()) if idx is not None else int((xs[0] != x).sum()) for x in xs[1:]]
    detail = None
    for fi, (f, r) in enumerate(zip(fused, refs)):
        if not torch.equal(f[1], r[1]):
            bad = (f[1] != r[1]).nonzero()
            tiles = sorted(set((bad[:, 1] // 64).tolist()))
            row0, tile0 = int(bad[0, 0]), int(bad[0, 1]) // 64
            G = f[3].shape[0]
            detail = {"fused_index": fi, "bad_bytes": int(bad.shape[0]), "tiles": tiles[:10], "rows": sorted(set(bad[:, 0].tolist())),
                      "int8_mismatch": int((f[0] != r[0]).sum()), "keeper_scale_mismatch": int((f[2] != r[2]).sum()),
                      "group_scale_mismatch": int((f[3] != r[3]).sum()),
                      "scale_rows_bad": sorted(set((f[3] != r[3]).nonzero()[:, 0].tolist()))[:10],
                      "first": {"row": row0, "tile": tile0,
                                "got": f[1][row0, tile0 * 64:(tile0 + 1) * 64].cpu().numpy().astype(np.uint8).tolist(),
                                "ref": r[1][row0, tile0 * 64:(tile0 + 1) * 64].cpu().numpy().astype(np.uint8).tolist(),
                                "got_scale": [float(x) for x in f[3][tile0].float().cpu().numpy()[:8]],
                                "ref_scale": [float(x) for x in r[3][tile0].float().cpu().numpy()[:8]]}}
            break
    return {"case": [m, inter, k], "detail": detail, "gate_gemm_vs_first": nd(gs), "up_gemm_vs_first": nd(us), "ref_int4_vs_first": nd(refs, 1),
            "fused_int4_vs_first": nd(fused, 1), "fused_vs_ref_int4": [int((f[1] != r[1]).sum()) for f, r in zip(fused, refs)]}

for rnd in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    for case in [(16, 256, 512), (5, 384, 1024), (32, 512, 1024), (64, 256, 512)]:
        run_case(*case, reps=1)
    print(json.dumps(run_case(16, 11008, 4096, reps=4)), flush=True)
