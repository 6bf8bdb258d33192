"""CPU suite: a thread-by-thread simulation of the experimental FP16-path GEMM's data movement
(atom_b200/csrc/gemm_f16path_sm100.cuh) -- the TMA boxes, the scale loader's slot layout, every converter thread's source /
destination / scale addresses (transcribed formula by formula from the kernel), the stage -> unit mapping, and the way
tcgen05.mma reads a K-major SWIZZLE_128B tile through a descriptor advanced by 32 B per K=16 step (the layout the validated INT8
kernel uses).  The simulated tile must equal the numpy model of the kernel's numerics.  What this cannot cover is the
hardware itself (barrier protocol, instruction descriptor): that is what the opt-in GPU test is for."""
import numpy as np
import pytest

from oracle import oracle as O

BM = 128


def sw128_chunk_offset(r, c):                      # w4_f16_convert.cuh
    return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)


def nib8_to_f16(w, s):                             # element order [e0 e4 e1 e5 e2 e6 e3 e7], each fp16(e * s)
    out = np.empty(8, np.float16)
    for q in range(4):
        for hi in (0, 1):
            e = (int(w) >> (4 * (q + 4 * hi))) & 0xF
            e = e - 16 if e >= 8 else e
            out[2 * q + hi] = np.float16(np.float32(e) * np.float32(s))
    return out


def i8x4_to_f16(w, s):
    b = np.frombuffer(np.uint32(w).tobytes(), np.int8)
    return (b.astype(np.float32) * np.float32(s)).astype(np.float16)


def simulate_tile(t, m, n, k, m0, n0, BN, conv_threads):
    a, b, a_scale, b_scale, ak, bk, aks, bks = t
    G = k // 128 - 1
    ldm = O.scale_size(m)
    a_scale = np.asarray(a_scale).reshape(G, ldm)
    b_scale = np.asarray(b_scale).reshape(-1)[: G * n].reshape(G, n)
    nstages = 2 * G + 2
    acc = np.zeros((BM, BN), np.float64)
    f16 = np.float16

    def tma_box(src, row0, rows, col0):            # box 64 B x rows, zero fill out of bounds
        img = np.zeros((rows, 64), np.uint8)
        r_hi = min(src.shape[0], row0 + rows)
        if r_hi > row0:
            img[: r_hi - row0] = src[row0:r_hi, col0:col0 + 64].view(np.uint8)
        return img.reshape(-1)

    for st in range(nstages):
        u = (st >> 1) if st < 2 * G else G + (st - 2 * G)
        is_i4 = st < 2 * G
        half_sel = (st & 1) if is_i4 else 0
        sg = u if is_i4 else G
        # ---- what the producer's TMA and the scale loader put into shared memory
        if u < G:
            pa, pb = tma_box(a, m0, BM, u * 64), tma_box(b, n0, BN, u * 64)
        else:
            pa, pb = tma_box(ak, m0, BM, (u - G) * 64), tma_box(bk, n0, BN, (u - G) * 64)
        as_row = np.asarray(aks) if sg == G else a_scale[sg]
        bs_row = np.asarray(bks) if sg == G else b_scale[sg]
        slot_a = np.full(128, np.nan, f16)          # 64 (lower, upper) words; NaN = never written (rows >= M)
        for w in range(BM // 2):
            blk, i = w >> 3, w & 7
            if m0 + 16 * blk + i < m:
                src = 64 * (m0 // 16 + blk) + 8 * i
                slot_a[2 * w:2 * w + 2] = as_row[src:src + 2]
        slot_b = np.full(BN, np.nan, f16)
        for c in range(BN // 8):
            if n0 + 8 * c < n:
                slot_b[8 * c:8 * c + 8] = bs_row[n0 + 8 * c:n0 + 8 * c + 8]
        # ---- converter threads
        ea = np.zeros(BM * 128, np.uint8)
        eb = np.zeros(BN * 128, np.uint8)
        for t_id in range(conv_threads):
            if is_i4:
                RP = conv_threads // 8
                j, r0 = t_id & 7, t_id >> 3
                src_off, dst_off = r0 * 64 + half_sel * 32 + j * 4, sw128_chunk_offset(r0, j)
                for kk in range(BM // RP):
                    w = pa[src_off + kk * RP * 64: src_off + kk * RP * 64 + 4].view(np.uint32)[0]
                    pw = slot_a[2 * (((r0 >> 4) + kk * (RP // 16)) * 8 + (r0 & 7)):][:2]
                    s = pw[1] if (r0 & 8) else pw[0]
                    o = dst_off + kk * RP * 128
                    ea[o:o + 16] = nib8_to_f16(w, s).view(np.uint8)
                for kk in range(BN // RP):
                    w = pb[src_off + kk * RP * 64: src_off + kk * RP * 64 + 4].view(np.uint32)[0]
                    s = f16(np.float32(slot_b[r0 + kk * RP]) * np.float32(256))
                    o = dst_off + kk * RP * 128
                    eb[o:o + 16] = nib8_to_f16(w, s).view(np.uint8)
            else:
                RP = conv_threads // 16
                j, r0 = t_id & 15, t_id >> 4
                src_off, dst_off = r0 * 64 + j * 4, sw128_chunk_offset(r0, j >> 1) + (j & 1) * 8
                for kk in range(BM // RP):
                    w = pa[src_off + kk * RP * 64: src_off + kk * RP * 64 + 4].view(np.uint32)[0]
                    pw = slot_a[2 * (((r0 >> 4) + kk * (RP // 16)) * 8 + (r0 & 7)):][:2]
                    s = pw[1] if (r0 & 8) else pw[0]
                    o = dst_off + kk * RP * 128
                    ea[o:o + 8] = i8x4_to_f16(w, s).view(np.uint8)
                for kk in range(BN // RP):
                    w = pb[src_off + kk * RP * 64: src_off + kk * RP * 64 + 4].view(np.uint32)[0]
                    s = f16(np.float32(slot_b[r0 + kk * RP]) * np.float32(256))
                    o = dst_off + kk * RP * 128
                    eb[o:o + 8] = i8x4_to_f16(w, s).view(np.uint8)
        # ---- the MMA's view: K-major SW128, logical (row, byte column c) lives at chunk (c/16) ^ (row%8)
        def logical(img, rows):
            out = np.empty((rows, 64), np.float16)
            for r in range(rows):
                row = np.concatenate([img[sw128_chunk_offset(r, c):sw128_chunk_offset(r, c) + 16] for c in range(8)])
                out[r] = row.view(np.float16)
            return out
        A, B = logical(ea, BM).astype(np.float64), logical(eb, BN).astype(np.float64)
        with np.errstate(invalid="ignore"):
            for k16 in range(4):                    # 4 x (K = 16): descriptor start advanced by 32 B per step
                acc += A[:, 16 * k16:16 * k16 + 16] @ B[:, 16 * k16:16 * k16 + 16].T
    return (acc.astype(np.float32) * np.float32(1 / 256)).astype(np.float16)


@pytest.mark.parametrize("m,n,k,BN,m0,n0", [(40, 136, 384, 128, 0, 0), (40, 136, 384, 128, 0, 128), (150, 256, 256, 256, 128, 0)])
def test_simulated_tile_equals_numeric_model(m, n, k, BN, m0, n0):
    t = O.make_gemm_inputs(m, n, k, seed=11 + m + n0, pair_shared=True)
    model = O.gemm_i4_o16_f16path_model(*t)
    tile = simulate_tile(t, m, n, k, m0, n0, BN, conv_threads=512)
    rows, cols = min(BM, m - m0), min(BN, n - n0)
    got, want = tile[:rows, :cols], model[m0:m0 + rows, n0:n0 + cols]
    assert np.isfinite(got.astype(np.float32)).all()
    ulp = np.abs(got.view(np.int16).astype(np.int32) - want.view(np.int16).astype(np.int32))
    assert ulp.max() <= 1 and (ulp != 0).mean() < 0.01          # float64 vs float64 with a different summation split
