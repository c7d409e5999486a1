"""Lane-level model of the transposed q8_0 decode tile (csrc/decode_stream.cuh, `T8` path), on the CPU.

The kernel computes, per warp and 16-key tile,  S^T = K Q^T  and  O^T += V^T P'^T  with mma.sync.m16n8k16: the q8_0
rows are the A operands (so N = 8 covers every live query row and no MMA quad is padding), the probabilities go
through movmatrix to become the B operand, block scales are applied in fp32 to per-block partial sums (K) or folded
into the probabilities per 32-dim block (V), and the int8 -> f16 conversion leaves its +1152 bias in the operand: the
bias is cancelled by initialising the QK accumulator with -1152 * sum(Q) of the block (V subtracts it explicitly).

This file restates the per-lane register algebra with numpy (fragment layouts of the PTX ISA) and checks it against
a direct evaluation, so that every index map in the kernel is pinned without a GPU.  The kernel source cites this file.
"""
import numpy as np
import pytest

LANES = np.arange(32)
G, T = LANES >> 2, LANES & 3


def mma_m16n8k16(a, b, c):
    """a[lane][4][2], b[lane][2][2], c[lane][4] -> d[lane][4]; PTX m16n8k16 .row.col fragment layouts."""
    A = np.zeros((16, 16)); B = np.zeros((16, 8)); C = np.zeros((16, 8))
    for l in range(32):
        g, t = l >> 2, l & 3
        for r, (row, col) in enumerate([(g, 2 * t), (g + 8, 2 * t), (g, 2 * t + 8), (g + 8, 2 * t + 8)]):
            A[row, col], A[row, col + 1] = a[l][r]
        for r, k in enumerate([2 * t, 2 * t + 8]):
            B[k, g], B[k + 1, g] = b[l][r]
        C[g, 2 * t], C[g, 2 * t + 1], C[g + 8, 2 * t], C[g + 8, 2 * t + 1] = c[l]
    Dm = A @ B + C
    d = np.zeros((32, 4))
    for l in range(32):
        g, t = l >> 2, l & 3
        d[l] = Dm[g, 2 * t], Dm[g, 2 * t + 1], Dm[g + 8, 2 * t], Dm[g + 8, 2 * t + 1]
    return d


def movmatrix_trans(x):
    """x[lane][2] (one b16x2 register per lane = M[lane/4][2*(lane%4) .. +1]) -> the same layout of M^T."""
    M = np.zeros((8, 8))
    for l in range(32):
        M[l >> 2, 2 * (l & 3)], M[l >> 2, 2 * (l & 3) + 1] = x[l]
    out = np.zeros((32, 2))
    for l in range(32):
        out[l] = M.T[l >> 2, 2 * (l & 3)], M.T[l >> 2, 2 * (l & 3) + 1]
    return out


def key_of_slot(s):
    """accumulator / contraction slot s = 8h + 2t + j of the tile -> key 4t + 2h + j (kernel: keyA, keyB = keyA + 2)"""
    h, r = s >> 3, s & 7
    return 4 * (r >> 1) + 2 * h + (r & 1)


def dim_of(mt, m):
    """O^T tile mt, fragment row m -> head dim (kernel: fold / record mapping)"""
    return 32 * (mt >> 1) + 4 * (m & 7) + 2 * (mt & 1) + (m >> 3)


@pytest.mark.parametrize("D,rows", [(128, 4), (128, 8), (128, 1), (64, 4), (64, 7)])
def test_transposed_q8_tile_matches_direct(D, rows):
    rng = np.random.default_rng(D + rows)
    NB = D // 32
    Kq = rng.integers(-128, 128, (16, D)).astype(np.float64); Vq = rng.integers(-128, 128, (16, D)).astype(np.float64)
    dK = rng.uniform(0.001, 0.02, (16, NB)); dV = rng.uniform(0.001, 0.02, (16, NB))
    Q = np.zeros((8, D)); Q[:rows] = rng.uniform(-1, 1, (rows, D))
    scale = 1.0 / np.sqrt(D)
    # ---- direct ----
    Kd = Kq * np.repeat(dK, 32, axis=1); Vd = Vq * np.repeat(dV, 32, axis=1)
    S = (Q @ Kd.T) * scale  # [row][key]
    Pm = np.exp2(S - S.max(axis=1, keepdims=True))
    O_ref = Pm @ Vd

    # ---- per-lane model ----
    keyA = np.array([key_of_slot(g) for g in G]); keyB = np.array([key_of_slot(g + 8) for g in G])
    assert (keyB == keyA + 2).all() and (keyA == 4 * (G >> 1) + (G & 1)).all()
    # Q B-fragments: lane (g, t) holds Q[row g][32b + 8t .. +7]; qsn = -1152 * block sums of rows 2t, 2t+1
    s_acc = np.zeros((32, 2, 2))  # [lane][kk][rr]
    for b in range(NB):
        c = np.zeros((32, 4))
        for l in range(32):
            for rr in range(2):
                c[l][rr] = c[l][2 + rr] = -1152.0 * Q[2 * T[l] + rr, 32 * b:32 * b + 32].sum()
        for i in range(2):  # two MMAs per block: payload bytes 8t + 4i .. + 3 of rows keyA / keyB (with the +1152 bias)
            a = np.zeros((32, 4, 2)); bq = np.zeros((32, 2, 2))
            for l in range(32):
                o = 32 * b + 8 * T[l] + 4 * i
                a[l][0] = Kq[keyA[l], o:o + 2] + 1152; a[l][1] = Kq[keyB[l], o:o + 2] + 1152
                a[l][2] = Kq[keyA[l], o + 2:o + 4] + 1152; a[l][3] = Kq[keyB[l], o + 2:o + 4] + 1152
                bq[l][0] = Q[G[l], o:o + 2]; bq[l][1] = Q[G[l], o + 2:o + 4]
            c = mma_m16n8k16(a, bq, c)
        for l in range(32):
            for rr in range(2):
                s_acc[l][0][rr] += c[l][rr] * dK[keyA[l], b]
                s_acc[l][1][rr] += c[l][2 + rr] * dK[keyB[l], b]
    # scores of lane: keys keyA / keyB, rows 2t + rr
    for l in range(32):
        for rr in range(2):
            r = 2 * T[l] + rr
            assert np.allclose(s_acc[l][0][rr] * scale, S[r, keyA[l]], atol=1e-9)
            assert np.allclose(s_acc[l][1][rr] * scale, S[r, keyB[l]], atol=1e-9)
    x = s_acc * scale
    m = np.zeros((32, 2))
    for l in range(32):
        for rr in range(2):
            m[l][rr] = S[2 * T[l] + rr].max()  # warp-wide row max (kernel: xor 4, 8, 16 butterflies)
    p = np.exp2(x - m[:, None, :])
    # P'_b = p * dV, packed (rows 2t, 2t+1) per key slot, transposed by movmatrix into the B fragments
    oT = np.zeros((2 * NB, 32, 4))
    for b in range(NB):
        pk0 = np.zeros((32, 2)); pk1 = np.zeros((32, 2))
        for l in range(32):
            pk0[l] = p[l][0] * dV[keyA[l], b]; pk1[l] = p[l][1] * dV[keyB[l], b]
        bf = np.stack([movmatrix_trans(pk0), movmatrix_trans(pk1)], axis=1)  # [lane][2][2]
        for half in range(2):
            mt = 2 * b + half
            a = np.zeros((32, 4, 2))
            for l in range(32):
                g, t = G[l], T[l]
                col = 32 * b + 4 * g + 2 * half  # payload bytes 4g + 2*half (m = g) and + 1 (m = g + 8) of keys 4t .. 4t+3
                a[l][0] = Vq[4 * t:4 * t + 2, col]; a[l][1] = Vq[4 * t:4 * t + 2, col + 1]
                a[l][2] = Vq[4 * t + 2:4 * t + 4, col]; a[l][3] = Vq[4 * t + 2:4 * t + 4, col + 1]
            oT[mt] = mma_m16n8k16(a, bf, oT[mt])
    # unpack: c0/c1 = (dim(mt, g), rows 2t / 2t+1), c2/c3 = (dim(mt, g + 8), ...)
    O = np.zeros((8, D))
    for mt in range(2 * NB):
        for l in range(32):
            g, t = G[l], T[l]
            for e in range(4):
                O[2 * t + (e & 1), dim_of(mt, g + 8 * (e >> 1))] = oT[mt][l][e]
    assert np.allclose(O[:rows], O_ref[:rows], rtol=1e-9, atol=1e-12)
    dims = sorted(dim_of(mt, mm) for mt in range(2 * NB) for mm in range(16))
    assert dims == list(range(D))
