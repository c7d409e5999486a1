"""-m gpu: mask slices of b200fa_flash_attn_ext2 — one mask per head and / or per batch entry, upstream ggml's
ne32 / ne33 broadcast.  The reference shares ONE mask between heads and batches (flash-llama.h:151,194), so parity here is against
the oracle's restatement of the broadcast rule only (oracle_flash_attn_ext3).  Typical use: per-sequence padding / attention
windows in a batched decode or a batched prefill."""
import numpy as np
import pytest

import oracle
from common import assert_close, synth_qkv
from gpu_common import pkg, to_dev

pytestmark = pytest.mark.gpu


def make_slices(ne33, ne32, n_q, n_kv, kind, seed=5):
    """f16 [ne33][ne32][rows padded to 32][n_kv]: per slice a different visible window (0 / -inf), optionally with noise."""
    rows = (n_q + 31) // 32 * 32
    rng = np.random.default_rng(seed)
    m = np.zeros((ne33, ne32, rows, n_kv), np.float32)
    for b in range(ne33):
        for h in range(ne32):
            length = int(rng.integers(max(1, n_kv // 3), n_kv + 1))      # this sequence's real length (right padding is masked)
            start = int(rng.integers(0, max(1, length // 4)))             # ... and a sliding-window start for some of them
            m[b, h, :, length:] = -np.inf
            if (b + h) % 2:
                m[b, h, :, :start] = -np.inf
            if kind == "causal":
                off = n_kv - n_q
                for i in range(n_q):
                    m[b, h, i, i + off + 1:] = -np.inf
                    m[b, h, i, max(0, min(start, i + off)):i + off + 1][-1:] = 0.0   # the diagonal stays visible: no empty rows
            elif kind == "noise":
                m[b, h, :n_q] += rng.uniform(-1, 1, (n_q, n_kv)).astype(np.float32)
    return m.astype(np.float16)


def run(D, n_q, n_kv, H, Hk, B, ne32, ne33, kind="window", q8=False, max_bias=0.0):
    import torch
    P = pkg()
    Q, K, V = synth_qkv(D, n_q, n_kv, H, Hk, n_batch=B)
    M = make_slices(ne33, ne32, n_q, n_kv, kind)
    if q8:
        Kq = oracle.quantize_q8_0(K.astype(np.float32)); Vq = oracle.quantize_q8_0(V.astype(np.float32))
        kview, vview, k, v = oracle.view_of(Kq, oracle.TYPE_Q8_0), oracle.view_of(Vq, oracle.TYPE_Q8_0), to_dev(Kq), to_dev(Vq)
    else:
        kview, vview, k, v = oracle.view_of(K), oracle.view_of(V), to_dev(K), to_dev(V)
    ref = oracle.flash_attn_ext(oracle.view_of(Q), kview, vview, oracle.view_of(M), 1 / np.sqrt(D), round_q_f16=True, max_bias=max_bias)
    out = P.flash_attn_ext(to_dev(Q), k, v, to_dev(M), max_bias=max_bias)
    torch.cuda.synchronize()
    assert_close(out.float().cpu().numpy(), ref, f"{P.last_dispatch()} ne32={ne32} ne33={ne33} {kind}")
    # and the slices really differ: with the first slice for everyone the result is another one
    if ne32 * ne33 > 1:
        shared = P.flash_attn_ext(to_dev(Q), k, v, to_dev(np.ascontiguousarray(M[0, 0])))
        torch.cuda.synchronize()
        assert float((shared - out).abs().max()) > 1e-3
    return P.last_dispatch()


@pytest.mark.parametrize("ne32,ne33", [(1, 5), (8, 1), (8, 5)])
@pytest.mark.parametrize("H,Hk", [(8, 8), (8, 2)])
def test_decode_per_sequence_and_per_head_masks(ne32, ne33, H, Hk):
    assert run(128, 1, 1000, H, Hk, 5, ne32, ne33) == "decode_stream"


@pytest.mark.parametrize("ne32,ne33", [(1, 3), (8, 3)])
def test_decode_q8_0_and_alibi_with_mask_slices(ne32, ne33):
    assert run(128, 1, 2048, 8, 2, 3, ne32, ne33, q8=True) == "decode_stream"        # transposed q8_0 tile: lines per row
    assert run(128, 1, 777, 8, 8, 3, ne32, ne33, kind="noise", max_bias=8.0) == "decode_stream"


@pytest.mark.parametrize("ne32,ne33", [(1, 2), (4, 2)])
def test_bursts_with_mask_slices(ne32, ne33):
    run(128, 3, 515, 4, 2, 2, ne32, ne33, kind="causal")     # 6 rows per kv head: stream kernel
    run(128, 12, 524, 4, 2, 2, ne32, ne33, kind="causal")    # 24 rows: virtual heads
    run(64, 5, 300, 4, 1, 2, ne32, ne33, kind="noise")


@pytest.mark.parametrize("ne32,ne33", [(1, 3), (4, 1), (4, 3)])
@pytest.mark.parametrize("n_q,n_kv,kind", [(300, 300, "causal"), (130, 517, "window"), (256, 640, "noise")])
def test_prefill_with_mask_slices(ne32, ne33, n_q, n_kv, kind):
    assert run(128, n_q, n_kv, 4, 2, 3, ne32, ne33, kind=kind) == "prefill_tcgen05"


def test_head_size_256_and_unaligned_slices():
    run(256, 1, 500, 4, 4, 2, 4, 2)                           # the 16-row kernel
    import torch
    P = pkg()
    # slice strides that are not multiples of 16 bytes: the staged (bulk-copy) path steps aside
    D, n_kv, H, B = 128, 333, 4, 3
    Q, K, V = synth_qkv(D, 1, n_kv, H, H, n_batch=B)
    M = make_slices(B, 1, 1, n_kv, "window")
    ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), oracle.view_of(M), 1 / np.sqrt(D), round_q_f16=True)
    out = P.flash_attn_ext(to_dev(Q), to_dev(K), to_dev(V), to_dev(M))
    torch.cuda.synchronize()
    assert_close(out.cpu().numpy(), ref, "odd n_kv, per-sequence masks")


def test_mask_slice_argument_errors():
    import torch
    P = pkg()
    q = torch.zeros((2, 4, 1, 128), device="cuda"); k = torch.zeros((2, 4, 64, 128), dtype=torch.float16, device="cuda")
    with pytest.raises(P.B200FAError):   # 3 slices for 4 heads
        P.flash_attn_ext(q, k, k, torch.zeros((1, 3, 32, 64), dtype=torch.float16, device="cuda"))
    with pytest.raises(P.B200FAError):   # 3 slices for 2 batch entries
        P.flash_attn_ext(q, k, k, torch.zeros((3, 1, 32, 64), dtype=torch.float16, device="cuda"))
