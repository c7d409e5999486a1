"""-m gpu: b200fa_flash_attn_ext2 — ALiBi slopes on the mask (max_bias) and logit soft-cap, the two score modifiers of upstream
ggml's flash_attn_ext.  Not in the reference: the oracle restates ggml's published formulae
(oracle/attn_oracle.c, oracle_flash_attn_ext2), so parity here is against that restatement only."""
import numpy as np
import pytest

import oracle
from common import assert_close, synth_qkv
from gpu_common import kv_cache_view, pkg, to_dev

pytestmark = pytest.mark.gpu


def alibi_mask(n_q, n_kv, causal=True):
    """ggml builds the ALiBi bias as -|distance| in the KQ mask; the slope comes from max_bias inside the op."""
    off = n_kv - n_q
    qi = np.arange(n_q)[:, None] + off
    kj = np.arange(n_kv)[None, :]
    m = -np.abs(qi - kj).astype(np.float32)
    if causal:
        m[kj > qi] = -np.inf
    return m.astype(np.float16)


def run(D, n_q, n_kv, H, Hk, B=1, max_bias=0.0, softcap=0.0, scale=None, q8=False, causal=True, mask=True, cache_view=False, q_f16=False):
    import torch
    P = pkg()
    Q, K, V = synth_qkv(D, n_q, n_kv, H, Hk, n_batch=B)
    if q_f16:
        Q = Q.astype(np.float16)
    scale = scale if scale is not None else 1.0 / np.sqrt(D)
    M = alibi_mask(n_q, n_kv, causal) if mask else None
    if q8:
        Kq = oracle.quantize_q8_0(K.astype(np.float32)); Vq = oracle.quantize_q8_0(V.astype(np.float32))
        kview, vview = oracle.view_of(Kq, oracle.TYPE_Q8_0), oracle.view_of(Vq, oracle.TYPE_Q8_0)
        k, v = to_dev(Kq), to_dev(Vq)
    else:
        kview, vview = oracle.view_of(K), oracle.view_of(V)
        k, v = to_dev(K), to_dev(V)
    ref = oracle.flash_attn_ext(oracle.view_of(Q), kview, vview, oracle.view_of(M) if M is not None else None, scale, round_q_f16=True,
                                max_bias=max_bias, logit_softcap=softcap)
    q = to_dev(Q)
    if cache_view:
        q, k, v = kv_cache_view(q), kv_cache_view(k), kv_cache_view(v)
    m = None
    if M is not None:
        rows = (n_q + 31) // 32 * 32
        mm = np.zeros((rows, n_kv), np.float16); mm[:n_q] = M
        m = to_dev(mm)
    out = P.flash_attn_ext(q, k, v, m, scale=scale, max_bias=max_bias, logit_softcap=softcap)
    torch.cuda.synchronize()
    assert_close(out.float().cpu().numpy(), ref, f"{P.last_dispatch()} max_bias={max_bias} softcap={softcap}")
    return P.last_dispatch()


MODS = [(8.0, 0.0), (0.0, 30.0), (8.0, 50.0), (2.5, 1.5)]


@pytest.mark.parametrize("max_bias,softcap", MODS)
@pytest.mark.parametrize("H,Hk", [(6, 6), (12, 3), (8, 8)])   # 6 and 12 heads: both branches of the slope formula
def test_decode(max_bias, softcap, H, Hk):
    assert run(128, 1, 1500, H, Hk, B=2, max_bias=max_bias, softcap=softcap, scale=0.5) == "decode_stream"


@pytest.mark.parametrize("max_bias,softcap", MODS)
def test_decode_q8_0_and_d64(max_bias, softcap):
    run(128, 1, 1024, 12, 4, max_bias=max_bias, softcap=softcap, scale=0.5, q8=True)
    run(64, 1, 700, 6, 2, max_bias=max_bias, softcap=softcap, scale=0.7, cache_view=True, q_f16=True)


@pytest.mark.parametrize("max_bias,softcap", MODS)
def test_burst_rows16(max_bias, softcap):
    run(128, 24, 400, 6, 2, max_bias=max_bias, softcap=softcap, scale=0.5)       # 72 rows per KV head -> n_q < 64: rows16 kernel
    run(80, 9, 300, 5, 5, max_bias=max_bias, softcap=softcap, scale=0.5)         # padded head size


@pytest.mark.parametrize("max_bias,softcap", MODS)
@pytest.mark.parametrize("n_q,n_kv", [(256, 256), (200, 456)])
def test_prefill(max_bias, softcap, n_q, n_kv):
    assert run(128, n_q, n_kv, 6, 3, B=2, max_bias=max_bias, softcap=softcap, scale=0.5) == "prefill_tcgen05"


def test_prefill_softcap_without_mask_and_small_heads():
    assert run(96, 300, 300, 4, 4, softcap=20.0, scale=1.0, mask=False) == "prefill_tcgen05"
    assert run(128, 130, 130, 3, 1, max_bias=4.0, causal=False) == "prefill_tcgen05"


def test_zero_modifiers_take_the_plain_entry_bitwise():
    """ext = {0, 0} is exactly b200fa_flash_attn_ext."""
    import ctypes as C
    import torch
    P = pkg()
    Q, K, V = synth_qkv(128, 1, 999, 8, 2)
    q, k, v = to_dev(Q), to_dev(K), to_dev(V)
    a = P.flash_attn_ext(q, k, v, None)
    ws = P.Workspace(P.workspace_size(0, 1, 128, 1, 8, 1, 999, 2, 1))
    b = torch.empty_like(a)
    ext = P.ExtParams(0.0, 0.0)
    rc = P.lib().b200fa_flash_attn_ext2(q.data_ptr(), k.data_ptr(), v.data_ptr(), None, b.data_ptr(), 1.0 / np.sqrt(128), 0, 1, 0,
                                        128, 1, 8, 1, 128, 999, 2, 1, 0, 0, 512, 512, 512 * 8, 256, 256 * 999, 256 * 999 * 2,
                                        256, 256 * 999, 256 * 999 * 2, 128, 8, 1, 1, C.byref(ext), 0, ws.ptr, ws.nbytes, None)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(a, b)


def test_invalid_modifiers():
    import torch
    P = pkg()
    q = torch.zeros((1, 2, 1, 128), device="cuda"); k = torch.zeros((1, 2, 64, 128), device="cuda", dtype=torch.float16)
    with pytest.raises(P.B200FAError):
        P.flash_attn_ext(q, k, k, None, max_bias=-1.0)
    with pytest.raises(P.B200FAError):
        P.flash_attn_ext(q, k, k, None, logit_softcap=float("inf"))


@pytest.mark.parametrize("max_bias,softcap", [(8.0, 0.0), (0.0, 30.0), (4.0, 20.0)])
@pytest.mark.parametrize("n_q,q8", [(1, False), (1, True), (40, False)])
def test_modifiers_in_the_sequence_split_entry(max_bias, softcap, n_q, q8):
    """b200fa_flash_attn_partial2: four KV slices with ALiBi / soft-cap, each given its own mask columns, merged = the unsplit result."""
    import torch
    P = pkg()
    D, n_kv, H, Hk = 128, 1024, 8, 2
    Q, K, V = synth_qkv(D, n_q, n_kv, H, Hk)
    M = alibi_mask(n_q, n_kv, True)
    if q8:
        Kq = oracle.quantize_q8_0(K.astype(np.float32)); Vq = oracle.quantize_q8_0(V.astype(np.float32))
        kview, vview, k, v = oracle.view_of(Kq, oracle.TYPE_Q8_0), oracle.view_of(Vq, oracle.TYPE_Q8_0), to_dev(Kq), to_dev(Vq)
    else:
        kview, vview, k, v = oracle.view_of(K), oracle.view_of(V), to_dev(K), to_dev(V)
    ref = oracle.flash_attn_ext(oracle.view_of(Q), kview, vview, oracle.view_of(M), 1 / np.sqrt(D), round_q_f16=True, max_bias=max_bias, logit_softcap=softcap)
    rows = (n_q + 31) // 32 * 32
    mm = np.zeros((rows, n_kv), np.float16); mm[:n_q] = M
    q = to_dev(Q)
    parts = []
    for i in range(4):
        sl = slice(i * 256, (i + 1) * 256)
        parts.append(P.flash_attn_partial(q, k[:, :, sl], v[:, :, sl], to_dev(np.ascontiguousarray(mm[:, sl])), kv_pos0=i * 256, n_kv_total=n_kv,
                                          max_bias=max_bias, logit_softcap=softcap))
    out = P.merge_partials(torch.stack(parts))
    torch.cuda.synchronize()
    assert_close(out.cpu().numpy().reshape(ref.shape), ref, f"partial2 max_bias={max_bias} softcap={softcap} n_q={n_q} q8={q8}")
