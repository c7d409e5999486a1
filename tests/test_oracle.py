"""Pin the CPU oracle (oracle/attn_oracle.c) before anything is allowed to trust it:
  1. the reference's own known-answer table (misc/flash-attn.cu:207-295),
  2. outputs of the reference's own host attention (utils.h compiled unmodified -> tests/golden/ref_host_cases.npz),
  3. ggml q8_0 blocks from gguf.quants (tests/golden/q8_0_gguf.npz),
  4. live against oracle/_ref/libref_host.so where it exists (this container; prebuilt on the GPU box).
CPU only — none of these touch the product.
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from common import GOLDEN, assert_close, load_kat, load_ref_host_cases, make_mask, synth_qkv


def _dense_views(Q, K, V, mask):
    """Dense per-head buffers -> ggml views.  Q [b][h][q][D], K/V [b][hk][kv][D], mask [q][kv]."""
    return oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), (oracle.view_of(mask) if mask is not None else None)


def test_f16_conversion_bit_exact_all_halves():
    bits = np.arange(65536, dtype=np.uint16)
    f = oracle.f16_bits_to_f32(bits)
    ref = bits.view(np.float16).astype(np.float32)
    np.testing.assert_array_equal(f.view(np.uint32), ref.view(np.uint32))
    finite = np.isfinite(ref)
    back = oracle.f32_to_f16_bits(ref[finite])
    np.testing.assert_array_equal(back, bits[finite])


def test_f32_to_f16_rounding_matches_numpy():
    rng = np.random.default_rng(0)
    x = np.concatenate([
        rng.standard_normal(200000).astype(np.float32) * np.float32(10.0) ** rng.integers(-9, 6, 200000).astype(np.float32),
        np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e9, -1e9, 5.96e-8, 2.98e-8, 2.9802322e-8, 6.1e-5, np.inf, -np.inf],
                 np.float32)])
    with np.errstate(over="ignore"):
        ref = x.astype(np.float16).view(np.uint16)
    np.testing.assert_array_equal(oracle.f32_to_f16_bits(x), ref)


def test_known_answer_table_from_reference():
    kat = load_kat()
    d, s, h = kat["d_head"], kat["seq_len"], kat["num_heads"]
    Q = np.array(kat["query"], np.float32).reshape(1, h, s, d)
    K = np.array(kat["key"], np.float32).reshape(1, h, s, d)
    V = np.ascontiguousarray(np.array(kat["value_T"], np.float32).reshape(1, h, d, s).transpose(0, 1, 3, 2))
    exp = np.array(kat["expected"], np.float32).reshape(h, s, d)
    out = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), None, 1.0 / np.sqrt(d))
    got = out[0].transpose(1, 0, 2)  # [q][head][d] -> [head][q][d]
    assert np.abs(got - exp).max() < 6e-5, np.abs(got - exp).max()  # table is printed to 4 decimals
    # and against fp64 numpy
    s64 = np.einsum("hqd,hkd->hqk", Q[0].astype(np.float64), K[0].astype(np.float64)) / np.sqrt(d)
    p = np.exp(s64 - s64.max(-1, keepdims=True)); p /= p.sum(-1, keepdims=True)
    ref = np.einsum("hqk,hkd->hqd", p, V[0].astype(np.float64))
    assert np.abs(got - ref).max() < 1e-6


CASES = ["c1_zero", "c1_tail", "c1_nomask", "llama_32h", "gqa_causal", "noise_mask", "d64"]


@pytest.mark.parametrize("tag", CASES)
def test_oracle_matches_reference_host_golden(tag):
    g = load_ref_host_cases()
    D, n_q, n_kv, n_head, n_head_kv, s1, s2, s3 = [int(x) for x in g[tag + "__meta"]]
    Q, K, V = synth_qkv(D, n_q, n_kv, n_head, n_head_kv, 1, (s1, s2, s3))
    np.testing.assert_allclose([Q.astype(np.float64).sum(), K.astype(np.float64).sum(), V.astype(np.float64).sum()],
                               g[tag + "__inputsum"], rtol=0, atol=1e-9)  # generator drift guard
    mask = make_mask(str(g[tag + "__mask"]), n_q, n_kv)
    out = oracle.flash_attn_ext(*_dense_views(Q, K, V, mask), 1.0 / np.sqrt(D), strict_ref=True)
    # same algorithm, same accumulation order -> equal up to expf/compiler contraction
    assert np.abs(out[0] - g[tag]).max() < 2e-6, np.abs(out[0] - g[tag]).max()


def test_oracle_matches_reference_kernel_test_path():
    g = load_ref_host_cases()
    D, n_kv, n_head, n_head_kv = 128, 512, 32, 8
    q = oracle.uniform_pm1(1, (n_head, D)); k = oracle.uniform_pm1(2, (n_head_kv, n_kv, D))
    v = oracle.uniform_pm1(3, (n_head_kv, n_kv, D)); m = oracle.uniform_pm1(4, (n_kv,))
    # kernel_test rounds everything through f16 (utils.h:10-11) and uses a 1-D mask (utils.h:14)
    Q = q.reshape(1, n_head, 1, D)
    K = k.astype(np.float16).reshape(1, n_head_kv, n_kv, D); V = v.astype(np.float16).reshape(1, n_head_kv, n_kv, D)
    mask = m.astype(np.float32).reshape(1, n_kv)
    # the 1-D mask is added in f32 by that overload; the ggml path carries it as f16 — compare on an f16-exact mask
    out = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V),
                                oracle.view_of(mask.astype(np.float16)), 1.0 / np.sqrt(D), round_q_f16=True)
    # P·V in the ktest path also rounds P through f16 (utils.h:10), which the ggml path does not -> 1e-3 band
    assert np.abs(out[0, 0] - g["ktest_512"]).max() < 1.5e-3


@pytest.mark.skipif(not oracle.ref_host_available(), reason="oracle/_ref/libref_host.so not built")
@pytest.mark.parametrize("shape", [(128, 1, 256, 4, 4), (128, 7, 300, 8, 2), (64, 3, 65, 2, 1), (128, 33, 64, 2, 2)])
def test_oracle_matches_live_reference_host(shape):
    D, n_q, n_kv, n_head, n_head_kv = shape
    Q, K, V = synth_qkv(D, n_q, n_kv, n_head, n_head_kv, 1, (11, 12, 13))
    mask = make_mask("causal", n_q, n_kv)
    lib = oracle.ref_host()
    VT = np.ascontiguousarray(V[0].transpose(0, 2, 1))
    res = np.zeros((n_q, n_head, D), np.float32); scores = np.zeros((n_head, n_q, n_kv), np.float32)
    rc = lib.ref_host_attention_llama(Q[0].ctypes.data, K[0].ctypes.data, VT.ctypes.data, mask.ctypes.data,
                                      res.ctypes.data, scores.ctypes.data, D, n_q, n_kv, n_head, n_head_kv,
                                      C.c_float(1.0 / np.sqrt(D)), 2)
    assert rc == 0
    out = oracle.flash_attn_ext(*_dense_views(Q, K, V, mask), 1.0 / np.sqrt(D), strict_ref=True)
    assert np.abs(out[0] - res).max() < 2e-6


def test_strict_ref_reproduces_reference_nan_on_masked_prefix():
    """utils.h:37-41: a row whose FIRST score is -inf comes out NaN in the reference host code.
    The default (non-strict) oracle keeps such rows finite, like the reference CUDA kernels."""
    Q, K, V = synth_qkv(128, 1, 64, 1, 1)
    mask = np.zeros((1, 64), np.float16); mask[0, :8] = -np.inf
    strict = oracle.flash_attn_ext(*_dense_views(Q, K, V, mask), 0.1, strict_ref=True)
    assert np.isnan(strict).all()
    robust = oracle.flash_attn_ext(*_dense_views(Q, K, V, mask), 0.1)
    ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K[:, :, 8:]), oracle.view_of(V[:, :, 8:]), None, 0.1)
    assert np.abs(robust - ref).max() < 1e-6


def test_strided_views_equal_dense():
    """ggml KV-cache view [kv][head][d] (flash-matrix.cu:203-204 strides) == dense per head."""
    D, n_q, n_kv, H, Hk = 128, 2, 40, 8, 4
    Q, K, V = synth_qkv(D, n_q, n_kv, H, Hk, 2)
    mask = make_mask("causal", n_q, n_kv)
    dense = oracle.flash_attn_ext(*_dense_views(Q, K, V, mask), 0.09)
    Kc = np.ascontiguousarray(K.transpose(0, 2, 1, 3)).transpose(0, 2, 1, 3)  # memory [b][kv][hk][d]
    Vc = np.ascontiguousarray(V.transpose(0, 2, 1, 3)).transpose(0, 2, 1, 3)
    Qc = np.ascontiguousarray(Q.transpose(0, 2, 1, 3)).transpose(0, 2, 1, 3)  # memory [b][q][h][d]
    assert not Kc.flags.c_contiguous
    strided = oracle.flash_attn_ext(oracle.view_of(Qc), oracle.view_of(Kc), oracle.view_of(Vc), oracle.view_of(mask), 0.09)
    np.testing.assert_array_equal(dense, strided)


def test_f16_q_and_f16_dst():
    Q, K, V = synth_qkv(128, 3, 50, 4, 2)
    Qh = Q.astype(np.float16)
    a = oracle.flash_attn_ext(oracle.view_of(Qh), oracle.view_of(K), oracle.view_of(V), None, 0.0884)
    b = oracle.flash_attn_ext(oracle.view_of(Qh.astype(np.float32)), oracle.view_of(K), oracle.view_of(V), None, 0.0884)
    np.testing.assert_array_equal(a, b)
    h = oracle.flash_attn_ext(oracle.view_of(Qh), oracle.view_of(K), oracle.view_of(V), None, 0.0884,
                              dst_type=oracle.TYPE_F16)
    assert h.dtype == np.float16
    np.testing.assert_array_equal(h, a.astype(np.float16))


# ---- q8_0: published ggml format, pinned against gguf.quants ----
def test_q8_0_quantize_matches_gguf_bytes():
    g = np.load(os.path.join(GOLDEN, "q8_0_gguf.npz"))
    got = oracle.quantize_q8_0(g["x"])
    np.testing.assert_array_equal(got, g["blocks"])


def test_q8_0_dequantize_matches_gguf_bits():
    g = np.load(os.path.join(GOLDEN, "q8_0_gguf.npz"))
    got = oracle.dequantize_q8_0(g["blocks"])
    np.testing.assert_array_equal(got.view(np.uint32), g["dequant"].view(np.uint32))


def test_q8_0_attention_equals_attention_on_dequantised_f32():
    D, n_q, n_kv, H, Hk = 128, 1, 96, 4, 2
    Q, K, V = synth_qkv(D, n_q, n_kv, H, Hk)
    Kq = oracle.quantize_q8_0(K.astype(np.float32)); Vq = oracle.quantize_q8_0(V.astype(np.float32))
    a = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(Kq, oracle.TYPE_Q8_0),
                              oracle.view_of(Vq, oracle.TYPE_Q8_0), None, 0.0884)
    Kd = oracle.dequantize_q8_0(Kq); Vd = oracle.dequantize_q8_0(Vq)
    b = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(Kd), oracle.view_of(Vd), None, 0.0884)
    np.testing.assert_array_equal(a, b)
    # and q8_0 stays close to the f16 original (sanity on the quantiser, not a parity claim)
    c = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), None, 0.0884)
    assert np.abs(a - c).max() < 5e-3


# ---- split-KV merge algebra (fa_reduce, flash_row_float.h:429-471) ----
@pytest.mark.parametrize("n_parts", [1, 2, 5, 16])
def test_merge_partials_equals_unsplit(n_parts):
    D, n_kv = 128, 64 * n_parts
    Q, K, V = synth_qkv(D, 1, n_kv, 1, 1)
    full = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), None, 0.0884)[0, 0, 0]
    ms, ls, Os = [], [], []
    q = Q[0, 0, 0].astype(np.float64)
    for p in range(n_parts):
        k = K[0, 0, p * 64:(p + 1) * 64].astype(np.float64); v = V[0, 0, p * 64:(p + 1) * 64].astype(np.float64)
        s = k @ q * 0.0884
        m = s.max(); e = np.exp(s - m)
        ms.append(m); ls.append(e.sum()); Os.append(e @ v)
    got = oracle.merge_partials(np.array(ms), np.array(ls), np.array(Os))
    assert np.abs(got - full).max() < 2e-6


def test_merge_partials_ignores_empty_parts():
    O = np.ones((3, 4), np.float32); O[1] = 7.0
    got = oracle.merge_partials(np.array([0.0, -np.inf, 0.0]), np.array([1.0, 0.0, 1.0]), O)
    np.testing.assert_allclose(got, np.ones(4), atol=1e-7)


@pytest.mark.parametrize("n_head,max_bias,softcap", [(6, 8.0, 0.0), (12, 4.0, 0.0), (8, 0.0, 30.0), (5, 2.5, 1.5)])
def test_score_modifiers_against_numpy(n_head, max_bias, softcap):
    """oracle_flash_attn_ext2 (ALiBi slopes + logit soft-cap, upstream ggml semantics — not in the reference) against an independent
    float64 numpy evaluation of the published formulae; 6, 12 and 5 heads exercise both branches of the slope rule."""
    D, n_q, n_kv = 32, 5, 40
    rs = np.random.RandomState(3)
    Q = rs.uniform(-1, 1, (1, n_head, n_q, D)).astype(np.float32)
    K = rs.uniform(-1, 1, (1, n_head, n_kv, D)).astype(np.float16)
    V = rs.uniform(-1, 1, (1, n_head, n_kv, D)).astype(np.float16)
    qi = np.arange(n_q)[:, None] + (n_kv - n_q); kj = np.arange(n_kv)[None, :]
    M = (-np.abs(qi - kj)).astype(np.float32); M[kj > qi] = -np.inf
    M = M.astype(np.float16)
    scale = 0.4
    got = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), oracle.view_of(M), scale,
                                max_bias=max_bias, logit_softcap=softcap)
    n2 = 2 ** int(np.floor(np.log2(n_head)))
    m0, m1 = 2.0 ** (-max_bias / n2), 2.0 ** (-(max_bias / 2.0) / n2)
    ref = np.zeros((1, n_q, n_head, D))
    for h in range(n_head):
        slope = 1.0 if max_bias <= 0 else (m0 ** (h + 1) if h < n2 else m1 ** (2 * (h - n2) + 1))
        s = Q[0, h].astype(np.float64) @ K[0, h].astype(np.float64).T * (scale / softcap if softcap else scale)
        if softcap:
            s = softcap * np.tanh(s)
        s = s + slope * M.astype(np.float64)
        p = np.exp(s - s.max(axis=1, keepdims=True)); p /= p.sum(axis=1, keepdims=True)
        ref[0, :, h, :] = p @ V[0, h].astype(np.float64)
    assert np.abs(got - ref).max() < 2e-6


def test_mask_slices_follow_the_ggml_broadcast_rule():
    """oracle_flash_attn_ext3: one mask per head / batch entry (upstream ggml's ne32 / ne33 broadcast; the reference's mask is shared,
    flash-llama.h:151,194).  Cross-checked against float64 numpy; a 4-D mask with ne32 = ne33 = 1 equals the shared 2-D mask."""
    D, n_q, n_kv, H, Hk, B = 32, 3, 40, 4, 2, 2
    Q, K, V = synth_qkv(D, n_q, n_kv, H, Hk, n_batch=B)
    rng = np.random.default_rng(0)
    for ne33, ne32 in [(B, H), (B, 1), (1, H), (1, 1)]:
        M = rng.uniform(-1, 1, (ne33, ne32, 32, n_kv)).astype(np.float16)
        ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), oracle.view_of(M), 1 / np.sqrt(D))
        for b in range(B):
            for h in range(H):
                s = (Q[b, h].astype(np.float64) @ K[b, h // 2].astype(np.float64).T) / np.sqrt(D) + M[b % ne33, h % ne32, :n_q].astype(np.float64)
                p = np.exp(s - s.max(1, keepdims=True)); p /= p.sum(1, keepdims=True)
                assert np.abs(p @ V[b, h // 2].astype(np.float64) - ref[b, :, h]).max() < 1e-5
        if ne32 == ne33 == 1:
            shared = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), oracle.view_of(np.ascontiguousarray(M[0, 0])), 1 / np.sqrt(D))
            assert np.array_equal(shared, ref)
