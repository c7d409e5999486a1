"""-m gpu parity tests for the tcgen05/TMEM/TMA prefill path, through the C ABI, against the CPU oracle
and (small case) the reference's own flash_attn_ext_f16 CUDA kernel."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from parity_log import three_way
from common import assert_close, make_mask, synth_qkv
from gpu_common import pkg, run_both, to_dev

pytestmark = pytest.mark.gpu


def _check_dispatch():
    assert pkg().last_dispatch() == "prefill_tcgen05", pkg().last_dispatch()


@pytest.mark.parametrize("n_q,n_kv,H,Hk", [(128, 128, 1, 1), (128, 256, 2, 2), (256, 256, 4, 4), (256, 512, 8, 2),
                                           (384, 384, 4, 1)])
def test_no_mask(n_q, n_kv, H, Hk):
    Q, K, V = synth_qkv(128, n_q, n_kv, H, Hk)
    run_both(Q, K, V, None)
    _check_dispatch()


@pytest.mark.parametrize("n_q,n_kv,H,Hk", [(128, 128, 2, 2), (256, 256, 4, 4), (512, 512, 8, 2), (256, 640, 4, 2)])
def test_causal_flag_mask_tensor_and_both(n_q, n_kv, H, Hk):
    Q, K, V = synth_qkv(128, n_q, n_kv, H, Hk)
    mask = make_mask("causal", n_q, n_kv)
    a, _ = run_both(Q, K, V, mask)                                   # mask tensor only -> tile classification pre-pass
    _check_dispatch()
    assert pkg().last_launch_count() == 2                            # one helper launch (q->f16 + mask scan), attention
    b, _ = run_both(Q, K, V, mask, flags=pkg().FLAG_CAUSAL)          # both
    c, _ = run_both(Q, K, V, mask, flags=pkg().FLAG_CAUSAL, drop_mask_for_product=True)  # flag only
    assert np.abs(a - b).max() < 1e-6 and np.abs(b - c).max() < 1e-6


@pytest.mark.parametrize("n_q,n_kv", [(65, 128), (129, 200), (200, 129), (300, 1000), (127, 127), (130, 70)])
def test_ragged_sizes(n_q, n_kv):
    H, Hk = 4, 2
    Q, K, V = synth_qkv(128, n_q, n_kv, H, Hk)
    run_both(Q, K, V, make_mask("causal", n_q, n_kv) if n_kv >= n_q else None)
    _check_dispatch()
    run_both(Q, K, V, make_mask("noise", n_q, n_kv))
    _check_dispatch()


def test_noise_mask_unaligned_rows():
    n_q, n_kv = 128, 250  # nb31 = 500 bytes: scalar mask loads
    Q, K, V = synth_qkv(128, n_q, n_kv, 2, 2)
    run_both(Q, K, V, make_mask("noise", n_q, n_kv))
    _check_dispatch()


def test_sliding_window_mask_skips_tiles():
    n_q = n_kv = 768
    Q, K, V = synth_qkv(128, n_q, n_kv, 2, 1)
    m = make_mask("causal", n_q, n_kv)
    for i in range(n_q):
        m[i, :max(0, i - 200)] = -np.inf  # window of 200 keys: leading tiles fully masked
    run_both(Q, K, V, m)
    _check_dispatch()


def test_fully_masked_rows_zero():
    n_q = n_kv = 256
    Q, K, V = synth_qkv(128, n_q, n_kv, 2, 2)
    m = np.zeros((n_q, n_kv), np.float16); m[128:, :] = -np.inf; m[5, :] = -np.inf
    got, _ = run_both(Q, K, V, m)
    assert np.all(got[0, 128:] == 0) and np.all(got[0, 5] == 0)


def test_f16_q_f16_dst_cache_view_batch():
    Q, K, V = synth_qkv(128, 256, 256, 8, 2, n_batch=2)
    run_both(Q, K, V, make_mask("causal", 256, 256), q_f16=True, dst_f16=True, cache_view=True, flags=pkg().FLAG_CAUSAL)
    _check_dispatch()
    run_both(Q, K, V, None, cache_view=True)
    _check_dispatch()


def test_large_scores_trigger_rescale():
    """Row maxima that keep growing across KV tiles force the lazy O rescale (threshold 2^8)."""
    n_q, n_kv = 128, 1024
    Q, K, V = synth_qkv(128, n_q, n_kv, 1, 1)
    ramp = np.linspace(0.2, 6.0, n_kv).astype(np.float32)[None, None, :, None]
    K2 = (K.astype(np.float32) * ramp).astype(np.float16)
    run_both(Q, K2, V, None, scale=1.0)
    _check_dispatch()


def test_matches_rows16_path():
    Q, K, V = synth_qkv(128, 256, 384, 4, 2)
    mask = make_mask("causal", 256, 384)
    a, _ = run_both(Q, K, V, mask, flags=pkg().FLAG_CAUSAL)
    _check_dispatch()
    b, _ = run_both(Q, K, V, mask, flags=pkg().FLAG_CAUSAL | pkg().FLAG_NO_TCGEN05)
    assert pkg().last_dispatch() == "rows16_mma"
    assert np.abs(a - b).max() < 1e-3


@pytest.mark.parametrize("seed", [0])
def test_c3_llama7b_prefill_2048_causal_sampled(seed):
    """BASELINE.json configs[2] at full size; the oracle checks 4 heads x sampled rows (it is O(n^2) on CPU)."""
    import torch
    P = pkg()
    n = 2048; H = 32
    Q, K, V = synth_qkv(128, n, n, H, H)
    mask = make_mask("causal", n, n)
    q, k, v, m = to_dev(Q.astype(np.float16)), to_dev(K), to_dev(V), to_dev(mask)
    out = P.flash_attn_ext(q, k, v, m, flags=P.FLAG_CAUSAL)
    torch.cuda.synchronize()
    _check_dispatch()
    out2 = P.flash_attn_ext(q, k, v, m)  # mask tensor only
    torch.cuda.synchronize()
    got = out.cpu().numpy(); got2 = out2.cpu().numpy()
    assert np.abs(got - got2).max() < 1e-6
    rows = np.array([0, 1, 127, 128, 129, 1000, 1023, 1024, 2046, 2047])
    for h in (0, 13, 31):
        Qs = np.ascontiguousarray(Q[:, h:h + 1, rows].astype(np.float16))
        ms = np.ascontiguousarray(mask[rows])
        ref = oracle.flash_attn_ext(oracle.view_of(Qs), oracle.view_of(np.ascontiguousarray(K[:, h:h + 1])),
                                    oracle.view_of(np.ascontiguousarray(V[:, h:h + 1])), oracle.view_of(ms), 1 / np.sqrt(128))
        assert_close(got[0, rows, h], ref[0, :, 0], f"c3 head {h}")
    # size-independent property: a causal row i equals full attention over the first i+1 keys; row 0 == V[0]
    np.testing.assert_allclose(got[0, 0], V[0, :, 0].astype(np.float32), atol=2e-3)


def test_vs_reference_cuda_flash_attn_ext_f16_prefill():
    """The reference's own flash_attn_ext_f16<128,16,128> on a small non-causal prefill (zero mask: its -inf block
    skip has a divergent barrier, flash-llama.h:276-280, so causal masks are not a safe input for it)."""
    import torch
    path = oracle.ref_gpu_path()
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_gpu.so not built")
    lib = C.CDLL(path)
    n_q, n_kv, H, Hk, D = 128, 256, 4, 2, 128
    Q, K, V = synth_qkv(D, n_q, n_kv, H, Hk)
    mask = make_mask("noise", n_q, n_kv)
    ours, ref32 = run_both(Q, K, V, mask)
    _check_dispatch()
    q = to_dev(Q[0]); k = to_dev(K[0]); v = to_dev(V[0]); m = to_dev(mask)
    dst = torch.zeros(n_q, H, D, dtype=torch.float32, device="cuda")
    rc = lib.ref_gpu_flash_attn_ext_f16(
        C.c_void_p(q.data_ptr()), C.c_void_p(k.data_ptr()), C.c_void_p(v.data_ptr()), C.c_void_p(m.data_ptr()),
        C.c_void_p(dst.data_ptr()), C.c_float(1 / np.sqrt(D)), D, n_q, H, 1, D, n_kv, Hk, 1, n_q, n_kv * 2,
        D * 4, D * 4 * n_q, D * 4 * n_q * H, D * 2, D * n_kv * 2, D * n_kv * Hk * 2, D, H, n_q, 1, None)
    torch.cuda.synchronize()
    assert rc == 0
    refgpu = dst.cpu().numpy()
    e, ref_ok = three_way("flash_attn_ext_f16 prefill 128x256 (4q/2kv, noise mask)", ours[0], refgpu, ref32[0],
                          note="flash-llama.h:5-438 launched as flash-matrix.cu:198-206")
    if not ref_ok:
        pytest.skip(f"reference kernel outside tolerance on this input: {e}")
    assert_close(ours[0], refgpu, "ours vs reference CUDA prefill", atol=4e-3, rtol=2e-2)


# ---- persistent scheduler: more work items than SMs, uneven item costs, odd tile counts, batches ----
@pytest.mark.parametrize("n_q,n_kv,H,Hk,B,causal", [(384, 384, 40, 8, 5, True),     # 2 pairs x 40 heads x 5 = 400 items > 148 SMs, odd tile count
                                                      (640, 1152, 16, 16, 2, True),    # n_kv > n_q (chunked prefill), causal offset 512
                                                      (130, 257, 12, 4, 13, False)])   # ragged everything, 156 items
def test_persistent_many_items(n_q, n_kv, H, Hk, B, causal):
    Q, K, V = synth_qkv(128, n_q, n_kv, H, Hk, n_batch=B)
    if causal:
        run_both(Q, K, V, make_mask("causal", n_q, n_kv), flags=pkg().FLAG_CAUSAL, drop_mask_for_product=True)
    else:
        run_both(Q, K, V, None)
    _check_dispatch()


def test_both_prefill_variants_agree(monkeypatch):
    """The persistent kernel and the one-CTA-per-item kernel run the same pipeline: results agree to the last bit
    of the f16 P rounding pattern (same tile order per row)."""
    import subprocess, sys, json
    code = r'''
import sys, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from common import synth_qkv
from gpu_common import pkg, to_dev
P = pkg()
Q, K, V = synth_qkv(128, 512, 512, 6, 2)
out = P.flash_attn_ext(to_dev(Q), to_dev(K), to_dev(V), None, flags=P.FLAG_CAUSAL)
torch.cuda.synchronize()
np.save(sys.argv[1], out.cpu().numpy())
'''
    import tempfile
    outs = []
    for variant in ("persistent", "cta"):
        with tempfile.NamedTemporaryFile(suffix=".npy") as f:
            env = dict(os.environ); env["B200FA_PREFILL"] = variant
            subprocess.run([sys.executable, "-c", code, f.name], check=True, env=env, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            outs.append(np.load(f.name))
    assert np.abs(outs[0] - outs[1]).max() < 1e-6


def test_full_size_properties_c3():
    """BASELINE.json configs[2] at full size, size-independent properties: (i) linearity in V, (ii) a causal row is a
    convex combination of the first i+1 value rows, (iii) rows of a non-causal call over a KV prefix equal the causal
    rows that see exactly that prefix."""
    import torch
    P = pkg()
    n, H = 2048, 32
    Q, K, V = synth_qkv(128, n, n, H, H)
    q, k, v = to_dev(Q.astype(np.float16)), to_dev(K), to_dev(V)
    a = P.flash_attn_ext(q, k, v, None, flags=P.FLAG_CAUSAL).cpu().numpy()
    b = P.flash_attn_ext(q, k, to_dev((2 * V.astype(np.float32)).astype(np.float16)), None, flags=P.FLAG_CAUSAL).cpu().numpy()
    torch.cuda.synchronize()
    assert np.abs(b - 2 * a).max() < 4e-3                                  # (i)
    vmax = np.abs(V.astype(np.float32)).max()
    assert np.abs(a).max() <= vmax + 1e-3                                   # (ii)
    pre = 1024
    c = P.flash_attn_ext(q[:, :, pre - 1:pre], k[:, :, :pre], v[:, :, :pre], None).cpu().numpy()   # decode-style call: row pre-1 over keys [0, pre)
    assert np.abs(c[0, 0] - a[0, pre - 1]).max() < 2e-3                     # (iii)


# ---- split-KV prefill: fewer work items than SMs and a long KV range (chunked prefill / long-context continuation) ----
@pytest.mark.parametrize("n_q,n_kv,H,Hk,B,kind", [
    (256, 4096, 4, 4, 1, "none"), (256, 4096, 4, 2, 1, "causal"), (130, 3000, 3, 1, 2, "noise"), (96, 8192, 2, 2, 1, "causal"),
    (384, 2048, 8, 8, 1, "causal"), (128, 2100, 1, 1, 1, "zeros")])
def test_split_kv_prefill(n_q, n_kv, H, Hk, B, kind):
    Q, K, V = synth_qkv(128, n_q, n_kv, H, Hk, n_batch=B)
    mask = make_mask(kind, n_q, n_kv)
    flags = pkg().FLAG_CAUSAL if kind == "causal" else 0
    a, _ = run_both(Q, K, V, mask, flags=flags)
    _check_dispatch()
    n_launch = pkg().last_launch_count()
    expect = 1 + 1 + 1   # helper launch (Q conversion, and the mask scan when there is no flag) + attention + combine
    assert n_launch == expect, f"expected {expect} launches (split-KV), got {n_launch}"


def test_split_kv_prefill_f16_io_and_mask_tensor_only():
    Q, K, V = synth_qkv(128, 200, 5000, 4, 4)
    mask = make_mask("causal", 200, 5000)
    run_both(Q, K, V, mask, q_f16=True, dst_f16=True, cache_view=True)       # mask tensor without the flag: classify + attention + combine
    _check_dispatch()
    assert pkg().last_launch_count() == 3
    run_both(Q, K, V, mask, flags=pkg().FLAG_CAUSAL, q_f16=True)
    assert pkg().last_launch_count() == 2


# ---- q8_0 K/V on the prefill path: dequantised once to f16 in the workspace, then the tensor-core kernel ----
@pytest.mark.parametrize("D,n_q,n_kv,H,Hk,B,kind", [
    (128, 256, 256, 4, 4, 1, "causal"), (128, 300, 517, 4, 2, 2, "noise"), (64, 200, 264, 4, 4, 1, "causal"),
    (128, 130, 130, 3, 1, 1, "none"), (128, 256, 4096, 4, 4, 1, "causal")])
def test_q8_0_prefill(D, n_q, n_kv, H, Hk, B, kind):
    Q, K, V = synth_qkv(D, n_q, n_kv, H, Hk, n_batch=B)
    mask = make_mask(kind, n_q, n_kv)
    flags = pkg().FLAG_CAUSAL if kind == "causal" else 0
    run_both(Q, K, V, mask, flags=flags, q8=True)
    _check_dispatch()
