"""-m gpu parity tests for the split-KV decode / rows16 path, through the C ABI, against the CPU oracle,
the committed reference-host goldens, and the reference's own CUDA kernels (oracle/_ref/libref_gpu.so)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from parity_log import three_way
from common import assert_close, load_ref_host_cases, make_mask, synth_qkv
from gpu_common import kv_cache_view, pkg, run_both, to_dev

pytestmark = pytest.mark.gpu


# ---- BASELINE.json configs[0]: the repo's flash-matrix test ----
@pytest.mark.parametrize("mask_kind", ["zeros", "tail56", "none"])
def test_c1_flash_matrix_case(mask_kind):
    Q, K, V = synth_qkv(128, 1, 256, 1, 1)
    run_both(Q, K, V, make_mask(mask_kind, 1, 256), mask_pad=32)
    assert pkg().last_dispatch() == "decode_stream"


@pytest.mark.parametrize("tag", ["c1_zero", "c1_tail", "c1_nomask", "llama_32h", "gqa_causal", "noise_mask", "d64"])
def test_against_reference_host_goldens(tag):
    """Product vs the outputs of the reference's own host attention (tests/golden/ref_host_cases.npz)."""
    import torch
    g = load_ref_host_cases()
    D, n_q, n_kv, n_head, n_head_kv, s1, s2, s3 = [int(x) for x in g[tag + "__meta"]]
    Q, K, V = synth_qkv(D, n_q, n_kv, n_head, n_head_kv, 1, (s1, s2, s3))
    mask = make_mask(str(g[tag + "__mask"]), n_q, n_kv)
    P = pkg()
    out = P.flash_attn_ext(to_dev(Q), to_dev(K), to_dev(V), to_dev(mask) if mask is not None else None)
    torch.cuda.synchronize()
    assert_close(out.cpu().numpy()[0], g[tag], tag)


# ---- BASELINE.json configs[1]: LLaMA-7B decode, KV-cache view strides + padded mask ----
def test_c2_llama7b_decode_kv4096_cache_view():
    Q, K, V = synth_qkv(128, 1, 4096, 32, 32)
    run_both(Q, K, V, make_mask("zeros", 1, 4096), cache_view=True, mask_pad=32)
    assert pkg().last_dispatch() == "decode_stream"
    assert pkg().last_launch_count() == 1  # splits are merged in-kernel by the last CTA of each row group


@pytest.mark.parametrize("n_kv", [1, 15, 16, 17, 77, 255, 256, 257, 1000, 5000])
def test_ragged_kv_lengths_gqa(n_kv):
    Q, K, V = synth_qkv(128, 1, n_kv, 32, 8)
    run_both(Q, K, V, make_mask("noise", 1, n_kv))


@pytest.mark.parametrize("n_q,H,Hk", [(1, 8, 1), (2, 32, 8), (5, 32, 8), (3, 16, 16), (16, 4, 4), (40, 8, 2)])
def test_several_queries_causal_mask_tensor_and_flag(n_q, H, Hk):
    n_kv = 300
    Q, K, V = synth_qkv(128, n_q, n_kv, H, Hk)
    mask = make_mask("causal", n_q, n_kv)
    a, _ = run_both(Q, K, V, mask, flags=pkg().FLAG_NO_TCGEN05)
    b, _ = run_both(Q, K, V, mask, flags=pkg().FLAG_CAUSAL | pkg().FLAG_NO_TCGEN05)
    c, _ = run_both(Q, K, V, mask, flags=pkg().FLAG_CAUSAL | pkg().FLAG_NO_TCGEN05, drop_mask_for_product=True)  # synthesised
    assert np.abs(a - b).max() < 1e-5 and np.abs(b - c).max() < 1e-5


def test_f16_q_and_f16_dst():
    Q, K, V = synth_qkv(128, 1, 700, 32, 8)
    run_both(Q, K, V, None, q_f16=True, dst_f16=True)


def test_head_dim_64():
    Q, K, V = synth_qkv(64, 2, 333, 8, 4)
    run_both(Q, K, V, make_mask("causal", 2, 333))


def test_batch_and_batch_broadcast():
    Q, K, V = synth_qkv(128, 1, 513, 8, 2, n_batch=3)
    run_both(Q, K, V, make_mask("noise", 1, 513), cache_view=True)
    # batch broadcast (ne03 = 4, ne13 = 2 -> ik3 = iq3 / 2, flash-llama.h:129,137)
    Q4 = oracle.uniform_pm1(1, (4, 8, 1, 128)); _, K2, V2 = synth_qkv(128, 1, 200, 8, 2, n_batch=2)
    run_both(Q4, K2, V2, None)


def test_fully_masked_rows_give_zeros_not_nan():
    import torch
    Q, K, V = synth_qkv(128, 2, 64, 4, 4)
    mask = np.zeros((2, 64), np.float16); mask[1, :] = -np.inf; mask[0, :10] = -np.inf
    got, ref = run_both(Q, K, V, mask)
    assert np.all(got[0, 1] == 0)


# ---- q8_0 K/V (BASELINE.json configs[4], reduced) ----
@pytest.mark.parametrize("n_kv,H,Hk", [(256, 32, 8), (1000, 32, 8), (4096, 8, 8)])
def test_q8_0_kv_in_loop_dequant(n_kv, H, Hk):
    Q, K, V = synth_qkv(128, 1, n_kv, H, Hk)
    run_both(Q, K, V, None, q8=True)


def test_q8_0_quantize_dequantize_kernels_bit_exact():
    import torch
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "q8_0_gguf.npz"))
    P = pkg()
    qb = P.quantize_q8_0(to_dev(g["x"]))
    np.testing.assert_array_equal(qb.cpu().numpy(), g["blocks"])
    dq = P.dequantize_q8_0(to_dev(g["blocks"]))
    np.testing.assert_array_equal(dq.cpu().numpy().view(np.uint32), g["dequant"].view(np.uint32))
    x = oracle.uniform_pm1(9, (64, 4096)).astype(np.float32) * 3
    np.testing.assert_array_equal(P.quantize_q8_0(to_dev(x)).cpu().numpy(), oracle.quantize_q8_0(x))
    np.testing.assert_array_equal(P.quantize_q8_0(to_dev(x.astype(np.float16))).cpu().numpy(),
                                  oracle.quantize_q8_0(x.astype(np.float16).astype(np.float32)))


# ---- sequence-split building blocks (C5 path): partial triples + merge == unsplit ----
@pytest.mark.parametrize("q8", [False, True])
@pytest.mark.parametrize("n_parts", [2, 8])
def test_partials_merge_equals_unsplit(n_parts, q8):
    import torch
    P = pkg()
    n_kv, H, Hk = 2048, 32, 8
    Q, K, V = synth_qkv(128, 1, n_kv, H, Hk)
    if q8:
        Kq, Vq = oracle.quantize_q8_0(K.astype(np.float32)), oracle.quantize_q8_0(V.astype(np.float32))
        ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(Kq, 8), oracle.view_of(Vq, 8), None, 1 / np.sqrt(128),
                                    round_q_f16=True)
        k, v = to_dev(Kq), to_dev(Vq)
    else:
        ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), None, 1 / np.sqrt(128),
                                    round_q_f16=True)
        k, v = to_dev(K), to_dev(V)
    q = to_dev(Q)
    step = n_kv // n_parts
    parts = [P.flash_attn_partial(q, k[:, :, i * step:(i + 1) * step], v[:, :, i * step:(i + 1) * step],
                                  kv_pos0=i * step, n_kv_total=n_kv) for i in range(n_parts)]
    out = P.merge_partials(torch.stack(parts))
    torch.cuda.synchronize()
    assert_close(out.cpu().numpy().reshape(ref.shape), ref, "partials+merge")
    # the triples themselves: m = row max in natural units, l = sum exp(s - m)
    p0 = parts[0].cpu().numpy()
    Kf = oracle.dequantize_q8_0(Kq) if q8 else K.astype(np.float32)
    s = (Q[0, 0, 0].astype(np.float16).astype(np.float64) @ Kf[0, 0, :step].astype(np.float64).T) / np.sqrt(128)
    assert abs(p0[0, 128] - s.max()) < 2e-3
    assert abs(p0[0, 129] - np.exp(s - s.max()).sum()) / np.exp(s - s.max()).sum() < 2e-3


def test_partial_with_causal_flag_and_positions():
    import torch
    P = pkg()
    n_q, n_kv, H, Hk = 4, 512, 8, 2
    Q, K, V = synth_qkv(128, n_q, n_kv, H, Hk)
    mask = make_mask("causal", n_q, n_kv)
    ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), oracle.view_of(mask),
                                1 / np.sqrt(128), round_q_f16=True)
    q, k, v = to_dev(Q), to_dev(K), to_dev(V)
    parts = [P.flash_attn_partial(q, k[:, :, i * 128:(i + 1) * 128], v[:, :, i * 128:(i + 1) * 128], kv_pos0=i * 128,
                                  n_kv_total=n_kv, flags=P.FLAG_CAUSAL) for i in range(4)]
    out = P.merge_partials(torch.stack(parts))
    torch.cuda.synchronize()
    assert_close(out.cpu().numpy().reshape(ref.shape), ref, "causal partials")


# ---- the reference's own CUDA kernels on the same B200 (legacy HMMA path) ----
def _ref_gpu():
    path = oracle.ref_gpu_path()
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_gpu.so not built")
    return C.CDLL(path)


@pytest.mark.parametrize("n_kv", [256, 512, 4096])
def test_vs_reference_cuda_flash_attn_row(n_kv):
    """flash_attn_row<128,8,2,256> + fa_reduce<128,8> (flash_row_float.h) vs ours, same inputs (kernel_test.h recipe)."""
    import torch
    lib = _ref_gpu()
    P = pkg()
    H, Hk, D = 32, 8, 128
    Q, K, V = synth_qkv(D, 1, n_kv, H, Hk)
    mask1d = make_mask("noise", 1, n_kv)
    ours, ref32 = run_both(Q, K, V, mask1d)
    q = to_dev(Q[0, :, 0, :]); k = to_dev(K[0]); vT = to_dev(np.ascontiguousarray(V[0].transpose(0, 2, 1)))
    m = to_dev(mask1d[0])
    lib.ref_gpu_row_tmp_halves.restype = C.c_size_t
    tmp = torch.zeros(lib.ref_gpu_row_tmp_halves(n_kv, H), dtype=torch.float16, device="cuda")
    dst = torch.zeros(H, D, dtype=torch.float32, device="cuda")
    rc = lib.ref_gpu_flash_attn_row(C.c_void_p(q.data_ptr()), C.c_void_p(k.data_ptr()), C.c_void_p(vT.data_ptr()),
                                    C.c_void_p(m.data_ptr()), C.c_void_p(tmp.data_ptr()), C.c_void_p(dst.data_ptr()),
                                    n_kv, C.c_float(1 / np.sqrt(D)), H, H // Hk, None)
    torch.cuda.synchronize()
    assert rc == 0
    refgpu = dst.cpu().numpy()
    e, ref_ok = three_way(f"flash_attn_row+fa_reduce n_kv={n_kv} (32q/8kv, noise mask)", ours[0, 0], refgpu, ref32[0, 0],
                          note="flash_row_float.h:4-200,415-472 vs b200fa vs fp32 oracle, kernel_test.h recipe")
    # ours is always held to the oracle (run_both asserted it); ours-vs-refgpu is only a meaningful gate where the f16-accumulating
    # reference kernel is itself inside the tolerance — otherwise say so loudly instead of passing having compared nothing
    if not ref_ok:
        pytest.skip(f"reference kernel outside tolerance on this input: {e}")
    assert_close(ours[0, 0], refgpu, "ours vs reference CUDA", atol=4e-3, rtol=2e-2)


def test_vs_reference_cuda_flash_attn_ext_f16_decode():
    """flash_attn_ext_f16<128,16,128> (flash-llama.h) on the kernel_test shape (kernel_test.h:191-198)."""
    import torch
    lib = _ref_gpu()
    H, Hk, D, n_kv = 32, 8, 128, 512
    Q, K, V = synth_qkv(D, 1, n_kv, H, Hk)
    mask = make_mask("noise", 1, n_kv)
    ours, ref32 = run_both(Q, K, V, mask, mask_pad=32)
    q = to_dev(Q[0, :, 0, :]); k = to_dev(K[0]); v = to_dev(V[0])
    mp = np.zeros((32, n_kv), np.float16); mp[0] = mask[0]
    m = to_dev(mp)
    dst = torch.zeros(H, D, dtype=torch.float32, device="cuda")
    rc = lib.ref_gpu_flash_attn_ext_f16(
        C.c_void_p(q.data_ptr()), C.c_void_p(k.data_ptr()), C.c_void_p(v.data_ptr()), C.c_void_p(m.data_ptr()),
        C.c_void_p(dst.data_ptr()), C.c_float(1 / np.sqrt(D)), D, 1, H, 1, D, n_kv, Hk, 1, 32, n_kv * 2,
        D * 4, D * 4, D * H * 4, D * 2, D * n_kv * 2, D * n_kv * Hk * 2, D, H, 1, 1, None)
    torch.cuda.synchronize()
    assert rc == 0
    refgpu = dst.cpu().numpy()
    e, ref_ok = three_way("flash_attn_ext_f16 decode n_kv=512 (32q/8kv, noise mask)", ours[0, 0], refgpu, ref32[0, 0],
                          note="flash-llama.h:5-438 launched as kernel_test.h:191-198")
    if not ref_ok:
        pytest.skip(f"reference kernel outside tolerance on this input: {e}")
    assert_close(ours[0, 0], refgpu, "ours vs reference CUDA ext_f16", atol=4e-3, rtol=2e-2)


def test_error_paths_on_device():
    import torch
    P = pkg()
    Q, K, V = synth_qkv(128, 1, 4096, 32, 32)
    class Tiny(P.Workspace):
        def __init__(self):
            super().__init__(256); self.nbytes = 256
    with pytest.raises(P.B200FAError) as e:
        P.flash_attn_ext(to_dev(Q), to_dev(K), to_dev(V), workspace=Tiny())
    assert e.value.status == -3


# ---- stream-K decomposition: CTAs whose chunk run crosses unit (kv head, batch) boundaries ----
@pytest.mark.parametrize("n_kv,H,Hk,B", [(200, 32, 8, 40), (64, 8, 8, 3), (130, 16, 1, 5), (4096, 16, 1, 1), (1, 4, 2, 300)])
def test_stream_units_cross_cta_boundaries(n_kv, H, Hk, B):
    Q, K, V = synth_qkv(128, 1, n_kv, H, Hk, n_batch=B)
    run_both(Q, K, V, make_mask("noise", 1, n_kv))
    assert pkg().last_dispatch() == "decode_stream"
    run_both(Q, K, V, None, dst_f16=True, q_f16=True)


@pytest.mark.parametrize("n_kv,H,Hk,B", [(200, 32, 8, 40), (1000, 8, 2, 1), (62, 4, 4, 2), (63, 4, 4, 2), (1, 2, 2, 1)])
def test_stream_q8_0_units_and_ragged_tails(n_kv, H, Hk, B):
    Q, K, V = synth_qkv(128, 1, n_kv, H, Hk, n_batch=B)
    run_both(Q, K, V, make_mask("noise", 1, n_kv), q8=True)
    # heads whose byte size is not a multiple of 16 (odd n_kv) cannot be bulk-copied: those go to the rows16 kernel
    assert pkg().last_dispatch() == ("decode_stream" if (n_kv * 136) % 16 == 0 or Hk * B == 1 else "decode_splitkv")


def test_stream_head_dim_64_q8_and_f16():
    Q, K, V = synth_qkv(64, 2, 777, 8, 2)
    run_both(Q, K, V, make_mask("causal", 2, 777))
    assert pkg().last_dispatch() == "decode_stream"
    Q, K, V = synth_qkv(64, 2, 776, 8, 2)  # 776 * 68 bytes per head: a multiple of 16, so q8_0 heads can be bulk-copied
    run_both(Q, K, V, None, q8=True)
    assert pkg().last_dispatch() == "decode_stream"


# ---- KV-cache append (+ q8_0 quantise on the fly), then decode over the cache ----
@pytest.mark.parametrize("q8", [False, True])
def test_kv_cache_append_then_decode(q8):
    import torch
    P = pkg()
    D, Hk, H, n_max, B = 128, 4, 16, 512, 2
    n_past, n_new = 300, 5
    Q, K, V = synth_qkv(D, 1, n_past + n_new, H, Hk, n_batch=B)
    Kf, Vf = K.astype(np.float32), V.astype(np.float32)
    if q8:
        kc = torch.zeros((B, Hk, n_max, D // 32 * 34), dtype=torch.uint8, device="cuda"); vc = torch.zeros_like(kc)
    else:
        kc = torch.zeros((B, n_max, Hk, D), dtype=torch.float16, device="cuda").permute(0, 2, 1, 3); vc = torch.zeros_like(kc)  # cache view
    # fill the past in one call, then append the new tokens (projection layout [b][tok][head][D], f32)
    for cache, X in ((kc, Kf), (vc, Vf)):
        P.kv_cache_append(to_dev(np.ascontiguousarray(X[:, :, :n_past].transpose(0, 2, 1, 3))), cache, 0, cache_type=P.TYPE_Q8_0 if q8 else None)
        P.kv_cache_append(to_dev(np.ascontiguousarray(X[:, :, n_past:].transpose(0, 2, 1, 3))), cache, n_past, cache_type=P.TYPE_Q8_0 if q8 else None)
    torch.cuda.synchronize()
    n = n_past + n_new
    if q8:  # byte-exact against the oracle's ggml quantiser
        np.testing.assert_array_equal(kc[:, :, :n].cpu().numpy(), oracle.quantize_q8_0(Kf))
        np.testing.assert_array_equal(vc[:, :, :n].cpu().numpy(), oracle.quantize_q8_0(Vf))
        Kq, Vq = oracle.quantize_q8_0(Kf), oracle.quantize_q8_0(Vf)
        ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(Kq, 8), oracle.view_of(Vq, 8), None, 1 / np.sqrt(D), round_q_f16=True)
    else:
        np.testing.assert_array_equal(kc[:, :, :n].cpu().numpy(), K)
        ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), None, 1 / np.sqrt(D), round_q_f16=True)
    out = P.flash_attn_ext(to_dev(Q), kc[:, :, :n], vc[:, :, :n], None, kv_type=P.TYPE_Q8_0 if q8 else None)
    torch.cuda.synchronize()
    assert_close(out.cpu().numpy(), ref, "decode over the appended cache")


# ---- sequence-split combine over peer-mapped memory (no NCCL): `world` ranks emulated on one device ----
@pytest.mark.parametrize("world,q8", [(1, False), (2, False), (4, True), (8, False)])
def test_peer_exchange_combine_emulated_ranks(world, q8):
    import torch
    P = pkg()
    n_kv, H, Hk, D = 4096, 32, 8, 128
    Q, K, V = synth_qkv(D, 1, n_kv, H, Hk)
    if q8:
        Kq, Vq = oracle.quantize_q8_0(K.astype(np.float32)), oracle.quantize_q8_0(V.astype(np.float32))
        ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(Kq, 8), oracle.view_of(Vq, 8), None, 1 / np.sqrt(D), round_q_f16=True)
        k, v = to_dev(Kq), to_dev(Vq)
    else:
        ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), None, 1 / np.sqrt(D), round_q_f16=True)
        k, v = to_dev(K), to_dev(V)
    q = to_dev(Q)
    xs = P.PeerExchange.local(world, H, D)
    try:
        for epoch in (1, 2, 3):  # three steps: both generations of the gathered area and the monotonic counters
            for r in range(world):
                sh = P.seq_shard(n_kv, r, world)
                P.flash_attn_partial_scatter(q, k[:, :, sh.kv_pos0:sh.kv_pos0 + sh.n_local], v[:, :, sh.kv_pos0:sh.kv_pos0 + sh.n_local], xs[r],
                                             kv_pos0=sh.kv_pos0, n_kv_total=n_kv)
            for r in range(world):
                out = P.merge_partials_wait(xs[r])
                torch.cuda.synchronize()
                assert_close(out.cpu().numpy().reshape(ref.shape), ref, f"peer exchange world={world} rank={r} epoch={epoch}")
    finally:
        for x in xs:
            x.close()


def test_fused_seqpar_step_world_1():
    """The one-kernel sequence-parallel step (decode + NVLink scatter + arrival wait + merge).  On one GPU only world = 1 can
    run it (a rank waits for its peers inside the kernel); tests/multi_gpu_check.py runs it on real peers."""
    import torch
    P = pkg()
    n_kv, H, Hk, D = 3000, 32, 8, 128
    Q, K, V = synth_qkv(D, 1, n_kv, H, Hk)
    ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), None, 1 / np.sqrt(D), round_q_f16=True)
    q, k, v = to_dev(Q), to_dev(K), to_dev(V)
    xs = P.PeerExchange.local(1, H, D)
    try:
        for step in range(3):
            out = P.flash_attn_seqpar(q, k, v, xs[0])
            torch.cuda.synchronize()
            assert P.last_dispatch() == "decode_stream_seqpar" and P.last_launch_count() == 1
            assert_close(out.cpu().numpy().reshape(ref.shape), ref, f"fused seqpar step {step}")
    finally:
        xs[0].close()


def test_graph_replay_back_to_back_calls_share_a_workspace():
    """Decode and prefill calls recorded back to back in one CUDA graph on one workspace (programmatic dependent launch lets each
    kernel's prologue overlap the previous kernel's tail; the split-KV counters are reused from call to call): every replay must
    reproduce the eagerly computed results."""
    import torch
    P = pkg()
    shapes = [(1, 4096, 32, 32), (1, 1000, 32, 8), (1, 8192, 8, 8), (256, 256, 4, 4), (1, 4096, 32, 32)]
    calls = []
    ws_bytes = 0
    for i, (n_q, n_kv, H, Hk) in enumerate(shapes):
        Q, K, V = synth_qkv(128, n_q, n_kv, H, Hk, seeds=(50 + i, 60 + i, 70 + i))
        q, k, v = to_dev(Q), to_dev(K), to_dev(V)
        ws_bytes = max(ws_bytes, P.workspace_size(0, 1, 128, n_q, H, 1, n_kv, Hk, 1))
        calls.append((q, k, v))
    ws = P.Workspace(ws_bytes)
    eager = []
    for (q, k, v) in calls:
        eager.append(P.flash_attn_ext(q, k, v, None, workspace=ws, flags=P.FLAG_WORKSPACE_ZEROED).clone())
    torch.cuda.synchronize()
    outs = [torch.zeros_like(e) for e in eager]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for (q, k, v), o in zip(calls, outs):  # warm-up on the capture stream
            P.flash_attn_ext(q, k, v, None, dst=o, workspace=ws, flags=P.FLAG_WORKSPACE_ZEROED, stream=s)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for rep in range(3):
                for (q, k, v), o in zip(calls, outs):
                    P.flash_attn_ext(q, k, v, None, dst=o, workspace=ws, flags=P.FLAG_WORKSPACE_ZEROED, stream=s)
    for rep in range(4):
        for o in outs:
            o.fill_(float("nan"))
        g.replay()
        torch.cuda.synchronize()
        for e, o in zip(eager, outs):
            assert torch.equal(e, o), f"replay {rep}: graph result differs from the eager one"


@pytest.mark.parametrize("n_q,H,Hk,B,n_kv,kind,q8", [
    (8, 32, 8, 2, 1500, "causal", False), (5, 16, 2, 1, 3000, "noise", False), (16, 8, 2, 3, 777, "causal", False),
    (4, 32, 4, 1, 2048, "zeros", True), (9, 12, 3, 1, 1000, "causal", False), (3, 24, 3, 2, 640, "none", True),
    (4, 12, 2, 2, 1500, "causal", False), (3, 36, 3, 1, 900, "noise", False), (5, 10, 2, 1, 2048, "causal", True)])
def test_gqa_bursts_17_to_64_rows(n_q, H, Hk, B, n_kv, kind, q8):
    """A burst of query positions under GQA (17..64 rows per KV head).  Power-of-two groups: the tensor-core kernel with the group's q
    heads packed into one 128-row tile (one pass over K/V per KV head).  Other group sizes (6, 12, 5 here): every KV head is split
    into virtual heads of <= 16 rows for the stream kernel."""
    Q, K, V = synth_qkv(128, n_q, n_kv, H, Hk, n_batch=B)
    mask = make_mask(kind, n_q, n_kv)
    flags = pkg().FLAG_CAUSAL if kind == "causal" else 0
    run_both(Q, K, V, mask, flags=flags, q8=q8, mask_pad=32 if mask is not None else None)
    gqa = H // Hk
    assert pkg().last_dispatch() == ("prefill_tcgen05" if gqa & (gqa - 1) == 0 else "decode_stream"), pkg().last_dispatch()
    a, _ = run_both(Q, K, V, mask, flags=flags, q8=q8, cache_view=not q8, dst_f16=True)


def test_kv_cache_append_bounds_and_padded_head_sizes():
    """appending past the end of the cache is an error (not an out-of-bounds write), and an f16 cache takes every head
    size the attention entries take (multiples of 8), e.g. 80."""
    import torch
    P = pkg()
    B, Hk, n_max = 2, 3, 64
    for D in (80, 128):
        cache = torch.zeros((B, Hk, n_max, D), dtype=torch.float16, device="cuda")
        guard = cache.clone()
        src = (torch.rand((B, 10, Hk, D), device="cuda") * 2 - 1)
        P.kv_cache_append(src, cache, 54)   # rows 54..63: exactly fits
        torch.cuda.synchronize()
        assert torch.equal(cache[:, :, 54:64], src.permute(0, 2, 1, 3).half())
        assert torch.equal(cache[:, :, :54], guard[:, :, :54])
        with pytest.raises(P.B200FAError):
            P.kv_cache_append(src, cache, 55)   # one row too many
    q8 = torch.zeros((B, Hk, n_max, 80 // 32 * 34), dtype=torch.uint8, device="cuda")
    with pytest.raises(P.B200FAError):  # q8_0 rows are whole 32-element blocks: 80 is not a q8_0 head size
        P.kv_cache_append(torch.zeros((B, 1, Hk, 80), device="cuda"), q8, 0, cache_type=P.TYPE_Q8_0)


def test_seqpar_argument_errors_and_emulated_exchange():
    """b200fa_flash_attn_seqpar validates before it plans (ne12 = 0 used to be a division by zero), and the fused
    one-kernel step refuses an exchange whose ranks live on one device (it could only dead-lock there)."""
    import ctypes as C
    import torch
    P = pkg()
    xs = P.PeerExchange.local(2, 32, 128)
    q = torch.zeros((1, 32, 1, 128), device="cuda"); k = torch.zeros((1, 8, 256, 128), dtype=torch.float16, device="cuda")
    with pytest.raises(P.B200FAError):
        P.flash_attn_seqpar(q, k, k, xs[0], kv_pos0=0, n_kv_total=512)
    ws = P.Workspace(1 << 20)
    x1 = P.PeerExchange.local(1, 32, 128)[0]
    args = [q.data_ptr(), k.data_ptr(), k.data_ptr(), None, q.data_ptr(), C.c_float(0.1), 0, 1, 0,
            128, 1, 32, 1, 128, 256, 0, 1,   # ne12 = 0
            0, 0, 512, 512, 16384, 256, 65536, 524288, 256, 65536, 524288, 0, 256,
            x1.own_ptr, x1.peers_dev.data_ptr(), 0, 1, 0, ws.ptr, ws.nbytes, None]
    assert P.lib().b200fa_flash_attn_seqpar(*args) == -1
    for x in xs + [x1]:
        x.close()


def test_peer_exchange_timeout_raises_a_flag_instead_of_a_trap():
    """a wait for peers that never arrive ends after the configured timeout with the exchange's error flag raised and a
    healthy context; after b200fa_peer_reset the exchange works again."""
    import torch
    P = pkg()
    D, rows = 128, 32
    xs = P.PeerExchange.local(2, rows, D)   # rank 1 never publishes
    Q, K, V = synth_qkv(D, 1, 512, 32, 8)
    q, k, v = to_dev(Q), to_dev(K), to_dev(V)
    xs[0].set_timeout(50)
    dst = torch.full((rows, D), 7.0, device="cuda")
    P.flash_attn_partial_scatter(q, k, v, xs[0], kv_pos0=0, n_kv_total=1024)
    P.merge_partials_wait(xs[0], dst=dst)
    torch.cuda.synchronize()            # no sticky error: the kernel ended normally
    assert xs[0].timed_out() and not xs[1].timed_out()
    assert float(dst.min()) == 7.0      # dst untouched
    for x in xs:
        x.reset()
    assert not xs[0].timed_out()
    # a complete step on the reset exchange: both emulated ranks publish, then both merge
    ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), None, 1 / np.sqrt(D), round_q_f16=True)
    for r in range(2):
        P.flash_attn_partial_scatter(q, k[:, :, r * 256:(r + 1) * 256], v[:, :, r * 256:(r + 1) * 256], xs[r], kv_pos0=r * 256, n_kv_total=512)
    for r in range(2):
        out = P.merge_partials_wait(xs[r])
        torch.cuda.synchronize()
        assert_close(out.cpu().numpy().reshape(ref.shape), ref, f"after reset, rank {r}")
    for x in xs:
        x.close()
