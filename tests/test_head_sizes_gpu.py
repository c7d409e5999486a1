"""-m gpu: head sizes other than the reference's 128 (the reference pads 80 -> 128 by hand,
flash-matrix.cu:33-35).  Every multiple of 8 up to 128 runs on the 64/128-wide kernels with the padding done on the fly
(TMA out-of-bounds zero fill / zero-filled chunks); results are compared with the CPU oracle on the UNPADDED tensors."""
import numpy as np
import pytest

from common import make_mask, synth_qkv
from gpu_common import pkg, run_both, to_dev

pytestmark = pytest.mark.gpu

SIZES = [8, 40, 72, 80, 96, 112]


@pytest.mark.parametrize("D", SIZES)
@pytest.mark.parametrize("n_kv,H,Hk,B", [(1000, 8, 8, 1), (4099, 8, 2, 2), (63, 4, 1, 1)])
def test_decode(D, n_kv, H, Hk, B):
    Q, K, V = synth_qkv(D, 1, n_kv, H, Hk, n_batch=B)
    run_both(Q, K, V, make_mask("noise", 1, n_kv), mask_pad=32)
    assert pkg().last_dispatch() == "decode_stream"


@pytest.mark.parametrize("D", SIZES)
@pytest.mark.parametrize("cache_view", [False, True])
def test_decode_cache_view_f16_io(D, cache_view):
    Q, K, V = synth_qkv(D, 1, 777, 6, 3)
    run_both(Q, K, V, None, q_f16=True, dst_f16=True, cache_view=cache_view)


@pytest.mark.parametrize("D", SIZES)
@pytest.mark.parametrize("n_q", [5, 20, 40])
def test_burst_rows(D, n_q):
    """2..64 live rows per KV head: stream kernel up to 16 rows, the rows16 kernel above."""
    n_kv = 300 + n_q
    Q, K, V = synth_qkv(D, n_q, n_kv, 4, 2)
    run_both(Q, K, V, make_mask("causal", n_q, n_kv), flags=pkg().FLAG_CAUSAL)


@pytest.mark.parametrize("D", SIZES + [64])
@pytest.mark.parametrize("n_q,n_kv,kind", [(300, 300, "causal"), (130, 517, "noise"), (256, 256, "none")])
def test_prefill(D, n_q, n_kv, kind):
    Q, K, V = synth_qkv(D, n_q, n_kv, 4, 2)
    flags = pkg().FLAG_CAUSAL if kind == "causal" else 0
    run_both(Q, K, V, make_mask(kind, n_q, n_kv), flags=flags)
    assert pkg().last_dispatch() == "prefill_tcgen05"


@pytest.mark.parametrize("D", [80, 96])
def test_prefill_f16_io_batches(D):
    Q, K, V = synth_qkv(D, 200, 264, 4, 4, n_batch=2)
    run_both(Q, K, V, make_mask("causal", 200, 264), q_f16=True, dst_f16=True, cache_view=True)
    assert pkg().last_dispatch() == "prefill_tcgen05"


def test_neighbours_untouched():
    """dst rows are Dr wide: a guard band after the output must stay as it was (nothing is written at the padded width)."""
    import torch
    P = pkg()
    for (D, n_q, n_kv) in [(80, 1, 500), (80, 24, 200), (80, 256, 256)]:
        Q, K, V = synth_qkv(D, n_q, n_kv, 4, 4)
        buf = torch.full((n_q * 4 * D + 4096,), 7.0, device="cuda")
        dst = buf[: n_q * 4 * D].view(1, n_q, 4, D)
        P.flash_attn_ext(to_dev(Q), to_dev(K), to_dev(V), None, dst=dst)
        torch.cuda.synchronize()
        assert bool((buf[n_q * 4 * D:] == 7.0).all()), (D, n_q)
        assert bool((dst != 7.0).all())


def test_unsupported_sizes():
    import torch
    P = pkg()
    for D in (12, 132, 264, 512):
        q = torch.zeros((1, 2, 1, D), device="cuda"); k = torch.zeros((1, 2, 64, D), device="cuda", dtype=torch.float16)
        with pytest.raises(P.B200FAError):
            P.flash_attn_ext(q, k, k, None)
    # q8_0 K/V: 64, 128 or 256 only
    q = torch.zeros((1, 2, 1, 96), device="cuda"); k = torch.zeros((1, 2, 64, 96 // 32 * 34), device="cuda", dtype=torch.uint8)
    with pytest.raises(P.B200FAError):
        P.flash_attn_ext(q, k, k, None)


# ---- head sizes 129..256 (SURVEY.md §8f row 3 lists 256): the 256-wide instantiation of the 16-row register-streaming kernel ----
BIG = [256, 160, 192]


@pytest.mark.parametrize("D", BIG)
@pytest.mark.parametrize("n_kv,H,Hk,B", [(1000, 8, 8, 1), (4099, 8, 2, 2), (63, 4, 1, 1)])
def test_decode_head_sizes_up_to_256(D, n_kv, H, Hk, B):
    Q, K, V = synth_qkv(D, 1, n_kv, H, Hk, n_batch=B)
    run_both(Q, K, V, make_mask("noise", 1, n_kv), mask_pad=32)
    assert pkg().last_dispatch() == "decode_splitkv"


@pytest.mark.parametrize("n_kv,H,Hk", [(2048, 8, 2), (777, 4, 4)])
def test_decode_head_size_256_q8_0_and_f16_io(n_kv, H, Hk):
    Q, K, V = synth_qkv(256, 1, n_kv, H, Hk)
    run_both(Q, K, V, None, q8=True)
    run_both(Q, K, V, None, q_f16=True, dst_f16=True, cache_view=True)


@pytest.mark.parametrize("D", BIG)
@pytest.mark.parametrize("n_q,n_kv,kind", [(5, 305, "causal"), (40, 340, "causal"), (300, 300, "causal"), (130, 517, "noise")])
def test_bursts_and_prefill_head_sizes_up_to_256(D, n_q, n_kv, kind):
    """More than 16 query positions at these head sizes also run on the 16-row kernel (row groups x KV splits): correct, not fast —
    the tcgen05 tile kernel is 128 wide."""
    Q, K, V = synth_qkv(D, n_q, n_kv, 4, 2)
    flags = pkg().FLAG_CAUSAL if kind == "causal" else 0
    run_both(Q, K, V, make_mask(kind, n_q, n_kv), flags=flags)
    assert pkg().last_dispatch() in ("decode_splitkv", "rows16_mma")


def test_neighbours_untouched_head_size_256_and_192():
    import torch
    P = pkg()
    for (D, n_q, n_kv) in [(256, 1, 500), (192, 24, 200), (192, 1, 300)]:
        Q, K, V = synth_qkv(D, n_q, n_kv, 4, 4)
        buf = torch.full((n_q * 4 * D + 4096,), 7.0, device="cuda")
        dst = buf[: n_q * 4 * D].view(1, n_q, 4, D)
        P.flash_attn_ext(to_dev(Q), to_dev(K), to_dev(V), None, dst=dst)
        torch.cuda.synchronize()
        assert bool((buf[n_q * 4 * D:] == 7.0).all()), (D, n_q)
        assert bool((dst != 7.0).all())
