"""GQA-packed tiles of the tensor-core kernel (csrc/prefill_persistent.cuh, pp_pack_shift): the q heads of a KV head share one
128-row tile, row r = query position r / gqa, head r % gqa.  The reference has no counterpart (its kernel re-reads K/V per q head,
flash-llama.h:128-151); parity is against the oracle like everywhere else."""
import numpy as np
import pytest

from common import make_mask, synth_qkv
from gpu_common import pkg, run_both

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("gqa", [2, 4, 8, 16, 32])
@pytest.mark.parametrize("n_q,n_kv,kind", [(1, 700, "none"), (7, 1000, "causal"), (33, 2000, "noise"), (64, 1536, "causal")])
def test_every_power_of_two_group(gqa, n_q, n_kv, kind):
    """Every group size the packing takes, from one position (rows = gqa) to several tile pairs; flag-less masks as well."""
    if n_q * gqa <= 16:
        pytest.skip("<= 16 rows per KV head: the stream kernel's shape")
    Hk = 2 if gqa < 32 else 1
    Q, K, V = synth_qkv(128, n_q, n_kv, Hk * gqa, Hk, n_batch=2)
    mask = make_mask(kind, n_q, n_kv)
    run_both(Q, K, V, mask, flags=pkg().FLAG_CAUSAL if kind == "causal" else 0, mask_pad=32 if mask is not None else None)
    packs = (n_q + 128 // gqa - 1) // (128 // gqa)
    plain = (n_q + 127) // 128
    expect_packed = (packs + 1) // 2 < ((plain + 1) // 2) * gqa
    assert pkg().last_dispatch() == "prefill_tcgen05"
    if kind == "causal":  # the same through the mask tensor alone (exact-causal detection) and with f16 output on a cache view
        run_both(Q, K, V, mask, flags=0, mask_pad=32, cache_view=True, dst_f16=True)
    assert expect_packed or n_q > 16


@pytest.mark.parametrize("n_q,H,Hk,B,n_kv", [(8, 32, 8, 1, 8192), (16, 32, 8, 2, 4100), (5, 16, 4, 1, 16384), (40, 8, 2, 1, 6000)])
def test_packed_tiles_with_split_kv(n_q, H, Hk, B, n_kv):
    """Few packed items and a long KV range: the KV tiles of every item are cut into segments (partial rows + combine launch)."""
    Q, K, V = synth_qkv(128, n_q, n_kv, H, Hk, n_batch=B)
    mask = make_mask("causal", n_q, n_kv)
    run_both(Q, K, V, mask, flags=pkg().FLAG_CAUSAL, mask_pad=32)
    assert pkg().last_dispatch() == "prefill_tcgen05" and pkg().last_launch_count() >= 2   # attention + combine (+ Q conversion)
    run_both(Q, K, V, make_mask("noise", n_q, n_kv), mask_pad=32, q_f16=True)


@pytest.mark.parametrize("D", [64, 80, 96])
def test_packed_tiles_with_padded_head_sizes(D):
    Q, K, V = synth_qkv(D, 12, 900, 16, 4, n_batch=2)
    run_both(Q, K, V, make_mask("causal", 12, 900), flags=pkg().FLAG_CAUSAL, mask_pad=32)
    assert pkg().last_dispatch() == "prefill_tcgen05"
    run_both(Q, K, V, make_mask("noise", 12, 900), mask_pad=32, dst_f16=True)


def test_packed_tiles_q8_0_cache_and_neighbours_untouched():
    """q8_0 K/V (dequantised copies) under packing; and the rows of dst that belong to other calls keep their contents."""
    import torch
    P = pkg()
    Q, K, V = synth_qkv(128, 6, 1280, 16, 2, n_batch=1)
    run_both(Q, K, V, make_mask("causal", 6, 1280), flags=P.FLAG_CAUSAL, q8=True, mask_pad=32)
    assert P.last_dispatch() == "prefill_tcgen05"
    # guard band: dst is a slice of a larger buffer filled with a sentinel
    q = torch.from_numpy(Q).cuda(); k = torch.from_numpy(K).cuda(); v = torch.from_numpy(V).cuda()
    big = torch.full((3, 6, 16, 128), 7.25, device="cuda")
    P.flash_attn_ext(q, k, v, None, dst=big[1:2])
    torch.cuda.synchronize()
    assert (big[0] == 7.25).all() and (big[2] == 7.25).all()
    assert not (big[1] == 7.25).any()


def test_modifiers_keep_the_unpacked_paths():
    """ALiBi / soft-cap / mask slices are per q head: such calls are not packed (bursts stay on the stream kernel's virtual heads)."""
    import torch
    P = pkg()
    Q, K, V = synth_qkv(128, 8, 1000, 16, 4, n_batch=1)
    q, k, v = (torch.from_numpy(x).cuda() for x in (Q, K, V))
    m = torch.zeros((32, 1000), dtype=torch.float16, device="cuda")
    P.flash_attn_ext(q, k, v, m, max_bias=8.0)
    torch.cuda.synchronize()
    assert P.last_dispatch() == "decode_stream"
