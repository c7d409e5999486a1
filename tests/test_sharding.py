"""Host-side multi-GPU logic without a GPU: the partition arithmetic, and the N > 1 protocol at world
size 2 over gloo with the CPU oracle standing in for the kernels (the product itself has no CPU path)."""
import os
import socket

import numpy as np
import pytest

import oracle
from __graft_entry__ import load_package
from common import assert_close, make_mask, synth_qkv

P = load_package()


# ---------------------------------------------------------------- partition arithmetic
@pytest.mark.parametrize("n_head,n_head_kv", [(32, 8), (32, 32), (8, 1), (40, 10), (6, 3)])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8, 16])
def test_head_shards_partition_the_heads_and_keep_gqa_groups(n_head, n_head_kv, world):
    gqa = n_head // n_head_kv
    shards = [P.head_shard(n_head, n_head_kv, r, world) for r in range(world)]
    assert sum(s.n_kv_heads for s in shards) == n_head_kv and sum(s.n_q_heads for s in shards) == n_head
    pos = 0
    for s in shards:
        assert s.kv_head0 == pos and s.q_head0 == pos * gqa and s.n_q_heads == s.n_kv_heads * gqa
        # every q head of the band maps to a kv head of the band (flash-llama.h:128-140: ik2 = iq2 / gqa)
        for h in range(s.q_head0, s.q_head0 + s.n_q_heads):
            assert s.kv_head0 <= h // gqa < s.kv_head0 + s.n_kv_heads
        pos += s.n_kv_heads
    sizes = [s.n_kv_heads for s in shards]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("n_kv", [1, 63, 64, 65, 1000, 4096, 131072])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_seq_shards_partition_the_sequence_on_chunk_boundaries(n_kv, world):
    shards = [P.seq_shard(n_kv, r, world) for r in range(world)]
    pos = 0
    for s in shards:
        assert s.kv_pos0 == pos or s.n_local == 0
        assert s.kv_pos0 % 64 == 0 or s.n_local == 0
        pos += s.n_local
    assert pos == n_kv
    with pytest.raises(ValueError):
        P.seq_shard(n_kv, world, world)


# ---------------------------------------------------------------- world size 2 over gloo
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _np_partial(q, k, v, mask, scale=None, kv_pos0=0, n_kv_total=None, flags=0):
    """CPU stand-in for b200fa_flash_attn_partial: per output row (O~[D], m, l), rows ordered (batch, q, head)."""
    import torch
    Q = q.numpy().astype(np.float16).astype(np.float64); K = k.numpy().astype(np.float64); V = v.numpy().astype(np.float64)
    n_b, H, n_q, D = Q.shape
    Hk, n_kv = K.shape[1], K.shape[2]
    scale = scale if scale is not None else 1.0 / np.sqrt(D)
    out = np.zeros((n_b, n_q, H, D + 2), np.float32)
    for b in range(n_b):
        for h in range(H):
            for i in range(n_q):
                s = K[b, h // (H // Hk)] @ Q[b, h, i] * scale
                if mask is not None:
                    s = s + mask.numpy()[i].astype(np.float64)
                if flags & P.FLAG_CAUSAL:
                    lim = i + (n_kv_total - n_q) - kv_pos0  # local keys <= lim visible
                    s[np.arange(n_kv) > lim] = -np.inf
                m = s.max() if n_kv else -np.inf
                if not np.isfinite(m):
                    out[b, i, h, D] = -np.inf
                    continue
                e = np.exp(s - m)
                out[b, i, h, :D] = e @ V[b, h // (H // Hk)]
                out[b, i, h, D] = m; out[b, i, h, D + 1] = e.sum()
    return torch.from_numpy(out.reshape(-1, D + 2))


def _oracle_merge(parts):
    import torch
    p = parts.numpy()
    n_parts, rows, d2 = p.shape
    out = np.stack([oracle.merge_partials(p[:, r, d2 - 2], p[:, r, d2 - 1], p[:, r, :d2 - 2]) for r in range(rows)])
    return torch.from_numpy(out)


def _oracle_attn(q, k, v, mask, **kw):
    import torch
    out = oracle.flash_attn_ext(oracle.view_of(q.numpy()), oracle.view_of(k.numpy()), oracle.view_of(v.numpy()),
                                oracle.view_of(mask.numpy()) if mask is not None else None, 1.0 / np.sqrt(q.shape[-1]), round_q_f16=True)
    return torch.from_numpy(out)


def _worker(rank, world, port, case, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        D = 128
        if case == "heads":
            n_q, n_kv, H, Hk = 3, 200, 12, 3   # 3 kv heads over 2 ranks: uneven bands 2 + 1
            Q, K, V = synth_qkv(D, n_q, n_kv, H, Hk, n_batch=2)
            mask = make_mask("noise", n_q, n_kv)
            q, k, v, m = (torch.from_numpy(x) for x in (Q, K, V, mask))
            local, sh = P.flash_attn_ext_head_parallel(q, k, v, m, rank, world, attn_fn=_oracle_attn)
            full = torch.zeros((2, n_q, H, D))
            full[:, :, sh.q_head0:sh.q_head0 + sh.n_q_heads] = local      # each rank fills its own column band ...
            dist.all_reduce(full)                                           # ... the bands are disjoint, so a sum assembles dst
            ref = _oracle_attn(q, k, v, m)
            ret[rank] = float((full - ref).abs().max())
        else:
            n_q, H, Hk = (1, 8, 2) if case != "seq_causal" else (3, 4, 4)
            n_kv = {"seq": 1000, "seq_causal": 333, "seq_empty_band": 40}[case]
            Q, K, V = synth_qkv(D, n_q, n_kv, H, Hk)
            flags = P.FLAG_CAUSAL if case == "seq_causal" else 0
            mask = make_mask("causal", n_q, n_kv) if flags else None
            sh = P.seq_shard(n_kv, rank, world)
            q = torch.from_numpy(Q)
            kl = torch.from_numpy(K[:, :, sh.kv_pos0:sh.kv_pos0 + sh.n_local]); vl = torch.from_numpy(V[:, :, sh.kv_pos0:sh.kv_pos0 + sh.n_local])
            out = P.flash_attn_ext_seq_parallel(q, kl, vl, None, n_kv, rank, world, partial_fn=_np_partial, merge_fn=_oracle_merge, flags=flags)
            ref = _oracle_attn(q, torch.from_numpy(K), torch.from_numpy(V), torch.from_numpy(mask) if mask is not None else None)
            ret[rank] = float((out.reshape(ref.shape) - ref).abs().max())
            if case == "seq_empty_band":
                assert P.seq_shard(n_kv, 1, world).n_local == 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["heads", "seq", "seq_causal", "seq_empty_band"])
def test_world_size_2_gloo(case):
    import torch.multiprocessing as mp
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), case, ret), nprocs=world, join=True)
        assert len(ret) == world
        for r in range(world):
            assert ret[r] < 1e-5, (case, r, ret[r])
