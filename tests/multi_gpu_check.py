#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun on N GPUs of one box (not collected by pytest: `-m gpu` runs on one GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
Head-parallel and sequence-parallel (f16 and q8_0) results are compared with the CPU oracle on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402
from common import assert_close, synth_qkv  # noqa: E402

P = load_package()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
D = 128
# head-parallel: Llama-3-8B GQA decode shape, small batch
Q, K, V = synth_qkv(D, 1, 1024, 32, 8, n_batch=4)
ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), None, 1 / np.sqrt(D), round_q_f16=True)
q, k, v = (torch.from_numpy(x).to(dev) for x in (Q, K, V))
out, sh = P.flash_attn_ext_head_parallel(q, k, v, None, rank, world)
torch.cuda.synchronize()
assert_close(out.cpu().numpy(), ref[:, :, sh.q_head0:sh.q_head0 + sh.n_q_heads], f"head-parallel rank {rank}")
# sequence-parallel, f16 and q8_0 K/V
for q8 in (False, True):
    n_kv = 8192
    Q, K, V = synth_qkv(D, 1, n_kv, 32, 8)
    ss = P.seq_shard(n_kv, rank, world)
    if q8:
        Kq, Vq = oracle.quantize_q8_0(K.astype(np.float32)), oracle.quantize_q8_0(V.astype(np.float32))
        ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(Kq, 8), oracle.view_of(Vq, 8), None, 1 / np.sqrt(D), round_q_f16=True)
        kl = torch.from_numpy(np.ascontiguousarray(Kq[:, :, ss.kv_pos0:ss.kv_pos0 + ss.n_local])).to(dev)
        vl = torch.from_numpy(np.ascontiguousarray(Vq[:, :, ss.kv_pos0:ss.kv_pos0 + ss.n_local])).to(dev)
    else:
        ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), None, 1 / np.sqrt(D), round_q_f16=True)
        kl = torch.from_numpy(np.ascontiguousarray(K[:, :, ss.kv_pos0:ss.kv_pos0 + ss.n_local])).to(dev)
        vl = torch.from_numpy(np.ascontiguousarray(V[:, :, ss.kv_pos0:ss.kv_pos0 + ss.n_local])).to(dev)
    out = P.flash_attn_ext_seq_parallel(torch.from_numpy(Q).to(dev), kl, vl, None, n_kv, rank, world)
    torch.cuda.synchronize()
    assert_close(out.cpu().numpy().reshape(ref.shape), ref, f"seq-parallel q8={q8} rank {rank}")
# sequence-parallel over peer-mapped memory (cudaIpc + NVLink stores, no NCCL on the data path)
n_kv = 8192
Q, K, V = synth_qkv(D, 1, n_kv, 32, 8)
ref = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(K), oracle.view_of(V), None, 1 / np.sqrt(D), round_q_f16=True)
ss = P.seq_shard(n_kv, rank, world)
kl = torch.from_numpy(np.ascontiguousarray(K[:, :, ss.kv_pos0:ss.kv_pos0 + ss.n_local])).to(dev)
vl = torch.from_numpy(np.ascontiguousarray(V[:, :, ss.kv_pos0:ss.kv_pos0 + ss.n_local])).to(dev)
qd = torch.from_numpy(Q).to(dev)
xch = P.PeerExchange.distributed(32, D)
dist.barrier()
for epoch in (1, 2, 3, 4):
    P.flash_attn_partial_scatter(qd, kl, vl, xch, kv_pos0=ss.kv_pos0, n_kv_total=n_kv)
    out = P.merge_partials_wait(xch)
    torch.cuda.synchronize()
    assert_close(out.cpu().numpy().reshape(ref.shape), ref, f"peer-exchange seq-parallel rank {rank} epoch {epoch}")
# the same step fused into ONE kernel per rank (f16 and q8_0)
for q8 in (False, True):
    if q8:
        Kq, Vq = oracle.quantize_q8_0(K.astype(np.float32)), oracle.quantize_q8_0(V.astype(np.float32))
        ref2 = oracle.flash_attn_ext(oracle.view_of(Q), oracle.view_of(Kq, 8), oracle.view_of(Vq, 8), None, 1 / np.sqrt(D), round_q_f16=True)
        kl2 = torch.from_numpy(np.ascontiguousarray(Kq[:, :, ss.kv_pos0:ss.kv_pos0 + ss.n_local])).to(dev)
        vl2 = torch.from_numpy(np.ascontiguousarray(Vq[:, :, ss.kv_pos0:ss.kv_pos0 + ss.n_local])).to(dev)
    else:
        ref2, kl2, vl2 = ref, kl, vl
    for step in range(4):
        out = P.flash_attn_seqpar(qd, kl2, vl2, xch, kv_pos0=ss.kv_pos0, n_kv_total=n_kv)
        torch.cuda.synchronize()
        assert P.last_dispatch() == "decode_stream_seqpar" and P.last_launch_count() == 1
        assert_close(out.cpu().numpy().reshape(ref2.shape), ref2, f"fused seq-parallel q8={q8} rank {rank} step {step}")
dist.barrier()
xch.close()
if rank == 0:
    print(f"multi_gpu_check ok on {world} GPUs: head-parallel + sequence-parallel (f16, q8_0; NCCL and peer-memory combine) match the oracle")
dist.destroy_process_group()
