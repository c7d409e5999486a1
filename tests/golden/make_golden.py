#!/usr/bin/env python
"""Generate the committed golden fixtures.  Run in the build container (needs /root/reference
and the `gguf` python package); the GPU box only ever reads the .json/.npz files written here.

  kat_flash_attn_f32.json   the reference's own known-answer table, parsed from
                            /root/reference/src/misc/flash-attn.cu:207-247 (inputs) and :286-293 (expected)
  ref_host_cases.npz        outputs of the reference's OWN host attention (utils.h, compiled unmodified into
                            oracle/_ref/libref_host.so) on seeded inputs, sequenced as test_llama
                            (flash-matrix.cu:88-111) and kernel_test (kernel_test.h:50-61)
  q8_0_gguf.npz             ggml q8_0 blocks + dequantised values from gguf.quants (the only offline q8_0 oracle)
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
sys.path.insert(0, os.path.dirname(HERE))
from common import make_mask  # noqa: E402

REF = "/root/reference/src"


def parse_kat():
    src = open(os.path.join(REF, "misc/flash-attn.cu")).read()

    def arr(name):
        m = re.search(r"float\s+%s\[24\]\s*=\s*\{(.*?)\};" % name, src, re.S)
        body = re.sub(r"//.*", "", m.group(1))
        return [float(x) for x in re.findall(r"-?\d+(?:\.\d+)?", body)]

    m = re.search(r"\* Expected values(.*?)\*/", src, re.S)
    expected = [float(x) for x in re.findall(r"-?\d+\.\d+", m.group(1))]
    assert len(expected) == 24
    kat = {
        "source": "reference src/misc/flash-attn.cu:207-247 (inputs), :286-293 (expected)",
        "d_head": 3, "seq_len": 4, "num_heads": 2, "scale": "1/sqrt(3)",
        "query_layout": "[head][seq][d]", "key_layout": "[head][seq][d]", "value_layout": "[head][d][seq] (transposed)",
        "query": arr("Query"), "key": arr("Key"), "value_T": arr("Value"), "expected": expected,
        "expected_layout": "[head][seq][d]", "decimals": 4,
    }
    json.dump(kat, open(os.path.join(HERE, "kat_flash_attn_f32.json"), "w"), indent=1)
    print("kat: ok")


def ref_host_cases():
    import ctypes as C
    lib = oracle.ref_host()
    out = {}

    def llama(tag, D, n_q, n_kv, n_head, n_head_kv, mask_kind, seeds=(1, 2, 3)):
        Q = oracle.uniform_pm1(seeds[0], (n_head, n_q, D))
        K = oracle.uniform_pm1(seeds[1], (n_head_kv, n_kv, D)).astype(np.float16)
        V = oracle.uniform_pm1(seeds[2], (n_head_kv, n_kv, D)).astype(np.float16)
        VT = np.ascontiguousarray(V.transpose(0, 2, 1))
        mask = make_mask(mask_kind, n_q, n_kv)
        res = np.zeros((n_q, n_head, D), np.float32)
        scores = np.zeros((n_head, n_q, n_kv), np.float32)
        rc = lib.ref_host_attention_llama(Q.ctypes.data, K.ctypes.data, VT.ctypes.data,
                                          mask.ctypes.data if mask is not None else None,
                                          res.ctypes.data, scores.ctypes.data, D, n_q, n_kv, n_head, n_head_kv,
                                          C.c_float(1.0 / np.sqrt(D)), 1)
        assert rc == 0
        out[tag] = res
        out[tag + "__meta"] = np.array([D, n_q, n_kv, n_head, n_head_kv, *seeds], np.int64)
        out[tag + "__mask"] = np.array(mask_kind)
        out[tag + "__inputsum"] = np.array([Q.astype(np.float64).sum(), K.astype(np.float64).sum(),
                                            V.astype(np.float64).sum()])

    # C1 (BASELINE.json configs[0]): 1 head, d128, n_kv 256, 1 query — zero mask and -inf tail
    llama("c1_zero", 128, 1, 256, 1, 1, "zeros")
    llama("c1_tail", 128, 1, 256, 1, 1, "tail56")
    llama("c1_nomask", 128, 1, 256, 1, 1, "none")
    # the fixture test's shape (flash-matrix.cu:76): 32 heads, kv 256, batch 1
    llama("llama_32h", 128, 1, 256, 32, 32, "zeros")
    # GQA 32/8 as kernel_test (kernel_test.h:25), several queries, causal mask, ragged kv
    llama("gqa_causal", 128, 5, 77, 32, 8, "causal")
    llama("noise_mask", 128, 3, 130, 8, 2, "noise")
    llama("d64", 64, 2, 96, 4, 4, "causal")

    # kernel_test path (kernel_test.h:45-61): all-f32 buffers rounded through f16, 1-D noise mask
    D, n_kv, n_head, n_head_kv = 128, 512, 32, 8
    q = oracle.uniform_pm1(1, (n_head, D)); k = oracle.uniform_pm1(2, (n_head_kv, n_kv, D))
    v = oracle.uniform_pm1(3, (n_head_kv, n_kv, D)); m = oracle.uniform_pm1(4, (n_kv,))
    res = np.zeros((n_head, D), np.float32); scores = np.zeros((n_head, n_kv), np.float32)
    rc = lib.ref_host_attention_ktest(q.ctypes.data, k.ctypes.data, v.ctypes.data, m.ctypes.data, res.ctypes.data,
                                      scores.ctypes.data, D, n_kv, n_head, n_head_kv, C.c_float(1.0 / np.sqrt(D)))
    assert rc == 0
    out["ktest_512"] = res
    np.savez_compressed(os.path.join(HERE, "ref_host_cases.npz"), **out)
    print("ref_host: ok", [k for k in out if "__" not in k])


def q8_0_cases():
    from gguf import quants
    from gguf.constants import GGMLQuantizationType as T
    x = oracle.uniform_pm1(7, (16, 128)).astype(np.float32)
    x[1] = 0.0                                # all-zero block -> d = 0
    x[2, :32] = np.linspace(-4, 4, 32)        # amax on both signs
    x[3] *= 1e-6                              # tiny d (subnormal f16 scale)
    x[4] *= 300.0                             # large d
    x[5, :32] = (np.arange(32) - 16) * 0.5 * (127 / 8.0) / 127 * 8 / 8  # exact .5 ties after scaling
    x[6] = np.round(x[6] * 4) / 4
    qb = quants.quantize(x, T.Q8_0)
    dq = quants.dequantize(qb, T.Q8_0)
    np.savez_compressed(os.path.join(HERE, "q8_0_gguf.npz"), x=x, blocks=qb, dequant=dq.astype(np.float32))
    print("q8_0: ok", qb.shape, dq.shape)


if __name__ == "__main__":
    oracle.build()
    parse_kat()
    ref_host_cases()
    q8_0_cases()
