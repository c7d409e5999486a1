"""ctypes front-end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` /
``--impl reference`` legs — never by the product package.  See ``attn_oracle.c`` for the
reference file:line each function restates.

Tensors are described the ggml way: ``ne`` = element counts (fastest dimension first) and
``nb`` = byte strides, exactly the arguments of the reference kernel (flash-llama.h:6-32).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TYPE_F32, TYPE_F16, TYPE_Q8_0 = 0, 1, 8
QK8_0, Q8_0_BLOCK_BYTES = 32, 34

_libs: dict[str, C.CDLL] = {}


def build(with_ref: bool = True) -> None:
    """Compile the checkers (plain-C oracle always; reference shims when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", HERE, "all" if with_ref else os.path.join(HERE, "_build/liboracle.so")],
                   check=True)


def _load(rel: str, build_if_missing: bool = True) -> C.CDLL:
    if rel not in _libs:
        path = os.path.join(HERE, rel)
        if not os.path.exists(path) and build_if_missing:
            build()
        _libs[rel] = C.CDLL(path)
    return _libs[rel]


def lib() -> C.CDLL:
    l = _load("_build/liboracle.so")
    l.oracle_flash_attn_ext.restype = C.c_int
    l.oracle_flash_attn_ext.argtypes = ([C.c_void_p] * 5 + [C.c_float] + [C.c_int] * 3 + [C.c_int64] * 23 + [C.c_int] * 3)
    l.oracle_flash_attn_ext2.restype = C.c_int
    l.oracle_flash_attn_ext2.argtypes = ([C.c_void_p] * 5 + [C.c_float] + [C.c_int] * 3 + [C.c_int64] * 23 + [C.c_int] * 3 + [C.c_float] * 2)
    l.oracle_flash_attn_ext3.restype = C.c_int
    l.oracle_flash_attn_ext3.argtypes = ([C.c_void_p] * 5 + [C.c_float] + [C.c_int] * 3 + [C.c_int64] * 23 + [C.c_int] * 3 + [C.c_float] * 2 + [C.c_int64] * 4)
    return l


def ref_host_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref/libref_host.so")) or os.path.exists("/root/reference/src/utils.h")


def ref_host() -> C.CDLL:
    l = _load("_ref/libref_host.so")
    l.ref_host_attention_llama.restype = C.c_int
    l.ref_host_attention_llama.argtypes = [C.c_void_p] * 6 + [C.c_int] * 5 + [C.c_float, C.c_int]
    l.ref_host_attention_ktest.restype = C.c_int
    l.ref_host_attention_ktest.argtypes = [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_float]
    if hasattr(l, "ref_host_load_tensor"):
        l.ref_host_load_tensor.restype = C.c_int
        l.ref_host_load_tensor.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_char_p, C.c_void_p, C.c_int64]
    return l


def ref_gpu_path() -> str:
    return os.path.join(HERE, "_ref/libref_gpu.so")


@dataclass
class View:
    """A ggml-style strided view over a numpy buffer (keeps the buffer alive)."""
    buf: np.ndarray
    ne: tuple  # (ne0, ne1, ne2, ne3) elements
    nb: tuple  # (nb0, nb1, nb2, nb3) bytes
    type: int
    offset: int = 0

    @property
    def ptr(self) -> int:
        return self.buf.ctypes.data + self.offset


def view_of(a: np.ndarray, type_: int | None = None) -> View:
    """numpy [n3][n2][n1][n0] (any strides, up to 4-D) -> ggml view with ne/nb reversed."""
    if type_ is None:
        type_ = {np.dtype(np.float32): TYPE_F32, np.dtype(np.float16): TYPE_F16}[a.dtype]
    shape = (1,) * (4 - a.ndim) + tuple(a.shape)
    strides = tuple(a.strides)
    strides = tuple([strides[0] * shape[4 - a.ndim]] * (4 - a.ndim)) + strides if a.ndim < 4 else strides
    ne = tuple(reversed(shape))
    nb = tuple(reversed(strides))
    if type_ == TYPE_Q8_0:  # bytes [..., n_rows, D/32*34] -> ne0 counts elements
        ne = (ne[0] // Q8_0_BLOCK_BYTES * QK8_0,) + ne[1:]
        nb = (Q8_0_BLOCK_BYTES,) + nb[1:]
    base = a
    while base.base is not None and isinstance(base.base, np.ndarray):
        base = base.base
    return View(base, ne, nb, type_, a.ctypes.data - base.ctypes.data)


def flash_attn_ext(q: View, k: View, v: View, mask: View | None, scale: float, dst_type: int = TYPE_F32,
                   round_q_f16: bool = False, strict_ref: bool = False, nthreads: int = 0, max_bias: float = 0.0,
                   logit_softcap: float = 0.0) -> np.ndarray:
    """softmax(scale·QKᵀ + mask)·V -> numpy [ne03][n_q][n_head][D]  (flash-llama.h:434 layout).
    max_bias / logit_softcap: upstream ggml extensions (ALiBi slopes on the mask, tanh soft-cap) — not in the reference.
    mask: a 2-D view is the reference's shared mask; a 4-D view [ne33][ne32][rows][n_kv] carries one slice per head (ne32 = n_head or 1)
    and per batch entry (ne33 = n_batch or 1) — upstream ggml's broadcast, not in the reference either."""
    D, n_q, n_head, n_b = q.ne
    out = np.empty((n_b, n_q, n_head, D), np.float16 if dst_type == TYPE_F16 else np.float32)
    if nthreads <= 0:
        nthreads = os.cpu_count() or 1
    rc = lib().oracle_flash_attn_ext3(
        q.ptr, k.ptr, v.ptr, mask.ptr if mask is not None else None, out.ctypes.data, scale,
        q.type, k.type, dst_type,
        *q.ne, *k.ne,
        mask.ne[1] if mask is not None else 0, mask.nb[1] if mask is not None else 0,
        q.nb[1], q.nb[2], q.nb[3], k.nb[1], k.nb[2], k.nb[3], v.nb[1], v.nb[2], v.nb[3],
        D, n_head, n_q, n_b,
        int(round_q_f16), int(strict_ref), nthreads, float(max_bias), float(logit_softcap),
        mask.ne[2] if mask is not None else 1, mask.ne[3] if mask is not None else 1,
        mask.nb[2] if mask is not None else 0, mask.nb[3] if mask is not None else 0)
    if rc != 0:
        raise ValueError(f"oracle_flash_attn_ext rejected the arguments (rc={rc})")
    return out


def quantize_q8_0(x: np.ndarray) -> np.ndarray:
    """f32 [..., D] -> uint8 [..., D/32*34] in ggml block_q8_0 layout."""
    x = np.ascontiguousarray(x, np.float32)
    assert x.shape[-1] % QK8_0 == 0
    y = np.empty(x.shape[:-1] + (x.shape[-1] // QK8_0 * Q8_0_BLOCK_BYTES,), np.uint8)
    lib().oracle_quantize_q8_0(C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), C.c_int64(x.size))
    return y


def dequantize_q8_0(y: np.ndarray) -> np.ndarray:
    y = np.ascontiguousarray(y, np.uint8)
    assert y.shape[-1] % Q8_0_BLOCK_BYTES == 0
    x = np.empty(y.shape[:-1] + (y.shape[-1] // Q8_0_BLOCK_BYTES * QK8_0,), np.float32)
    lib().oracle_dequantize_q8_0(C.c_void_p(y.ctypes.data), C.c_void_p(x.ctypes.data), C.c_int64(x.size))
    return x


def merge_partials(m: np.ndarray, l: np.ndarray, O: np.ndarray) -> np.ndarray:
    """fa_reduce algebra (flash_row_float.h:429-471) in fp32: m,l [P]; O [P][D] -> [D]."""
    m = np.ascontiguousarray(m, np.float32); l = np.ascontiguousarray(l, np.float32)
    O = np.ascontiguousarray(O, np.float32)
    out = np.empty(O.shape[1], np.float32)
    lib().oracle_merge_partials(C.c_void_p(m.ctypes.data), C.c_void_p(l.ctypes.data), C.c_void_p(O.ctypes.data),
                                C.c_int64(O.shape[0]), C.c_int64(O.shape[1]), C.c_void_p(out.ctypes.data))
    return out


def f32_to_f16_bits(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty(x.shape, np.uint16)
    lib().oracle_f32_to_f16(C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), C.c_int64(x.size))
    return y


def f16_bits_to_f32(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.uint16)
    y = np.empty(x.shape, np.float32)
    lib().oracle_f16_to_f32(C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), C.c_int64(x.size))
    return y


# ---- deterministic synthetic data: the reference recipe 1 - 2*rand()/RAND_MAX (utils.h:57-61),
# with rand() replaced by a fixed-seed splitmix64 stream so runs are reproducible.
def uniform_pm1(seed: int, shape) -> np.ndarray:
    n = int(np.prod(shape))
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed) * np.uint64(0xD1B54A32D192ED03)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / (1 << 24))  # [0,1)
    return (np.float32(1.0) - np.float32(2.0) * u).reshape(shape)
