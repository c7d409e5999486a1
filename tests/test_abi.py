"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads, exports every symbol
include/b200fa.h declares, and refuses to compute without an sm_100 device (no fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from __graft_entry__ import ROOT, load_package


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include/b200fa.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200fa_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ["b200fa_flash_attn_ext", "b200fa_workspace_size", "b200fa_flash_attn_partial", "b200fa_merge_partials",
              "b200fa_quantize_q8_0", "b200fa_dequantize_q8_0", "b200fa_status_string", "b200fa_version"]:
        assert s in syms


def test_library_loads_and_exports_every_declared_symbol():
    P = load_package()
    lib = P.lib()
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/b200fa.h but not exported"
    assert lib.b200fa_version() >= 100


def test_header_compiles_as_plain_c():
    src = '#include "b200fa.h"\nint main(void){return b200fa_version()>0?0:1;}\n'
    res = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                         input=src, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_sass_is_sm100a_and_has_tensor_paths():
    P = load_package()
    lib = P.build()
    out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    assert "HMMA" in sass          # decode: mma.sync fragments
    assert "UTCHMMA" in sass       # prefill: tcgen05.mma
    assert "UTMALDG" in sass       # prefill: TMA loads
    assert "LDTM" in sass and "STTM" in sass  # TMEM traffic


def test_status_strings():
    P = load_package()
    lib = P.lib()
    assert lib.b200fa_status_string(0) == b"ok"
    for s in (-1, -2, -3, -4):
        assert len(lib.b200fa_status_string(s)) > 0


def test_argument_validation_needs_no_gpu():
    P = load_package()
    lib = P.lib()
    i64 = C.c_int64
    def call(q=1 << 20, k=1 << 21, v=1 << 22, dst=1 << 23, D=128, n_q=1, H=32, Hk=8, n_kv=256, qt=0, kt=1, dt=0, nb11=256):
        return lib.b200fa_flash_attn_ext(q, k, v, None, dst, 0.1, qt, kt, dt, D, n_q, H, 1, D, n_kv, Hk, 1, 0, 0,
                                         D * 4, D * 4 * n_q, D * 4 * n_q * H, nb11, nb11 * n_kv, nb11 * n_kv * Hk,
                                         nb11, nb11 * n_kv, nb11 * n_kv * Hk, D, H, n_q, 1, 0, None, 0, None)
    assert call(q=None) == -1                 # NULL pointer
    assert call(H=30) == -1                   # heads not a multiple of kv heads
    assert call(nb11=250) == -1               # misaligned rows
    assert call(D=264, nb11=528) == -2        # head size not built (above 256)
    assert call(D=100) == -2                  # head size not a multiple of 8
    assert call(D=96, kt=8, nb11=102) == -2   # q8_0 K/V: 64, 128 or 256 only
    assert call(kt=2) == -2                   # type not built (q4_0)
    assert call(k=(1 << 21) + 2) == -1        # misaligned base


def test_no_device_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    P = load_package()
    lib = P.lib()
    D, n_kv = 128, 256
    rc = lib.b200fa_flash_attn_ext(1 << 20, 1 << 21, 1 << 22, None, 1 << 23, 0.1, 0, 1, 0, D, 1, 1, 1, D, n_kv, 1, 1, 0, 0,
                                   512, 512, 512, 256, 256 * n_kv, 256 * n_kv, 256, 256 * n_kv, 256 * n_kv,
                                   D, 1, 1, 1, 0, None, 0, None)
    assert rc == -4
    with pytest.raises(P.B200FAError):
        P.flash_attn_ext(torch.zeros(1, 1, 1, 128), torch.zeros(1, 1, 4, 128, dtype=torch.float16),
                         torch.zeros(1, 1, 4, 128, dtype=torch.float16))


def test_workspace_size_is_deterministic_and_covers_partials():
    P = load_package()
    a = P.workspace_size(0, 1, 128, 1, 32, 1, 4096, 32, 1)
    b = P.workspace_size(0, 1, 128, 1, 32, 1, 4096, 32, 1)
    assert a == b and a >= 32 * 130 * 4
    assert P.workspace_size(0, 1, 128, 1, 32, 1, 64, 32, 1) >= 256


def test_exchange_buffer_layout_needs_no_gpu():
    """b200fa_xchg_bytes: header, this rank's staged triples, two generations of gathered triples (NCCL-free three-launch path) and
    two generations of the fused step's flag-in-data area, where every float travels as an 8-byte {value, step tag} pair."""
    P = load_package()
    for world, rows, D in ((1, 32, 128), (2, 32, 128), (8, 256, 64)):
        n = rows * (D + 2)
        assert P.PeerExchange.nbytes(world, rows, D) == 256 + (1 + 2 * world) * n * 4 + 2 * world * n * 8
    assert P.PeerExchange.nbytes(0, 32, 128) == 0 and P.PeerExchange.nbytes(2, 0, 128) == 0
