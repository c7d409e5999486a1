"""The compiled CUDA C++ caller of the C ABI (examples/kernel_test_dropin.cu): the reference's `kernel_test` flow
(kernel_test.h:25-61, 161-162, 191-198) with the kernel launches replaced by b200fa_flash_attn_ext.
CPU: it compiles and links against include/b200fa.h + libb200fa.so.  GPU: it runs and its own host check passes."""
import re
import subprocess

import pytest

from gpu_common import pkg


def test_dropin_example_builds_and_links():
    path = pkg().build_example()
    out = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
    assert "libb200fa.so" in out and "not found" not in out.split("libb200fa.so")[1].split("\n")[0], out


@pytest.mark.gpu
@pytest.mark.parametrize("kv", [256, 4096, 1000])
def test_dropin_example_runs_and_matches_its_host_attention(kv):
    path = pkg().build_example()
    res = subprocess.run([path, str(kv)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    m = re.search(r"max_abs_diff ([0-9.eE+-]+)", res.stdout)
    assert m and float(m.group(1)) < 2e-3, res.stdout
    assert "dispatch decode_stream" in res.stdout and "launches 1" in res.stdout
