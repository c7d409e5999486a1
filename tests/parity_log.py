"""Parity numbers the -m gpu tests measure are also written down: profiles/parity_r2.json (and gpurun_out/, which is what
travels back from a GPU box), one entry per test case, merged across runs."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(name: str, entry: dict) -> None:
    for d in ("profiles", "gpurun_out"):
        path = os.path.join(ROOT, d, "parity_r2.json")
        if not os.path.isdir(os.path.dirname(path)):
            continue
        try:
            data = json.load(open(path))
        except Exception:  # noqa: BLE001
            data = {}
        data[name] = entry
        with open(path, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)


def three_way(name, ours, refgpu, ref32, atol=2e-3, rtol=1e-2, note=""):
    """ours / the reference's CUDA kernel / the fp32 oracle on the same inputs.  Returns (entry, reference kernel inside tolerance?)."""
    import numpy as np
    bound = atol + rtol * np.abs(ref32)
    e = {
        "ours_vs_fp32_max_abs": float(np.abs(ours - ref32).max()),
        "refgpu_vs_fp32_max_abs": float(np.abs(refgpu - ref32).max()),
        "ours_vs_refgpu_max_abs": float(np.abs(ours - refgpu).max()),
        "ours_worst_err_over_bound": float((np.abs(ours - ref32) / bound).max()),
        "refgpu_worst_err_over_bound": float((np.abs(refgpu - ref32) / bound).max()),
        "tolerance": f"|x - ref| <= {atol} + {rtol} |ref|", "note": note,
    }
    ok = bool(e["refgpu_worst_err_over_bound"] <= 1.0)
    e["reference_kernel_inside_tolerance"] = ok
    record(name, e)
    return e, ok
