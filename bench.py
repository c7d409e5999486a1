#!/usr/bin/env python
"""bench.py — measures the flash-attention hot path on B200 through the C ABI (include/b200fa.h).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--quick]

A "step" is one pass of the hot path over one batch of synthetic input.  The headline workload, at EVERY N, is
BASELINE.json configs[3] — Llama-3-8B GQA decode: 32 q / 8 kv heads, d=128, batch 64, KV 8192 f16 with mask — a fixed
2.1 GB problem partitioned by KV head over the N ranks (GQA groups stay together, no collective): STRONG scaling,
value = the whole problem's K/V bytes / max-over-ranks step time.  `workloads` carries the other BASELINE configs, each with its
own roofline and parity block: C2 (LLaMA-7B decode KV 4096, the reference's own call shape), the 32K-KV decode shape north_star
quotes, C3 (2Kx2K causal prefill, TFLOP/s), C5 (128K q8_0 decode, sequence-split over the N ranks), the shapes next to them
(`next_rows`), and — at N = 1 — `ref_gpu`: the reference's OWN CUDA kernels (oracle/_ref/libref_gpu.so, compiled from
/root/reference) timed by the same event loop on the same B200, as context (they are never on the product path).

Timing: CUDA events on the launching stream around CUDA-graph replays of the steps (>= 3 warm-ups), inputs rotated over
buffer sets whose total exceeds the 126 MB L2, barrier + synchronize on both sides, max over ranks.
Parity: every workload compares the CUDA result of one of its timed inputs with the CPU oracle (oracle/, test infrastructure:
used here only as the checker) on sampled units, on every rank; the worst error over the tolerance bound is max-reduced over ranks.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "attn TFLOP/s (prefill) & KV-read GB/s (decode) vs B200 roofline"
L2_BYTES = 126e6
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
ATOL, RTOL = 2e-3, 1e-2  # north_star tolerance
D = 128
C4 = dict(Hq=32, Hk=8, B=64, n_kv=8192)
WORKLOAD = ("c4: Llama-3-8B GQA decode, 32 q / 8 kv heads, d=128, batch 64, KV 8192 f16 + mask, K/V 2147483648 B per step, "
            "head-sharded over the GPUs (BASELINE.json configs[3])")
GROUP_BYTES = 2 * C4["n_kv"] * D * 2  # K + V of one (sequence, kv head) unit: what its 4 q heads read


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return dict(FALLBACK_PEAKS), "fallback (B200_PROFILING.md)"


def csrc_sha() -> str:
    """Identity of the kernel sources: ncu traffic figures committed under profiles/ are only quoted for the build they were taken from."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "ggml-cuda-experiments_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(f.encode()); h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(key: str):
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p))
    return t.get(key) if t.get("csrc_sha") == csrc_sha() else None  # stale capture: say nothing rather than something old


# --------------------------------------------------------------------------------------------------
# the reference's own host attention (utils.h through oracle/_ref, else the oracle port) on C4 units
# --------------------------------------------------------------------------------------------------
class HostC4:
    """Times `groups` (sequence, kv head) units of C4 — 4 q heads against 8192 keys each — on the host cores."""

    def __init__(self):
        import numpy as np
        import oracle
        self.np, self.oracle = np, oracle
        self.cores = os.cpu_count() or 1
        self.kind = "reference" if oracle.ref_host_available() else "port"

    def make(self, groups, seed0=0):
        np, o = self.np, self.oracle
        n_kv = C4["n_kv"]
        Q = o.uniform_pm1(1 + seed0, (groups * 4, 1, D))
        K = o.uniform_pm1(2 + seed0, (groups, n_kv, D)).astype(np.float16)
        V = o.uniform_pm1(3 + seed0, (groups, n_kv, D)).astype(np.float16)
        return Q, K, V

    def run(self, Q, K, V):
        """Q f32 [4g][1][D], K/V f16 [g][n_kv][D] -> f32 [4g][D]; returns (out, threads used)."""
        np, o = self.np, self.oracle
        g, n_kv = K.shape[0], K.shape[1]
        threads = max(1, min(self.cores, 4 * g))
        scale = 1.0 / np.sqrt(D)
        if self.kind == "reference":
            VT = np.ascontiguousarray(V.transpose(0, 2, 1))
            mask = np.zeros((1, n_kv), np.float16)
            out = np.zeros((1, 4 * g, D), np.float32); scores = np.zeros((4 * g, 1, n_kv), np.float32)
            rc = o.ref_host().ref_host_attention_llama(Q.ctypes.data, K.ctypes.data, VT.ctypes.data, mask.ctypes.data, out.ctypes.data,
                                                       scores.ctypes.data, D, 1, n_kv, 4 * g, g, C.c_float(scale), threads)
            assert rc == 0
            return out[0], threads
        mask = np.zeros((1, n_kv), np.float16)
        out = o.flash_attn_ext(o.view_of(Q[None]), o.view_of(K[None]), o.view_of(V[None]), o.view_of(mask), scale, strict_ref=True, nthreads=threads)
        return out[0, 0], threads

    def timed(self, steps, warmup, budget_s):
        """`steps` timed passes over a sample sized to the budget; returns (seconds per step, groups per step, threads)."""
        g0 = max(1, (self.cores + 3) // 4)
        Q, K, V = self.make(g0)
        t0 = time.perf_counter(); self.run(Q, K, V); t1 = max(time.perf_counter() - t0, 1e-4)
        per_step = budget_s / max(steps + warmup, 1)
        if per_step >= t1:
            groups = min(C4["B"] * C4["Hk"], g0 * max(1, int(per_step / t1)))
        else:
            groups = max(1, int(g0 * per_step / t1))
        if groups != g0:
            Q, K, V = self.make(groups)
        threads = 1
        for _ in range(warmup):
            self.run(Q, K, V)
        t0 = time.perf_counter()
        for _ in range(steps):
            _, threads = self.run(Q, K, V)
        return (time.perf_counter() - t0) / max(steps, 1), groups, threads


def run_reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    host = HostC4()
    dt, groups, threads = host.timed(args.steps, args.warmup, budget_s=100.0)
    gbs = groups * GROUP_BYTES / dt / 1e9
    sample = (f"{groups} of {C4['B'] * C4['Hk']} (sequence, kv head) units per step (4 q heads x 8192 keys each), {args.steps} steps; "
              f"{'the reference utils.h mulmat_cpu/softmax sequenced as flash-matrix.cu:88-102, heads over std::threads' if host.kind == 'reference' else 'oracle port of utils.h'}")
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": host.kind, "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------
# helpers for our arm
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={dev}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        rows = [r.strip().split(", ") for r in self.f.read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, load = [], [], set(), []
        for r in rows:
            try:
                clk, cmax, pw, util = float(r[1]), float(r[2]), float(r[3]), float(r[4])
            except Exception:  # noqa: BLE001
                continue
            sm.append(clk); mx.append(cmax)
            if util >= 50 or pw >= 300:
                load.append(clk)
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[6:10]):
                if val.strip().lower().startswith("active"):
                    reasons.add(name)
        use = load or sm
        return {"sm_mhz": statistics.median(use) if use else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(load)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--quick", action="store_true", help="headline only: skip the extra workloads, the reference kernels and the CPU baseline")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    # Libraries (NCCL's version banner, torch warnings) may write to stdout; the contract is ONE JSON line there.
    # Everything but that line goes to stderr: fd 1 is pointed at fd 2 for the run and restored for the final print.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist
    import oracle  # the CHECKER of the parity blocks and the cpu_baseline leg; nothing timed on the GPU touches it
    from __graft_entry__ import load_package
    P = load_package()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — b200fa has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peaks, peak_src = load_peaks()
    warmup = max(args.warmup, 3)
    P.lib()  # fail loudly if the CUDA library is missing
    ncores = os.cpu_count() or 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def rand_f16(shape, seed, of_rank=None):
        g = torch.Generator(device=dev); g.manual_seed(seed + 1000 * (rank if of_rank is None else of_rank))
        return (torch.rand(shape, generator=g, device=dev, dtype=torch.float32) * 2 - 1).to(torch.float16)

    def rand_f32(shape, seed):
        g = torch.Generator(device=dev); g.manual_seed(seed + 1000 * rank)
        return torch.rand(shape, generator=g, device=dev, dtype=torch.float32) * 2 - 1

    def time_steps(step_fn, n_steps, n_warm, chunk=500):
        """Times exactly n_steps calls of step_fn(i) as CUDA-graph replays; returns (ms_per_step_local, max over ranks)."""
        for i in range(n_warm):
            step_fn(i)
        torch.cuda.synchronize()
        chunk = max(1, min(chunk, n_steps))
        reps, rem = n_steps // chunk, n_steps % chunk

        def capture(count):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(count):
                    step_fn(i)
            return g

        g_main = capture(chunk)
        g_rem = capture(rem) if rem else None
        g_main.replay(); torch.cuda.synchronize()  # one untimed replay (graph upload)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(reps):
            g_main.replay()
        if g_rem is not None:
            g_rem.replay()
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ms = e0.elapsed_time(e1) / n_steps
        return ms, max_over_ranks(ms)

    def parity_block(pairs, checked):
        """pairs: (got, ref) numpy arrays of sampled rows.  Errors against the north_star bound, max-reduced over the ranks."""
        max_abs = max_rel = worst = 0.0
        finite = True
        for got, ref in pairs:
            got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
            finite = finite and bool(np.isfinite(got).all())
            err = np.abs(np.nan_to_num(got, nan=1e30) - ref)
            max_abs = max(max_abs, float(err.max()))
            max_rel = max(max_rel, float((err / np.maximum(np.abs(ref), 1e-3)).max()))
            worst = max(worst, float((err / (ATOL + RTOL * np.abs(ref))).max()))
        if world > 1:
            t = torch.tensor([max_abs, max_rel, worst, 0.0 if finite else 1.0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            max_abs, max_rel, worst, bad = (float(x) for x in t.tolist()); finite = bad == 0.0
        return {"max_abs": max_abs, "max_rel": max_rel, "worst_err_over_bound": worst, "ok": bool(finite and worst <= 1.0),
                "tolerance": f"|x - ref| <= {ATOL} + {RTOL} |ref|", "checked": checked + (f", on each of {world} ranks" if world > 1 else ""),
                "oracle": "oracle/attn_oracle.c (fp32 restatement of utils.h:5-49)"}

    def oracle_attn(q, k, v, mask_row=None, q8=False):
        """q f32/f16 [h][n_q][D], k/v f16 [hk][n_kv][D] (or q8_0 bytes) numpy -> [n_q][h][D] by the CPU oracle."""
        kt = oracle.TYPE_Q8_0 if q8 else None
        out = oracle.flash_attn_ext(oracle.view_of(q[None]), oracle.view_of(k[None], kt), oracle.view_of(v[None], kt),
                                    oracle.view_of(mask_row) if mask_row is not None else None, 1.0 / np.sqrt(D), round_q_f16=True, nthreads=ncores)
        return out[0]

    def hbm_roofline(kernel, bytes_per_launch, us, traffic_key=None):
        a = bytes_per_launch / (us * 1e-6) / 1e9
        return {"bound": "hbm", "kernel": kernel, "achieved": a, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": a / peaks["hbm_gbs"],
                "traffic": ncu_traffic(traffic_key) if traffic_key else None, "peak_source": peak_src, "kernel_us": us,
                "algorithmic_bytes_per_launch": bytes_per_launch, "frac_of_nominal_8TBs": a / 8000.0}

    results = {}

    # ---------------------------------------------------------------- C4: the headline workload (strong scaling over kv heads)
    Hq, Hk, B, n_kv = C4["Hq"], C4["Hk"], C4["B"], C4["n_kv"]
    hs = P.head_shard(Hq, Hk, rank, world)  # this rank's band of kv heads (+ their 4 q heads each)
    hk_l, hq_l = hs.n_kv_heads, hs.n_q_heads
    per_gpu = 2 * B * hk_l * n_kv * D * 2
    total_bytes = 2 * B * Hk * n_kv * D * 2
    nsets = max(2, int(3 * L2_BYTES // max(per_gpu, 1)) + 1)
    c4_k = [rand_f16((B, hk_l, n_kv, D), 60 + s) for s in range(nsets)]
    c4_v = [rand_f16((B, hk_l, n_kv, D), 80 + s) for s in range(nsets)]
    c4_q = rand_f32((B, hq_l, 1, D), 59)
    c4_mask = torch.zeros((32, n_kv), dtype=torch.float16, device=dev)  # shared by heads and sequences (flash-llama.h:151), row 0 used
    c4_dst = torch.empty((B, 1, hq_l, D), dtype=torch.float32, device=dev)
    c4_ws = P.Workspace(P.workspace_size(P.TYPE_F32, P.TYPE_F16, D, 1, hq_l, B, n_kv, hk_l, B))

    def c4_step(i):  # the workspace is zero-filled once and only b200fa touches it: the zeroed-workspace contract, one launch per step
        P.flash_attn_ext(c4_q, c4_k[i % nsets], c4_v[i % nsets], c4_mask, dst=c4_dst, flags=P.FLAG_WORKSPACE_ZEROED, workspace=c4_ws)

    sampler = ClockSampler(local) if rank == 0 else None
    c4_step(0); torch.cuda.synchronize()
    launches_per_step = P.last_launch_count()
    dispatch = P.last_dispatch()
    ms_local, ms = time_steps(c4_step, args.steps, warmup, chunk=100)
    value = total_bytes / (ms * 1e-3) / 1e9
    roofline = hbm_roofline("fa_decode_stream<128,f16,1>", per_gpu, ms * 1e3, "c4_fa_decode_stream" if world == 1 else None)
    roofline["note"] = "the step IS one launch of this kernel (in-kernel split-KV merge); per-GPU bytes / max-over-ranks step time"

    def c4_parity():
        c4_step(0); torch.cuda.synchronize()
        got_all = c4_dst.cpu().numpy()  # [B][1][hq_l][D]
        pairs = []
        units = [(0, 0), (B // 2 + 1, hk_l - 1), (B - 1, hk_l // 2)]
        for (b, h) in units:
            q = c4_q[b, 4 * h:4 * h + 4].cpu().numpy(); k = c4_k[0][b, h:h + 1].cpu().numpy(); v = c4_v[0][b, h:h + 1].cpu().numpy()
            ref = oracle_attn(q, k, v, np.zeros((1, n_kv), np.float16))  # [1][4][D]
            pairs.append((got_all[b, 0, 4 * h:4 * h + 4], ref[0]))
        return parity_block(pairs, f"{len(units)} (sequence, kv head) units = {4 * len(units)} q heads x 8192 keys of K/V set 0")

    parity = c4_parity()

    # ---------------------------------------------------------------- e2e: host buffers through the ABI, copies inside the timed region
    def measure_e2e(n_steps=5):
        hk = torch.empty((B, hk_l, n_kv, D), dtype=torch.float16).pin_memory(); hk.copy_(c4_k[0].cpu())
        hv = torch.empty((B, hk_l, n_kv, D), dtype=torch.float16).pin_memory(); hv.copy_(c4_v[0].cpu())
        hq = torch.empty((B, hq_l, 1, D), dtype=torch.float32).pin_memory(); hq.copy_(c4_q.cpu())
        hm = torch.zeros((32, n_kv), dtype=torch.float16).pin_memory()
        ho = torch.empty((B, 1, hq_l, D), dtype=torch.float32).pin_memory()
        dk, dv, dq, dm = c4_k[-1], c4_v[-1], torch.empty_like(c4_q), torch.empty_like(c4_mask)  # device staging = the last rotating set
        def one():
            dk.copy_(hk, non_blocking=True); dv.copy_(hv, non_blocking=True); dq.copy_(hq, non_blocking=True); dm.copy_(hm, non_blocking=True)
            P.flash_attn_ext(dq, dk, dv, dm, dst=c4_dst, workspace=c4_ws)  # the plain public call: no flags
            ho.copy_(c4_dst, non_blocking=True)
        for _ in range(2):
            one()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps):
            one()
        e1.record(); torch.cuda.synchronize(); barrier()
        t = max_over_ranks(e0.elapsed_time(e1) / n_steps)
        h2d = hk.numel() * 2 + hv.numel() * 2 + hq.numel() * 4 + hm.numel() * 2
        return {"value": total_bytes / (t * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": ho.numel() * 4,
                "ms_per_step": t, "note": "per rank: pinned host K/V/Q/mask -> device, b200fa_flash_attn_ext, result -> pinned host; bytes are per rank"}

    e2e = measure_e2e()

    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1 only): the reference's host attention on C4 units
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.quick:
        try:
            host = HostC4()
            dt, groups, threads = host.timed(steps=3, warmup=1, budget_s=12.0)
            # the same host code on units of the GPU's own inputs: reference CPU result vs our CUDA result, directly
            c4_step(0); torch.cuda.synchronize()
            got = c4_dst.cpu().numpy()
            g = min(4, hk_l)
            Qh = np.ascontiguousarray(c4_q[0, :4 * g].cpu().numpy()); Kh = np.ascontiguousarray(c4_k[0][0, :g].cpu().numpy())
            Vh = np.ascontiguousarray(c4_v[0][0, :g].cpu().numpy())
            ref_out, _ = host.run(Qh, Kh, Vh)
            cpu_baseline = {"value": groups * GROUP_BYTES / dt / 1e9, "unit": "GB/s", "cores": threads, "kind": host.kind,
                            "sample": f"{groups} of {B * Hk} (sequence, kv head) units (4 q heads x 8192 keys each) x 3 passes, {dt * 1e3:.1f} ms per pass; "
                                      "reference utils.h host attention sequenced as flash-matrix.cu:88-102, heads over std::threads",
                            "max_abs_diff_gpu_vs_cpu": float(np.abs(got[0, 0, :4 * g] - ref_out).max())}
        except Exception as e:  # noqa: BLE001
            cpu_baseline = {"error": repr(e)}
    del c4_k, c4_v
    torch.cuda.empty_cache()

    # ---------------------------------------------------------------- extra workloads (reported under "workloads")
    def run_c2():
        H, n = 32, 4096
        ns = 6
        ks = [rand_f16((1, n, H, D), 10 + s).permute(0, 2, 1, 3) for s in range(ns)]  # KV-cache view [kv][head][d]
        vs = [rand_f16((1, n, H, D), 20 + s).permute(0, 2, 1, 3) for s in range(ns)]
        q = rand_f32((1, 1, H, D), 9).permute(0, 2, 1, 3)                               # f32 [q][head][d] view
        mask = torch.zeros((32, n), dtype=torch.float16, device=dev)
        dst = torch.empty((1, 1, H, D), dtype=torch.float32, device=dev)
        ws = P.Workspace(P.workspace_size(P.TYPE_F32, P.TYPE_F16, D, 1, H, 1, n, H, 1))
        nbytes = 2 * H * n * D * 2
        def step(i):
            P.flash_attn_ext(q, ks[i % ns], vs[i % ns], mask, dst=dst, flags=P.FLAG_WORKSPACE_ZEROED, workspace=ws)
        step(0); torch.cuda.synchronize()
        nl, disp = P.last_launch_count(), P.last_dispatch()
        got = dst.cpu().numpy()[0, 0]
        heads = [0, 13, 31]
        pairs = [(got[h], oracle_attn(q[0, h:h + 1].cpu().numpy(), ks[0][0, h:h + 1].cpu().numpy(), vs[0][0, h:h + 1].cpu().numpy(),
                                      np.zeros((1, n), np.float16))[0, 0]) for h in heads]
        _, t = time_steps(step, 2000, 10)
        return {"config": "c2: LLaMA-7B decode, 32 heads, d=128, batch 1, KV 4096 f16 + mask, KV-cache view strides (BASELINE.json configs[1]); each rank its own 32 heads",
                "gbps": nbytes / (t * 1e-3) / 1e9, "us_per_step": t * 1e3, "launches_per_step": nl, "dispatch": disp,
                "roofline": hbm_roofline("fa_decode_stream<128,f16,1>", nbytes, t * 1e3, "c2_fa_decode_stream"),
                "parity": parity_block(pairs, f"heads {heads} of K/V set 0"), "l2": f"{ns} rotating K/V sets of {nbytes / 1e6:.0f} MB"}

    def run_c2_32k():
        H, n = 32, 32768
        ns = 2
        ks = [rand_f16((1, n, H, D), 110 + s).permute(0, 2, 1, 3) for s in range(ns)]
        vs = [rand_f16((1, n, H, D), 120 + s).permute(0, 2, 1, 3) for s in range(ns)]
        q = rand_f32((1, 1, H, D), 109).permute(0, 2, 1, 3)
        mask = torch.zeros((32, n), dtype=torch.float16, device=dev)
        dst = torch.empty((1, 1, H, D), dtype=torch.float32, device=dev)
        ws = P.Workspace(P.workspace_size(P.TYPE_F32, P.TYPE_F16, D, 1, H, 1, n, H, 1))
        nbytes = 2 * H * n * D * 2
        def step(i):
            P.flash_attn_ext(q, ks[i % ns], vs[i % ns], mask, dst=dst, flags=P.FLAG_WORKSPACE_ZEROED, workspace=ws)
        step(0); torch.cuda.synchronize()
        nl, disp = P.last_launch_count(), P.last_dispatch()
        got = dst.cpu().numpy()[0, 0]
        heads = [5, 30]
        pairs = [(got[h], oracle_attn(q[0, h:h + 1].cpu().numpy(), ks[0][0, h:h + 1].cpu().numpy(), vs[0][0, h:h + 1].cpu().numpy(),
                                      np.zeros((1, n), np.float16))[0, 0]) for h in heads]
        _, t = time_steps(step, 200, 5, chunk=50)
        return {"config": "LLaMA-7B decode at 32K KV: 32 heads, d=128, batch 1, f16 + mask, KV-cache view strides (north_star's >= 80 % HBM target shape)",
                "gbps": nbytes / (t * 1e-3) / 1e9, "us_per_step": t * 1e3, "kv_bytes": nbytes, "launches_per_step": nl, "dispatch": disp,
                "roofline": hbm_roofline("fa_decode_stream<128,f16,1>", nbytes, t * 1e3),
                "parity": parity_block(pairs, f"heads {heads} of K/V set 0"), "l2": f"{ns} rotating K/V sets of {nbytes / 1e6:.0f} MB"}

    def run_c3():
        n, H = 2048, 32
        ns = 4  # 4 x 48 MB of Q/K/V > 126 MB L2
        qs = [rand_f16((1, H, n, D), 30 + s) for s in range(ns)]
        ks = [rand_f16((1, H, n, D), 40 + s) for s in range(ns)]
        vs = [rand_f16((1, H, n, D), 50 + s) for s in range(ns)]
        mask = torch.full((n, n), float("-inf"), dtype=torch.float16, device=dev).triu(1)
        dst = torch.empty((1, n, H, D), dtype=torch.float32, device=dev)
        ws = P.Workspace(P.workspace_size(P.TYPE_F16, P.TYPE_F16, D, n, H, 1, n, H, 1))
        flops = 4 * H * n * n * D / 2
        out = {}
        # mask_tensor_only is the reference's actual call (flash-llama.h:6-32 has no causal flag); causal_flag tells the kernel what the mask is
        for name, flags, fl in (("mask_tensor_only", P.FLAG_WORKSPACE_ZEROED, flops), ("causal_flag", P.FLAG_CAUSAL | P.FLAG_WORKSPACE_ZEROED, flops)):
            def step(i):
                P.flash_attn_ext(qs[i % ns], ks[i % ns], vs[i % ns], mask, dst=dst, flags=flags, workspace=ws)
            step(0); torch.cuda.synchronize()
            nl, disp = P.last_launch_count(), P.last_dispatch()
            got = dst.cpu().numpy()[0]  # [n][H][D]
            heads = [0, 9, 22, 31]     # four WHOLE heads, every row
            mk = mask.cpu().numpy()
            pairs = [(got[:, h], oracle_attn(qs[0][0, h:h + 1].cpu().numpy(), ks[0][0, h:h + 1].cpu().numpy(), vs[0][0, h:h + 1].cpu().numpy(), mk)[:, 0])
                     for h in heads]
            _, t = time_steps(step, 400, 10, chunk=50)
            tf = fl / (t * 1e-3) / 1e12
            out[name] = {"tflops": tf, "us_per_step": t * 1e3, "launches_per_step": nl, "dispatch": disp,
                         "roofline": {"bound": "tensor", "kernel": "fa_prefill_persistent", "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                      "frac": tf / peaks["bf16_tflops"], "traffic": ncu_traffic("c3_fa_prefill_persistent") if name == "causal_flag" else None, "peak_source": peak_src + " burst", "frac_of_nominal_2250": tf / 2250.0,
                                      "algorithmic_flops_per_launch": fl},
                         "parity": parity_block(pairs, f"all 2048 rows of heads {heads} of input set 0")}
        out["config"] = "c3: LLaMA-7B prefill, 32 heads, d=128, 2048x2048 causal f16 Q/K/V, f32 out (BASELINE.json configs[2]); causal FLOPs 34.36 G; each rank its own copy"
        return out

    def run_c5():
        Hq5, Hk5, n5 = 32, 8, 131072
        ss = P.seq_shard(n5, rank, world)
        n_local = ss.n_local
        ns = max(2, int(3 * L2_BYTES // (2 * Hk5 * n_local * 136)) + 1)
        def make_set(s, of_rank=None, n_keys=n_local):
            kf = rand_f16((1, Hk5, n_keys, D), 70 + s, of_rank); vf = rand_f16((1, Hk5, n_keys, D), 90 + s, of_rank)
            return P.quantize_q8_0(kf), P.quantize_q8_0(vf)
        ksets, vsets = zip(*[make_set(s) for s in range(ns)])
        g = torch.Generator(device=dev); g.manual_seed(69)  # the SAME query on every rank
        q = torch.rand((1, Hq5, 1, D), generator=g, device=dev, dtype=torch.float32) * 2 - 1
        rows = Hq5
        part = torch.empty((rows, D + 2), dtype=torch.float32, device=dev)
        gathered = torch.empty((world, rows, D + 2), dtype=torch.float32, device=dev)
        dst = torch.empty((rows, D), dtype=torch.float32, device=dev)
        ws = P.Workspace(P.workspace_size(P.TYPE_F32, P.TYPE_Q8_0, D, 1, Hq5, 1, n_local, Hk5, 1))
        def local_step(i):
            P.flash_attn_partial(q, ksets[i % ns], vsets[i % ns], kv_pos0=ss.kv_pos0, n_kv_total=n5, workspace=ws, out=part, flags=P.FLAG_WORKSPACE_ZEROED)
        def full_step(i):
            local_step(i)
            if world > 1:
                dist.all_gather_into_tensor(gathered.view(world * rows, D + 2), part)
                P.merge_partials(gathered, dst=dst)
            else:
                P.merge_partials(part.view(1, rows, D + 2), dst=dst)
        per = 2 * Hk5 * n_local * (D // 32 * 34)
        local_step(0); torch.cuda.synchronize()
        nl, disp = P.last_launch_count(), P.last_dispatch()
        # parity of the FINAL output on every rank: the oracle sees the whole 131072-key K/V of two kv heads (8 q heads), rebuilt
        # from every rank's seeded shard generator
        full_step(0); torch.cuda.synchronize()
        got = dst.cpu().numpy()
        pairs = []
        for h in (1, 6):
            kparts, vparts = [], []
            for r in range(world):
                sr = P.seq_shard(n5, r, world)
                kq, vq = make_set(0, of_rank=r, n_keys=sr.n_local)
                kparts.append(kq[0, h].cpu().numpy()); vparts.append(vq[0, h].cpu().numpy()); del kq, vq
            kh = np.concatenate(kparts, 0)[None]; vh = np.concatenate(vparts, 0)[None]
            ref = oracle_attn(q[0, 4 * h:4 * h + 4].cpu().numpy(), kh, vh, None, q8=True)[0]
            pairs.append((got[4 * h:4 * h + 4], ref))
        par = parity_block(pairs, "final merged output of kv heads 1 and 6 (8 q heads) over all 131072 q8_0 keys of K/V set 0 (oracle: exact f32 d*q dequant)")
        _, t_local = time_steps(local_step, 100, 5, chunk=20)
        for i in range(5):
            full_step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(60):
            full_step(i)
        e1.record(); torch.cuda.synchronize(); barrier()
        t_full = max_over_ranks(e0.elapsed_time(e1) / 60)
        # the same step with the combine over peer-mapped memory (NVLink stores + device-side wait) instead of NCCL
        t_peer = None; t_fused = None; par_fused = None
        if world > 1:
            try:
                xch = P.PeerExchange.distributed(rows, D)
                barrier()
                def peer_step(i):
                    P.flash_attn_partial_scatter(q, ksets[i % ns], vsets[i % ns], xch, kv_pos0=ss.kv_pos0, n_kv_total=n5, workspace=ws, flags=P.FLAG_WORKSPACE_ZEROED)
                    P.merge_partials_wait(xch, dst=dst)
                _, t_peer = time_steps(peer_step, 120, 5, chunk=20)  # the step number lives on the device: graph-replayable
                def fused_step(i):
                    P.flash_attn_seqpar(q, ksets[i % ns], vsets[i % ns], xch, kv_pos0=ss.kv_pos0, n_kv_total=n5, workspace=ws, dst=dst, flags=P.FLAG_WORKSPACE_ZEROED)
                dst.zero_(); barrier()
                fused_step(0); torch.cuda.synchronize(); barrier()
                gf = dst.cpu().numpy()
                par_fused = parity_block([(gf[4 * h:4 * h + 4], ref_h) for h, (_, ref_h) in zip((1, 6), pairs)], "fused one-kernel step, same heads")
                _, t_fused = time_steps(fused_step, 120, 5, chunk=20)
                xch.close()
            except Exception as e:  # noqa: BLE001
                t_peer = repr(e)
        best = min([x for x in (t_full, t_peer, t_fused) if isinstance(x, float)])
        total5 = 2 * Hk5 * n5 * (D // 32 * 34)
        return {"config": f"c5: Llama-3-8B decode, 32 q / 8 kv heads, KV 131072 q8_0 (34 B / 32 elems), sequence-split over {world} GPU(s) "
                          f"(BASELINE.json configs[4]); K/V {total5} B in total",
                "scaling": "strong", "gbps_total_best_step": total5 / (best * 1e-3) / 1e9, "us_best_step": best * 1e3,
                "stream_us": t_local * 1e3, "stream_frac_of_measured_hbm": per / (t_local * 1e-3) / 1e9 / peaks["hbm_gbs"],
                "end_to_end_us_nccl_all_gather_merge": t_full * 1e3,
                "end_to_end_us_peer_memory_combine": (t_peer * 1e3 if isinstance(t_peer, float) else t_peer),
                "end_to_end_us_fused_one_kernel": (t_fused * 1e3 if isinstance(t_fused, float) else t_fused),
                "roofline": hbm_roofline("fa_decode_stream<128,q8_0,1,T8> (local stream, partial triples out)", per, t_local * 1e3,
                                         "c5_fa_decode_stream" if world == 1 else None),
                "parity": par, "parity_fused": par_fused, "dispatch": disp,
                "kv_bytes_per_gpu": per, "combine_payload_bytes_per_rank": rows * (D + 2) * 4, "launches_local": nl,
                "l2": f"{ns} rotating q8_0 K/V sets of {per / 1e6:.0f} MB per GPU"}

    def run_next_rows():
        """Shapes either side of the configs (SURVEY.md §8f row 3 and the dispatch boundaries): chunked prefill against a long
        cache (split-KV prefill), prefill on a q8_0 cache, a speculative-decoding burst under GQA (virtual KV heads)."""
        out = {}
        def one(name, n_q, n_k, Hq_, Hk_, B_, causal, q8, note):
            ns = 2
            ks = [rand_f16((B_, Hk_, n_k, D), 200 + s) for s in range(ns)]
            vs = [rand_f16((B_, Hk_, n_k, D), 210 + s) for s in range(ns)]
            if q8:
                ks = [P.quantize_q8_0(k) for k in ks]; vs = [P.quantize_q8_0(v) for v in vs]
            q = rand_f16((B_, Hq_, n_q, D), 220)
            dst = torch.empty((B_, n_q, Hq_, D), dtype=torch.float32, device=dev)
            ws = P.Workspace(P.workspace_size(P.TYPE_F16, P.TYPE_Q8_0 if q8 else P.TYPE_F16, D, n_q, Hq_, B_, n_k, Hk_, B_))
            flags = (P.FLAG_CAUSAL if causal else 0) | P.FLAG_WORKSPACE_ZEROED
            def step(i):
                P.flash_attn_ext(q, ks[i % ns], vs[i % ns], None, dst=dst, flags=flags, workspace=ws)
            step(0); torch.cuda.synchronize()
            nl, disp = P.last_launch_count(), P.last_dispatch()
            # parity: one kv head of the last sequence with its whole GQA group (the flag-only causal mask is rebuilt for the oracle)
            gq = Hq_ // Hk_
            got = dst.cpu().numpy()[B_ - 1][:, (Hk_ - 1) * gq:Hk_ * gq]
            mk = None
            if causal:
                mk = np.zeros((n_q, n_k), np.float16)
                for i in range(n_q):
                    mk[i, i + (n_k - n_q) + 1:] = -np.inf
            ref = oracle_attn(q[B_ - 1, (Hk_ - 1) * gq:].cpu().numpy(), ks[0][B_ - 1, Hk_ - 1:].cpu().numpy(), vs[0][B_ - 1, Hk_ - 1:].cpu().numpy(), mk, q8=q8)
            _, t = time_steps(step, 40, 4, chunk=10)
            fl = 4.0 * B_ * Hq_ * n_q * n_k * D * (0.5 if (causal and n_q == n_k) else 1.0)
            kvb = 2.0 * B_ * Hk_ * n_k * (136 if q8 else 256)
            out[name] = {"config": note, "us_per_step": t * 1e3, "tflops": fl / (t * 1e-3) / 1e12, "kv_gbps": kvb / (t * 1e-3) / 1e9,
                         "dispatch": disp, "launches_per_step": nl, "parity": parity_block([(got, ref)], "last kv head of the last sequence, its whole GQA group, every row")}
        one("chunked_prefill_256x32k", 256, 32768, 32, 32, 1, False, False, "256 new queries against a 32K f16 cache, 32 heads (split-KV prefill)")
        one("prefill_2k_q8_0_cache", 2048, 2048, 32, 32, 1, True, True, "C3's shape with q8_0 K/V (dequantised once to f16 workspace copies)")
        one("burst_8x_gqa4_b8_kv8192", 8, 8192, 32, 8, 8, False, False, "8 query positions x GQA 4 = 32 rows per KV head, batch 8, KV 8192 f16 (the group packed into one 128-row tile per KV head, 2 KV segments)")
        return out

    def run_ref_gpu():
        """The reference's own CUDA kernels on this B200 (context only; compiled from /root/reference into oracle/_ref/libref_gpu.so)."""
        path = oracle.ref_gpu_path()
        if not os.path.exists(path):
            return {"unavailable": "oracle/_ref/libref_gpu.so not built (needs /root/reference at build time)"}
        lib = C.CDLL(path)
        vp, ci = C.c_void_p, C.c_int
        lib.ref_gpu_flash_attn_row.restype = ci
        lib.ref_gpu_flash_attn_row.argtypes = [vp] * 6 + [ci, C.c_float, ci, ci, vp]
        lib.ref_gpu_row_tmp_halves.restype = C.c_size_t
        lib.ref_gpu_row_tmp_halves.argtypes = [ci, ci]
        lib.ref_gpu_flash_attn_ext_f16.restype = ci
        lib.ref_gpu_flash_attn_ext_f16.argtypes = [vp] * 5 + [C.c_float] + [ci] * 20 + [vp]
        out = {"note": "reference kernels run as legacy HMMA (f16 accumulate) code on sm_100a; their host-side V^T / dense-K repack "
                       "(flash-matrix.cu:130-165) is done outside the timed region"}
        scale = 1.0 / np.sqrt(D)
        # --- C2 through flash_attn_row<128,8,2,256> + fa_reduce<128,8> (flash-matrix.cu:210-227)
        H, n = 32, 4096
        ns = 6
        ks = [rand_f16((H, n, D), 310 + s) for s in range(ns)]                            # dense [head][kv][D]
        vts = [rand_f16((H, n, D), 320 + s).transpose(1, 2).contiguous() for s in range(ns)]  # V^T [head][D][kv]
        q = rand_f32((H, D), 309)
        mask1 = torch.zeros((n,), dtype=torch.float16, device=dev)
        tmp = torch.empty((lib.ref_gpu_row_tmp_halves(n, H),), dtype=torch.float16, device=dev)
        dst = torch.empty((H, D), dtype=torch.float32, device=dev)
        def row_step(i):
            rc = lib.ref_gpu_flash_attn_row(q.data_ptr(), ks[i % ns].data_ptr(), vts[i % ns].data_ptr(), mask1.data_ptr(), tmp.data_ptr(), dst.data_ptr(),
                                            n, C.c_float(scale), H, 1, torch.cuda.current_stream().cuda_stream)
            assert rc == 0
        row_step(0); torch.cuda.synchronize()
        got = dst.cpu().numpy()
        hsel = [0, 17]
        pr = parity_block([(got[h], oracle_attn(q[h:h + 1, None].cpu().numpy(), ks[0][h:h + 1].cpu().numpy(),
                                                vts[0][h:h + 1].transpose(1, 2).contiguous().cpu().numpy(), np.zeros((1, n), np.float16))[0, 0]) for h in hsel],
                          f"heads {hsel}; the reference accumulates in f16")
        _, t = time_steps(row_step, 1000, 10)
        nbytes = 2 * H * n * D * 2
        out["c2_flash_attn_row_plus_fa_reduce"] = {"kernels": "flash_attn_row<128,8,2,256> + fa_reduce<128,8> (flash_row_float.h:4-200,415-472)",
                                                   "us_per_step": t * 1e3, "gbps": nbytes / (t * 1e-3) / 1e9, "frac_of_measured_hbm": nbytes / (t * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                                   "launches_per_step": 2, "parity_vs_oracle": pr}
        del ks, vts
        # --- 2K x 2K NON-causal prefill through flash_attn_ext_f16<128,16,128> (flash-llama.h:5-438, launch flash-matrix.cu:198-206), and ours on the same inputs
        n, H = 2048, 32
        q3 = rand_f32((1, H, n, D), 330); k3 = rand_f16((1, H, n, D), 331); v3 = rand_f16((1, H, n, D), 332)
        mask0 = torch.zeros((n, n), dtype=torch.float16, device=dev)
        dst3 = torch.empty((1, n, H, D), dtype=torch.float32, device=dev)
        def ext_step(i):
            rc = lib.ref_gpu_flash_attn_ext_f16(q3.data_ptr(), k3.data_ptr(), v3.data_ptr(), mask0.data_ptr(), dst3.data_ptr(), C.c_float(scale),
                                                D, n, H, 1, D, n, H, 1, n, n * 2, D * 4, n * D * 4, H * n * D * 4, D * 2, n * D * 2, H * n * D * 2,
                                                D, H, n, 1, torch.cuda.current_stream().cuda_stream)
            assert rc == 0
        ext_step(0); torch.cuda.synchronize()
        got = dst3.cpu().numpy()[0]
        ref0 = oracle_attn(q3[0, 3:4].cpu().numpy(), k3[0, 3:4].cpu().numpy(), v3[0, 3:4].cpu().numpy(), None)[:, 0]
        pr3 = parity_block([(got[:, 3], ref0)], "all 2048 rows of head 3; the reference accumulates in f16")
        _, t = time_steps(ext_step, 20, 3, chunk=5)
        fl = 4.0 * H * n * n * D
        ws3 = P.Workspace(P.workspace_size(P.TYPE_F32, P.TYPE_F16, D, n, H, 1, n, H, 1))
        def ours_step(i):
            P.flash_attn_ext(q3, k3, v3, mask0, dst=dst3, flags=P.FLAG_WORKSPACE_ZEROED, workspace=ws3)
        ours_step(0); torch.cuda.synchronize()
        got_o = dst3.cpu().numpy()[0]
        pro = parity_block([(got_o[:, 3], ref0)], "all 2048 rows of head 3")
        _, to = time_steps(ours_step, 200, 5, chunk=50)
        out["prefill_2k_noncausal_flash_attn_ext_f16"] = {"kernels": "flash_attn_ext_f16<128,16,128>, 2 warps (flash-llama.h:5-438)", "us_per_step": t * 1e3,
                                                          "tflops": fl / (t * 1e-3) / 1e12, "parity_vs_oracle": pr3,
                                                          "ours_same_inputs": {"us_per_step": to * 1e3, "tflops": fl / (to * 1e-3) / 1e12, "parity": pro,
                                                                               "ours_vs_refgpu_max_abs": float(np.abs(got_o - got).max()), "speedup": t / to}}
        return out

    if not (args.quick or args.no_extras):
        todo = [("c2_decode_kv4096", run_c2), ("c2_decode_32k_kv", run_c2_32k), ("c3_prefill", run_c3), ("c5_q8_0_split_kv", run_c5)]
        if world == 1:
            todo += [("next_rows", run_next_rows), ("ref_gpu", run_ref_gpu)]
        for name, fn in todo:
            try:
                results[name] = fn()
            except Exception as e:  # noqa: BLE001
                results[name] = {"error": repr(e)}
            torch.cuda.empty_cache()
        if world == 1 and isinstance(results.get("ref_gpu"), dict) and "c2_flash_attn_row_plus_fa_reduce" in results["ref_gpu"] and "us_per_step" in results.get("c2_decode_kv4096", {}):
            r = results["ref_gpu"]["c2_flash_attn_row_plus_fa_reduce"]
            r["ours_c2_us_per_step"] = results["c2_decode_kv4096"]["us_per_step"]
            r["speedup"] = r["us_per_step"] / results["c2_decode_kv4096"]["us_per_step"]

    clocks = sampler.stop() if sampler else None

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": f"inputs rotate over {nsets} K/V sets of {per_gpu / 1e6:.0f} MB per GPU (> 126 MB L2)",
                       "parallelism": f"head-parallel x{world}: {hk_l} of 8 kv heads (+ their GQA groups) per rank, no collective",
                       "dispatch": dispatch, "timing": "CUDA events around CUDA-graph replays, max over ranks"},
            "roofline": roofline, "parity": parity, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "workloads": results,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
