#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native faer hot path.

Workload at N=1 (BASELINE.json configs[1]): f64 Cholesky LLT, n = 16384, synthetic SPD input
A = G G^T + n I (G ~ N(0,1), the reference's bench generator, faer/examples/bench.rs:1513-1515), column-major,
resident in HBM when the timed region starts. One "step" = restore the input (device copy of the 2.1 GB matrix;
faer's own bench also times `copy_from_triangular_lower`, bench.rs:1531-1540) + one in-place factorisation through
the C ABI `libfaer_v0_23_llt_factor_in_place_f64`.
  value  = n^3/3 flop per factorisation x N ranks / max-over-ranks device time   [TFLOP/s]
  e2e    = same metric through the same C-ABI call with HOST (pinned) buffers: H2D + factor + D2H inside the timed region
N > 1 (torchrun, one rank per GPU): ONE matrix is factored by all N GPUs — 1-D block-column-cyclic layout, NCCL broadcast of
each factored panel with look-ahead (csrc/dist.cu); weak scaling: n = 16384 * N^(1/3) rounded to whole blocks, so the
per-GPU flop stays at the N=1 value; value = n^3/3 / (max-over-ranks device time).

--impl reference: times the CPU restatement of the reference's algorithm (oracle/, OpenMP over all host cores; faer
itself needs a Rust toolchain that this image does not have) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "f64 GEMM & LU/LLT TFLOP/s at n=16384; % of B200 tensor-core peak"
UNIT = "TFLOP/s"
N_DEFAULT = 16384


def llt_flops(n: int) -> float:
    return n ** 3 / 3.0


# ---------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.time(), line.strip()))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for (ts, line) in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples inside the timed region"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU legs (oracle = test infrastructure; this is one of the two places allowed to execute it)
# ---------------------------------------------------------------------------------------------------
def cpu_llt_sample(target_seconds: float = 15.0, n_cap: int = 12288):
    """Time the CPU restatement of faer's LLT on all host cores on a bounded sample (same generator, smaller n)."""
    from oracle import oracle as orc
    orc.load()
    # The restatement keeps faer's 128-wide recursion, i.e. many small OpenMP regions: beyond ~32 threads the fork/join
    # cost dominated on the GPU box (first version: 0.0007 TFLOP/s with 128 threads vs 0.017 with 64), so the team is
    # capped. Since then the products run cache-blocked on packed panels and the triangular solves fork once per solve
    # (bitwise the same results; 5x faster on 8 cores here), but the cap has not been re-measured on the box.
    cores = max(1, min(os.cpu_count() or 1, 32))
    orc.set_num_threads(cores)
    rng = np.random.default_rng(0)

    def run(n):
        G = rng.standard_normal((n, n))
        A = np.asfortranarray(G @ G.T + n * np.eye(n))
        t = time.perf_counter()
        fail, _ = orc.llt(A)
        dt = time.perf_counter() - t
        assert fail == -1
        return dt

    run(512)  # thread start-up
    t_probe = run(2048)
    rate = llt_flops(2048) / t_probe
    n = int((target_seconds * rate * 3.0) ** (1.0 / 3.0)) // 256 * 256
    n = max(2048, min(n_cap, n))
    dt = run(n)
    return {"value": llt_flops(n) / dt / 1e12, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle (C++/OpenMP restatement of faer's LLT, not faer itself: no Rust toolchain) LLT n={n}, "
                      f"same generator as the GPU workload, {dt:.2f} s"}, n, dt


def lapack_proxy(n: int):
    """scipy/OpenBLAS dpotrf on all cores at the full size — a PROXY for an optimised CPU library, not faer."""
    try:
        import scipy.linalg as sla
        rng = np.random.default_rng(0)
        G = rng.standard_normal((n, n))
        A = np.asfortranarray(G @ G.T + n * np.eye(n))
        t = time.perf_counter()
        sla.cholesky(A, lower=True, overwrite_a=True, check_finite=False)
        dt = time.perf_counter() - t
        return {"value": llt_flops(n) / dt / 1e12, "unit": UNIT, "what": f"scipy.linalg.cholesky (OpenBLAS) n={n}, {dt:.2f} s",
                "cores": os.cpu_count()}
    except Exception as e:  # pragma: no cover
        return {"value": None, "what": f"unavailable: {e}"}


def run_reference_arm(args):
    """--impl reference: the reference's algorithm on the host cores (oracle port; faer cannot be built here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    times = []
    base, n, _ = cpu_llt_sample(target_seconds=6.0, n_cap=8192)
    from oracle import oracle as orc
    rng = np.random.default_rng(1)
    for it in range(args.warmup + args.steps):
        if it >= 1 and it < args.warmup:
            continue  # one warm-up is enough for a CPU loop; keep the whole run within minutes
        G = rng.standard_normal((n, n))
        A = np.asfortranarray(G @ G.T + n * np.eye(n))
        t = time.perf_counter()
        orc.llt(A)
        dt = time.perf_counter() - t
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = llt_flops(n) / (ms * 1e-3) / 1e12
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"f64 Cholesky LLT, bounded CPU sample n={n} of the n={args.n} workload (SPD = G G^T + n I)",
                   "n": n, "n_full": args.n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": base["cores"], "kind": "port",
                         "sample": f"oracle LLT n={n} per step (C++/OpenMP restatement; faer needs Rust, absent here)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
DIST_NB = 1024


def weak_n(world: int, nb: int = DIST_NB) -> int:
    """Weak scaling: per-GPU flop (n^3 / 3 / N) held at the N=1 value => n ~ 16384 * N^(1/3), rounded to whole blocks."""
    if world == 1:
        return N_DEFAULT
    return int(round(N_DEFAULT * world ** (1.0 / 3.0) / nb)) * nb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=0, help="matrix dimension (default: 16384 at N=1, weak-scaled for N>1)")
    ap.add_argument("--nb", type=int, default=DIST_NB, help="block-column width of the distributed layout (N>1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        if not args.n:
            args.n = N_DEFAULT
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    import faer_b200
    from faer_b200 import linalg as la

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = faer_b200.load()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        faer_b200.dist.init_from_torch_distributed()
    lay = faer_b200.dist
    stream = torch.cuda.current_stream()
    lib.faer_b200_set_stream(stream.cuda_stream)

    nb = args.nb
    n = args.n or weak_n(world, nb)
    distributed = world > 1

    # ---- synthetic SPD input: A = G G^T + n I (same G on every rank); each rank keeps its block columns ----
    torch.manual_seed(1234)
    G = torch.randn((n, n), dtype=torch.float64, device=dev)
    if distributed:
        gcols = torch.as_tensor(lay.global_col_indices(n, nb, world, rank), device=dev)
        A0 = (G @ G[gcols, :].T)  # n x local_cols (row-major storage)
        A0[gcols, torch.arange(gcols.numel(), device=dev)] += n
        A0 = A0.T.contiguous().T  # column-major local matrix
    else:
        gcols = None
        A0 = torch.addmm(n * torch.eye(n, dtype=torch.float64, device=dev), G, G.T).T  # symmetric, column-major view
    del G
    A = A0.clone(memory_format=torch.preserve_format)

    def factor():
        if distributed:
            fail, _ = lay.cholesky_in_place(A, n, nb=nb)
            assert fail == -1
        else:
            la.cholesky_in_place(A)

    def step():
        A.copy_(A0)
        factor()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    launches0 = lib.faer_b200_launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    t_wall1 = time.time()
    my_launches = lib.faer_b200_launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    # one factorisation of the (global) n x n matrix per step, whatever the number of ranks
    value = llt_flops(n) / (ms_per_step * 1e-3) / 1e12

    # ---- correctness guard on the timed data (cheap probe): A x == L (L^T x), reduced over the ranks ----
    torch.manual_seed(99)
    x = torch.randn((n, 2), dtype=torch.float64, device=dev)
    if distributed:
        rows = torch.arange(n, device=dev)
        Lloc = torch.where(rows[:, None] >= gcols[None, :], A, torch.zeros((), dtype=torch.float64, device=dev))
        lhs = A0 @ x[gcols, :]
        rhs = Lloc @ (Lloc.T @ x)
        amax = A0.abs().max().reshape(1)
        dist.all_reduce(lhs); dist.all_reduce(rhs); dist.all_reduce(amax, op=dist.ReduceOp.MAX)
        resid = float((lhs - rhs).abs().max()) / (float(amax.item()) * n)
        del Lloc
    else:
        L = torch.tril(A)
        resid = float((A0 @ x - L @ (L.T @ x)).abs().max()) / (float(A0.abs().max()) * n)
        del L

    # ---- roofline of the dominant kernel (the DMMA GEMM doing the trailing updates), rank 0's view ----
    roof = None
    import ctypes as C
    # The profiled step runs the SAME block-column schedule with the two-stream look-ahead switched off: with look-ahead
    # the panel-chain GEMMs share the SMs with the trailing update, and per-launch event times of co-running kernels
    # overlap (their sum exceeds the step), which says nothing about the kernel. Serial launches give its own duration.
    A.copy_(A0)
    lib.faer_b200_profile_begin()
    fail, _ = lay.cholesky_in_place(A, n, nb=(nb if distributed else 256), lookahead=False)
    assert fail == -1
    barrier()
    flops = C.c_double(0); ms = C.c_double(0); cnt = C.c_ulonglong(0)
    lib.faer_b200_profile_end(C.byref(flops), C.byref(ms), C.byref(cnt))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "profiles", "r01_f64_peaks.json")))
    except Exception:
        pass
    peak = peaks.get("dmma_tflops_sustained", 36.9)
    ach = flops.value / (ms.value * 1e-3) / 1e12 if ms.value > 0 else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_gemm_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    bf16 = None
    try:
        bf16 = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained")
    except Exception:
        pass
    roof = {"bound": "tensor", "kernel": "gemm_f64_kernel (DMMA.8x8x4 trailing updates)", "achieved": ach, "peak": peak,
            "unit": "TFLOP/s", "frac": (ach / peak) if ach else None, "traffic": traffic,
            "launches_per_step": int(cnt.value), "ms_in_kernel_per_step": ms.value,
            "flops_in_kernel_per_step": flops.value,
            "peak_source": "measured on this pool: DMMA.8x8x4 issue-bound peak, profiles/r01_f64_peaks.json "
                           "(tcgen05 has no f64 kind; MEASURED_PEAKS.json only has bf16: "
                           f"{bf16} TF/s sustained => frac_of_bf16 = {(ach / bf16) if (ach and bf16) else None})",
            "how": "CUDA events around every launch of the kernel on the launching stream, one extra profiled step right "
                   "after the timed region (rank 0), same block-column schedule with the look-ahead overlap off so that "
                   "launches do not share the SMs; achieved = sum(algorithmic flop per launch) / sum(duration)"}

    # ---- e2e: same metric with HOST (pinned) buffers, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        rows_, cols_ = A0.shape
        hA0 = torch.empty((cols_, rows_), dtype=torch.float64, pin_memory=True)  # storage of the column-major matrix
        hA0.copy_(A0.T)
        hA = torch.empty((cols_, rows_), dtype=torch.float64, pin_memory=True)
        reps = max(2, min(args.steps, 3))
        bytes_h2d = rows_ * cols_ * 8
        if distributed:
            dA = torch.empty((cols_, rows_), dtype=torch.float64, device=dev)

            def e2e_call():
                dA.copy_(hA, non_blocking=True)            # H2D of this rank's block columns
                fail, _ = lay.cholesky_in_place(dA.T, n, nb=nb)
                hA.copy_(dA, non_blocking=True)            # D2H of the factor
                torch.cuda.synchronize()
                assert fail == -1
            how = "faer_b200.dist.cholesky_in_place on this rank's block columns, pinned host <-> device copies timed"
        else:
            hv = hA.numpy().T  # column-major view over pinned memory

            def e2e_call():
                la.cholesky_in_place(hv)  # H2D + factorisation + D2H inside the C-ABI call, synchronous
            # the call streams block columns (width 256) through the factorization: only the part on / below the
            # diagonal blocks crosses PCIe (the strict upper triangle is neither read nor written by LLT)
            bw = int(os.environ.get("FAER_B200_NB", "0")) or 256
            bytes_h2d = sum((n - j0) * min(bw, n - j0) * 8 for j0 in range(0, n, bw))
            how = ("libfaer_v0_23_llt_factor_in_place_f64 on a pinned HOST matrix (wall clock around the synchronous call); "
                   "block columns are uploaded / downloaded on copy streams while the factorization runs")
        hA.copy_(hA0); e2e_call()  # warm-up (pool allocation)
        ts = []
        for _ in range(reps):
            hA.copy_(hA0)
            barrier()
            t0 = time.perf_counter()
            e2e_call()
            ts.append(time.perf_counter() - t0)
        tt = torch.tensor([float(np.mean(ts))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": llt_flops(n) / float(tt.item()) / 1e12, "unit": UNIT,
               "h2d_bytes_per_step": bytes_h2d * world, "d2h_bytes_per_step": bytes_h2d * world,
               "ms_per_step": 1e3 * float(tt.item()), "how": how}
        if not distributed:
            # same probe as above on the factor that came back to the host (not timed)
            Lh = torch.tril(hA.to(dev).T)
            e2e["probe_residual"] = float((A0 @ x - Lh @ (Lh.T @ x)).abs().max()) / (float(A0.abs().max()) * n)
            del Lh
        del hA, hA0

    # ---- the other two numbers BASELINE.json's metric names (f64 GEMM and LU at the same n), device-resident, N = 1 only;
    # informational: `value` stays the LLT of configs[1] ----
    also = None
    if world == 1 and not args.no_e2e:
        def _best_ms(f, reps=2):
            f(); torch.cuda.synchronize(); best = 1e30
            for _ in range(reps):
                a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
                a0.record(stream); f(); a1.record(stream); torch.cuda.synchronize()
                best = min(best, a0.elapsed_time(a1))
            return best
        del A
        torch.manual_seed(4321)
        X = torch.randn((n, n), dtype=torch.float64, device=dev).T
        Y = torch.randn((n, n), dtype=torch.float64, device=dev).T
        Z = torch.empty((n, n), dtype=torch.float64, device=dev).T
        t_gemm = _best_ms(lambda: la.matmul(Z, la.Accum.Replace, X, Y, 1.0))
        del Y, Z
        Xw = X.clone(memory_format=torch.preserve_format)
        pf = torch.zeros(n, dtype=torch.int64, device=dev); pi = torch.zeros(n, dtype=torch.int64, device=dev)
        t_copy = _best_ms(lambda: Xw.copy_(X))

        def _lu():
            Xw.copy_(X)
            la.lu_in_place(Xw, pf, pi)
        t_lu = _best_ms(_lu) - t_copy
        also = {"gemm_f64_tflops": 2.0 * n ** 3 / t_gemm / 1e9, "gemm_ms": t_gemm,
                "lu_f64_tflops": 2.0 * n ** 3 / 3.0 / t_lu / 1e9, "lu_ms": t_lu, "n": n,
                "what": "device-resident f64 GEMM (Replace, alpha = 1) and partial-pivoting LU (u64 indices) at the same n"}
        del X, Xw
        # ---- the other BASELINE.json configs (f32 QR 65536 x 4096, bidiagonalization and c64 GEMM at n = 8192), timed by
        # tools/bench_other_configs.py in a CHILD process: informational, and nothing that happens there (not even an abort)
        # can touch this line's headline fields ----
        try:
            import subprocess
            child = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools",
                                                                 "bench_other_configs.py")],
                                   capture_output=True, text=True, timeout=300)
            last = [ln for ln in child.stdout.strip().splitlines() if ln.startswith("{")]
            also["other_configs"] = json.loads(last[-1]) if last else {"error": f"exit {child.returncode}: {child.stderr[-300:]}"}
        except Exception as e:  # pragma: no cover
            also["other_configs"] = {"error": repr(e)}

    cpu = None
    proxy = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _, _ = cpu_llt_sample()
        proxy = lapack_proxy(min(n, 8192))

    if rank == 0:
        if distributed:
            wl = (f"f64 Cholesky LLT, ONE n={n} matrix factored by {world} GPUs (1-D block-column-cyclic, nb={nb}, NCCL panel "
                  f"broadcast + look-ahead); weak scaling of BASELINE.json configs[1] (n=16384 at N=1): n^3/(3N) flop per GPU held "
                  "constant; SPD = G G^T + n I; step = restore copy + factor")
        else:
            wl = f"f64 Cholesky LLT n={n} (BASELINE.json configs[1]); SPD = G G^T + n I; step = restore copy + factor"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl, "n": n, "layout": "column-major",
                       "parallelism": f"block-column-cyclic x{world}" if distributed else "single GPU",
                       "l2": f"inputs ({A0.numel() * 8 / 1e9:.1f} GB per GPU) exceed the 126 MB L2; no flush needed",
                       "probe_residual": resid},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(my_launches),
            "roofline": roof, "cpu_baseline": cpu, "cpu_lapack_proxy": proxy, "also": also,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        faer_b200.dist.finalize()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
