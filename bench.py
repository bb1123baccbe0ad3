#!/usr/bin/env python
"""bench.py -- headline benchmark of libb200blas (BASELINE.json: DGEMM TFLOP/s at m=n=k=16384).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 16384]

A "step" is one DGEMM C := A*B (NN, alpha=1, beta=0, f64, column-major) of BASELINE.json
configs[1] on synthetic U(-1,1) matrices.
  value : whole-job TFLOP/s with operands resident in HBM, timed with CUDA events on the stream
          the kernel is launched on (inputs 6.4 GB >> 126 MB L2, so no L2 flush is needed).
  e2e   : the same metric through the reference-facing symbol `dgemm_` with pinned HOST buffers --
          the library stages A and B host->device and C device->host inside the timed region.
  N > 1 : one process per GPU (torchrun); the same 16384^3 problem is 2-D tiled over the ranks
          (strong scaling), operands start on rank 0, timed span = distribute + compute + gather.
  --impl reference : the reference's CPU path for the same call -- the CPU BLAS (OpenBLAS) that
          its interposer forwards to (runtime-blas.c:59-69) -- on a bounded k-slice of the workload.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FP64_PEAK_NOMINAL = 148 * 128 * 1.965e9 / 1e12     # 37.2 TFLOP/s: 148 SMs x 128 flop/clk x 1.965 GHz
FP64_PEAK_MEASURED = 37.05                         # DMMA-only register loop, profiles/r01_probe_peaks_b200.txt


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # "under load": the upper half of the samples (idle samples before/after the region drag the median)
        load = [x for x in sm if x > 0.5 * (sm[-1] if sm else 0)]
        return {"sm_mhz": load[len(load) // 2] if load else None,
                "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "power_w_max": max((float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()), default=None),
                "reasons": sorted(reasons), "samples": len(self.rows)}


def verify_rows(torch, A, B, C, n, nrows=128):
    """Parity of the TIMED result at the benchmarked size (outside the timed region): 128 random rows of C x all columns against a
    float64 numpy product of the same operands on the host; bound ||C_R - A_R B||_F <= 2 k eps ||A_R||_F ||B||_F (c = 2).
    torch tensors are row-major, i.e. the column-major operands are their transposes: C_cm[r, :] = A_cm[r, :] B_cm."""
    import numpy as np
    rows = torch.from_numpy(np.random.default_rng(11).choice(n, size=min(nrows, n), replace=False)).to(A.device)
    A_rows = A[:, rows].T.contiguous().cpu().numpy()
    B_cm = B.cpu().numpy().T
    C_rows = C[:, rows].T.contiguous().cpu().numpy()
    ref = A_rows @ B_cm
    err = float(np.linalg.norm(C_rows - ref)); bound = 2.0 * n * 2.0 ** -53 * float(np.linalg.norm(A_rows)) * float(np.linalg.norm(B_cm))
    return {"rows_checked": int(rows.numel()), "against": "float64 numpy panel product on the host", "fro_err": err, "fro_bound_c2": bound,
            "ok": bool(err <= bound), "max_abs_err": float(np.abs(C_rows - ref).max())}


def e2e_pageable(g, n, nA, nB, C_dev, flops, reps=2):
    """dgemm_ on ordinary (pageable) numpy buffers -- what an unmodified program's malloc'd operands are when the tracker did not
    place them (the reference's miss path, runtime-mem.hpp:84-112)."""
    import numpy as np
    pA = np.empty((n, n)); pB = np.empty((n, n)); pC = np.empty((n, n))
    np.copyto(pA, nA); np.copyto(pB, nB)
    g.call("dgemm_", "N", "N", n, n, n, 1.0, pA, n, pB, n, 0.0, pC, n)
    t0 = time.perf_counter()
    for _ in range(reps):
        g.call("dgemm_", "N", "N", n, n, n, 1.0, pA, n, pB, n, 0.0, pC, n)
    dt = (time.perf_counter() - t0) / reps
    d = float(np.abs(pC[:256, :256] - C_dev[:256, :256].cpu().numpy()).max())
    return {"value": flops / dt / 1e12, "unit": "TFLOP/s", "ms_per_step": dt * 1e3, "host_memory": "pageable (numpy.empty)", "steps": reps,
            "max_abs_diff_vs_resident": d}


def e2e_managed_first_touch(g, lib, n, nA, nB, C_dev, flops, reps=2):
    """dgemm_ on tracked managed blocks (what the interposed calloc hands out) that the CPU has just filled: the timed call
    includes the bulk migration to the device (make_resident) and ends with a host read of C."""
    import numpy as np
    nbytes = n * n * 8
    best = None
    d = None
    for _ in range(reps):
        ptrs = [lib.b200blas_malloc_managed(nbytes) for _ in range(3)]
        assert all(ptrs)
        mA, mB, mC = (np.frombuffer((ctypes.c_char * nbytes).from_address(p), dtype=np.float64).reshape(n, n) for p in ptrs)
        np.copyto(mA, nA); np.copyto(mB, nB); mC[:] = 0.0          # first touch on the CPU
        t0 = time.perf_counter()
        g.call("dgemm_", "N", "N", n, n, n, 1.0, g.DevPtr(ptrs[0]), n, g.DevPtr(ptrs[1]), n, 0.0, g.DevPtr(ptrs[2]), n)
        probe = float(mC[0, 0]) + float(mC[n - 1, n - 1])            # the host reads the result (faults two pages back)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        d = float(np.abs(mC[:64, :64] - C_dev[:64, :64].cpu().numpy()).max())
        del mA, mB, mC
        for p_ in ptrs:
            lib.b200blas_free_managed(p_)
    return {"value": flops / best / 1e12, "unit": "TFLOP/s", "ms_per_step": best * 1e3, "host_memory": "tracked managed (b200blas_malloc_managed), CPU-filled",
            "steps": reps, "max_abs_diff_vs_resident": d, "_probe": probe}


def mg_stats(lib):
    buf = (ctypes.c_ulonglong * 5)()
    lib.b200blas_mg_stats(buf)
    return list(buf)


def other_routines_partitioned(g, lib, torch, dev, peaks, out, world):
    """SGEMM 16384^3, ZGEMM 8192^3, DSYRK / DTRSM / DTRMM 16384 through the Fortran symbols with devices=N (bulk mode: operands land,
    then the ordinary kernels), and the Cholesky workload."""
    def timed(fn, reps=3, warm=1):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]
    n = 16384
    A = torch.rand((n, n), dtype=torch.float32, device=dev) * 2 - 1; B = torch.rand((n, n), dtype=torch.float32, device=dev) * 2 - 1
    C = torch.zeros((n, n), dtype=torch.float32, device=dev)
    c0 = mg_stats(lib)[0]
    ms = timed(lambda: g.call("sgemm_", "N", "N", n, n, n, 1.0, A, n, B, n, 0.0, C, n))
    out["sgemm_16384"] = {"tflops": 2.0 * n ** 3 / ms / 1e9, "ms": ms, "devices": world, "partitioned_calls": mg_stats(lib)[0] - c0}
    del A, B, C
    n = 8192
    A = torch.rand((n, n), dtype=torch.complex128, device=dev); B = torch.rand((n, n), dtype=torch.complex128, device=dev)
    C = torch.zeros((n, n), dtype=torch.complex128, device=dev)
    c0 = mg_stats(lib)[0]
    ms = timed(lambda: g.call("zgemm_", "N", "N", n, n, n, 0.7 - 0.9j, A, n, B, n, 1.3 - 1.1j, C, n))
    out["zgemm_8192"] = {"tflops": 8.0 * n ** 3 / ms / 1e9, "ms": ms, "devices": world, "partitioned_calls": mg_stats(lib)[0] - c0,
                         "frac_of_fp64_peak_per_gpu": 8.0 * n ** 3 / ms / 1e9 / world / FP64_PEAK_NOMINAL}
    del A, B, C
    # DSYRK 16384^2 x 16384 and DTRSM / DTRMM 16384^2 through dsyrk_ / dtrsm_ / dtrmm_ with devices=N (csrc/multi_level3.cu)
    n = 16384
    A = torch.rand((n, n), dtype=torch.float64, device=dev) * 2 - 1
    C = torch.zeros((n, n), dtype=torch.float64, device=dev)
    c0 = mg_stats(lib)[0]
    ms = timed(lambda: g.call("dsyrk_", "L", "N", n, n, 1.0, A, n, 0.0, C, n))
    out["dsyrk_LN_16384"] = {"tflops": float(n) ** 3 / ms / 1e9, "ms": ms, "devices": world, "partitioned_calls": mg_stats(lib)[0] - c0,
                             "frac_of_fp64_peak_per_gpu": float(n) ** 3 / ms / 1e9 / world / FP64_PEAK_NOMINAL}
    # check (outside the timed region): 64 sampled columns of the lower triangle against a float64 product of the same operands;
    # torch tensors are row-major, so the column-major operand the library saw is A^T and its C := A_cm A_cm^T lower is torch's upper
    eps = 2.0 ** -53
    cols = torch.randperm(n, device=dev)[:64]
    Acm_cols = A[:, cols]                                    # column-major rows `cols` of the operand = torch columns
    ref = A.T @ Acm_cols                                     # (A_cm A_cm^T)[:, cols]
    got = C[cols, :].T                                       # column-major C[:, cols]
    mask = torch.arange(n, device=dev)[:, None] >= cols[None, :]          # referenced (lower) part of those columns
    err = ((got - ref) * mask).norm().item(); bound = 4 * (n + 2) * eps * (A.norm().item() * Acm_cols.norm().item())
    out["dsyrk_LN_16384"]["verified"] = {"ok": bool(err <= bound), "err_fro_64_columns": err, "bound": bound,
                                         "upper_untouched": bool((torch.tril(C, -1) == 0).all().item())}
    del ref, got, mask, Acm_cols
    T = torch.triu(A).contiguous(); T.mul_(1.0 / n); T.diagonal().fill_(1.0)          # row-major upper == column-major lower, well conditioned
    for name in ("dtrsm_", "dtrmm_"):
        C.uniform_(-1, 1)
        c0 = mg_stats(lib)[0]
        ms = timed(lambda: g.call(name, "L", "L", "N", "N", n, n, 1.0, T, n, C, n))
        out[name + "LLNN_16384"] = {"tflops": float(n) ** 3 / ms / 1e9, "ms": ms, "devices": world, "partitioned_calls": mg_stats(lib)[0] - c0,
                                    "frac_of_fp64_peak_per_gpu": float(n) ** 3 / ms / 1e9 / world / FP64_PEAK_NOMINAL}
        # one more call on known right-hand sides, checked on 64 of them: column-major B[:, j] is torch row j
        C.uniform_(-1, 1); B0 = C[:64].clone()
        g.call(name, "L", "L", "N", "N", n, n, 1.0, T, n, C, n); torch.cuda.synchronize()
        Lcm = T.T                                            # the column-major lower triangle as a torch matrix
        X = C[:64].T
        if name == "dtrsm_":
            err = (Lcm @ X - B0.T).norm().item(); bound = 4 * n * eps * (Lcm.norm().item() * X.norm().item() + B0.norm().item())
        else:
            err = (X - Lcm @ B0.T).norm().item(); bound = 4 * (n + 2) * eps * Lcm.norm().item() * B0.norm().item()
        out[name + "LLNN_16384"]["verified"] = {"ok": bool(err <= bound), "err_fro_64_rhs": err, "bound": bound}
        del B0, X
    del A, C, T
    out.update(cholesky_single_call(lib, torch, dev, world))


def cholesky_single_call(lib, torch, dev, world, n=32768):
    """BASELINE.json configs[3]: the blocked Cholesky workload in one call (b200blas_cholesky_lower, csrc/multi_gemm.cu) on the
    devices selected with devices=<n>; matrix resident on GPU 0, timed span = distribute + factor + gather, wall clock."""
    lib.b200blas_cholesky_lower.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int]
    lib.b200blas_cholesky_lower.restype = ctypes.c_int
    M = torch.rand((n, n), dtype=torch.float64, device=dev) * 2 - 1
    M = torch.tril(M, -1); M = M + M.T; M.diagonal().fill_(float(n))
    W = torch.empty_like(M)
    x = torch.rand(n, dtype=torch.float64, device=dev)
    res = {}
    for nb in ((2048,) if world == 1 else ((1024, 2048) if world == 2 else (256, 512, 1024))):
        best, info = None, 0
        for _ in range(3):
            W.copy_(M); torch.cuda.synchronize()
            t0 = time.perf_counter(); info = lib.b200blas_cholesky_lower(n, ctypes.c_void_p(W.data_ptr()), n, nb); torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        L = torch.triu(W)          # row-major upper == column-major lower
        r = (M @ x - L.T @ (L @ x)).norm().item() / (M.norm().item() * x.norm().item())
        cur = {"tflops": n ** 3 / 3.0 / best / 1e12, "ms": best * 1e3, "info": int(info), "nb": nb, "devices": world, "probe_residual": r,
               "frac_of_fp64_peak_per_gpu": n ** 3 / 3.0 / best / 1e12 / world / FP64_PEAK_NOMINAL,
               "note": "b200blas_cholesky_lower: diagonal-block factorisation + DTRSM panel + masked-GEMM trailing update, look-ahead 1; flops n^3/3"}
        if "cholesky_32768_single_call" not in res or cur["ms"] < res["cholesky_32768_single_call"]["ms"]:
            res["cholesky_32768_single_call"] = cur
    del M, W
    return res


def l1_c_driver(peaks):
    """Level-1 at the symbol boundary from C, not ctypes: tests/drivers/l1_chain.c under LD_PRELOAD=libb200blas.so on calloc'd
    (tracked -> managed) vectors of 2^26 doubles -- BASELINE.json configs[2] as an unmodified program runs it; wall clock per call
    including launch, completion and the scalar's way back."""
    from test_preload import build_driver, fields, run
    hbm = peaks.get("hbm_gbs", 6553.9)
    exe = build_driver("l1_chain")
    out, _ = run(exe, [1 << 26, 40, 1 << 28], preload=True, timeout=600)
    r = dict(kv.split("=", 1) for kv in [l for l in out.splitlines() if l.startswith("GBS")][0].split() if "=" in kv)
    res = {}
    for k, size in (("ddot", "2^26"), ("daxpy", "2^26"), ("dnrm2", "2^26"), ("idamax", "2^28")):
        res[k + "_" + size] = {"gbs": float(r[k]), "frac_of_measured_hbm": float(r[k]) / hbm}
    chk = [l for l in out.splitlines() if l.startswith("RESULT")][0]
    res["how"] = ("C driver (tests/drivers/l1_chain.c) under LD_PRELOAD=libb200blas.so: calloc'd (tracked -> managed) vectors filled by the CPU, "
                  "clock_gettime around each call at the symbol, mean of 39 steady calls; " + chk)
    return res


def traffic_from_profile():
    """dram__bytes_read.sum + dram__bytes_write.sum of the DGEMM kernel from the committed ncu --set full capture."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "dgemm16384_traffic.json")))
    except Exception:
        return None


def other_routines(g, torch, dev, peaks, out):
    """The remaining routines BASELINE.json's metric names, at its config sizes, device-resident, CUDA-event timed
    (median of 5 after 2 warm-ups; every operand set is larger than L2 or rotated so nothing is served from L2)."""
    hbm = peaks.get("hbm_gbs", 6553.9)

    def timed(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    n = 16384
    A = torch.rand((n, n), dtype=torch.float32, device=dev) * 2 - 1; B = torch.rand((n, n), dtype=torch.float32, device=dev) * 2 - 1
    C = torch.zeros((n, n), dtype=torch.float32, device=dev)
    ms = timed(lambda: g.call("sgemm_", "N", "N", n, n, n, 1.0, A, n, B, n, 0.0, C, n), reps=3, warm=1)
    tf = 2.0 * n ** 3 / ms / 1e9
    tf32_peak = peaks.get("bf16_tflops", 1667.5) / 2.0     # tf32 dense = half the bf16 rate; 3 MMAs per product
    out["sgemm_16384"] = {"tflops": tf, "ms": ms, "variant": g.last_variant(), "frac_of_fp32_ffma_peak_74.4": tf / 74.4,
                          "frac_of_tf32_pipe_div3": tf / (tf32_peak / 3.0),
                          "frac_of_tf32_pipe_div3_sustained": (tf / (peaks["bf16_tflops_sustained"] / 2.0 / 3.0)) if peaks.get("bf16_tflops_sustained") else None,
                          "note": "3xTF32 on tcgen05 incl. the split pass; tensor denominator = measured bf16 burst / 2 / 3 (and, second key, the sustained bf16 figure: back-to-back launches run power-limited); cuBLAS SGEMM (FFMA) on this part: 67 TFLOP/s"}
    del A, B, C
    n = 8192
    A = torch.rand((n, n), dtype=torch.complex128, device=dev); B = torch.rand((n, n), dtype=torch.complex128, device=dev)
    C = torch.zeros((n, n), dtype=torch.complex128, device=dev)
    ms = timed(lambda: g.call("zgemm_", "N", "N", n, n, n, 0.7 - 0.9j, A, n, B, n, 1.3 - 1.1j, C, n), reps=3, warm=1)
    out["zgemm_8192"] = {"tflops": 8.0 * n ** 3 / ms / 1e9, "ms": ms, "variant": g.last_variant(), "frac_of_fp64_peak": 8.0 * n ** 3 / ms / 1e9 / FP64_PEAK_NOMINAL}
    del A, B, C
    m = 32768
    A = torch.rand((m, m), dtype=torch.float64, device=dev); x = torch.rand(m, dtype=torch.float64, device=dev); y = torch.zeros(m, dtype=torch.float64, device=dev)
    for tr in "NT":
        ms = timed(lambda: g.call("dgemv_", tr, m, m, 1.0, A, m, x, 1, 0.0, y, 1))
        gbs = 8.0 * (m * m + 3 * m) / ms / 1e6
        out["dgemv_%s_32768" % tr] = {"gbs": gbs, "ms": ms, "frac_of_measured_hbm": gbs / hbm}
    del A, x, y
    n = 1 << 26
    x = torch.rand(n, dtype=torch.float64, device=dev); y = torch.rand(n, dtype=torch.float64, device=dev)
    for name, byts, fn in (("ddot_2^26", 16.0 * n, lambda: g.call("ddot_", n, x, 1, y, 1, restype=ctypes.c_double)),
                           ("daxpy_2^26", 24.0 * n, lambda: g.call("daxpy_", n, 1e-9, x, 1, y, 1)),
                           ("dnrm2_2^26", 8.0 * n, lambda: g.call("dnrm2_", n, x, 1, restype=ctypes.c_double))):
        ms = timed(fn, reps=9)
        out[name] = {"gbs": byts / ms / 1e6, "ms": ms, "frac_of_measured_hbm": byts / ms / 1e6 / hbm}
    del x, y
    n = 1 << 28
    z = torch.rand(n, dtype=torch.float64, device=dev)
    ms = timed(lambda: g.call("idamax_", n, z, 1, restype=ctypes.c_int), reps=9)
    out["idamax_2^28"] = {"gbs": 8.0 * n / ms / 1e6, "ms": ms, "frac_of_measured_hbm": 8.0 * n / ms / 1e6 / hbm}
    del z
    # stand-alone DSYRK / DTRSM / DTRMM (north_star (1)); flops: SYRK k*n*(n+1), TRSM/TRMM left m^2*n, right m*n^2
    fp64 = peaks.get("fp64_tflops_probe") or FP64_PEAK_NOMINAL
    n = 16384
    A = torch.rand((n, n), dtype=torch.float64, device=dev) * 2 - 1
    C = torch.zeros((n, n), dtype=torch.float64, device=dev)
    ms = timed(lambda: g.call("dsyrk_", "L", "N", n, n, 1.0, A, n, 0.0, C, n), reps=3, warm=1)
    tf = float(n) * n * (n + 1) / ms / 1e9
    out["dsyrk_LN_16384"] = {"tflops": tf, "ms": ms, "frac_of_fp64_peak": tf / FP64_PEAK_NOMINAL, "frac_of_fp64_probe": tf / fp64}
    del C
    T = torch.triu(A[:8192, :8192]).contiguous()      # row-major upper == column-major LOWER triangle, 8192 x 8192
    T.mul_(1.0 / 8192); T.diagonal().fill_(1.0)      # small off-diagonals: the solution stays O(1)
    Bm = torch.rand((8192, 8192), dtype=torch.float64, device=dev)
    m = 8192
    for name, fn, fl in (("dtrsm_LLNN_8192", lambda: g.call("dtrsm_", "L", "L", "N", "N", m, m, 1.0, T, m, Bm, m), float(m) ** 3),
                         ("dtrmm_LLNN_8192", lambda: g.call("dtrmm_", "L", "L", "N", "N", m, m, 1.0, T, m, Bm, m), float(m) ** 3)):
        Bm.uniform_(-1, 1)
        ms = timed(fn, reps=3, warm=1)
        tf = fl / ms / 1e9
        out[name] = {"tflops": tf, "ms": ms, "frac_of_fp64_peak": tf / FP64_PEAK_NOMINAL, "frac_of_fp64_probe": tf / fp64}
    # the Cholesky panel solve: X * L^T = B with L 2048 x 2048, B 30720 x 2048 (flops m*n^2)
    mm, nn = 30720, 2048
    Lp = torch.triu(A[:nn, :nn]).contiguous(); Lp.mul_(1.0 / nn); Lp.diagonal().fill_(1.0)
    Bp = torch.rand((nn, mm), dtype=torch.float64, device=dev)          # column-major 30720 x 2048
    ms = timed(lambda: g.call("dtrsm_", "R", "L", "T", "N", mm, nn, 1.0, Lp, nn, Bp, mm), reps=3, warm=1)
    tf = float(mm) * nn * nn / ms / 1e9
    out["dtrsm_RLTN_30720x2048"] = {"tflops": tf, "ms": ms, "frac_of_fp64_peak": tf / FP64_PEAK_NOMINAL, "frac_of_fp64_probe": tf / fp64}
    # the same routines on HOST operands (pinned), through the symbol: chunked staging under the compute (csrc/staged_level3.cuh);
    # wall clock around the synchronous call, best of 3; "serial_copy_ms" = what whole-array copies at 50 GB/s would add in series
    try:
        hn = 8192
        hA = (torch.rand((hn, hn), dtype=torch.float64) * 2 - 1).pin_memory(); hC = torch.zeros((hn, hn), dtype=torch.float64).pin_memory()
        hT = torch.triu(hA).contiguous(); hT.mul_(1.0 / hn); hT.diagonal().fill_(1.0); hT = hT.pin_memory()
        def wall(fn, reps=3):
            fn(); torch.cuda.synchronize(); best = None
            for _ in range(reps):
                t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            return best * 1e3
        ms = wall(lambda: g.call("dsyrk_", "L", "N", hn, hn, 1.0, hA, hn, 0.0, hC, hn))
        out["dsyrk_LN_8192_host_pinned"] = {"tflops": float(hn) * hn * (hn + 1) / ms / 1e9, "ms": ms, "h2d_bytes": 8 * hn * hn, "d2h_bytes_max": int(0.66 * 8 * hn * hn)}
        hC.uniform_(-1, 1)
        ms = wall(lambda: g.call("dtrsm_", "L", "L", "N", "N", hn, hn, 1.0, hT, hn, hC, hn))
        out["dtrsm_LLNN_8192_host_pinned"] = {"tflops": float(hn) ** 3 / ms / 1e9, "ms": ms, "h2d_bytes": 16 * hn * hn, "d2h_bytes": 8 * hn * hn}
        del hA, hC, hT
    except Exception as exc:      # pinned allocation can fail on a small host; the resident lines above stand on their own
        out["level3_host_pinned_error"] = repr(exc)
    del A, T, Bm, Lp, Bp
    # blocked Cholesky workload (BASELINE.json configs[3]) on one GPU: wall clock, the driver synchronises per panel
    from libgpublas_b200.cholesky import blocked_cholesky
    n = 32768
    M = torch.rand((n, n), dtype=torch.float64, device=dev) * 2 - 1
    M = torch.tril(M, -1); M = M + M.T; M.diagonal().fill_(float(n))
    W = torch.empty_like(M)
    best = None
    for _ in range(3):
        W.copy_(M); torch.cuda.synchronize()
        t0 = time.perf_counter(); info = blocked_cholesky(n, W, n, 2048); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    out["cholesky_32768"] = {"tflops": n ** 3 / 3.0 / best / 1e12, "ms": best * 1e3, "info": info, "nb": 2048,
                             "frac_of_fp64_peak": n ** 3 / 3.0 / best / 1e12 / FP64_PEAK_NOMINAL,
                             "note": "device potrf + dtrsm_ + dsyrk_ through the Fortran symbols; flops n^3/3"}
    del M, W
    out.update(cholesky_single_call(g.load(), torch, dev, 1))
    # banded / packed Level-2 (csrc/level2_struct.cu): algorithmic bytes = the stored part of the matrix once
    nb, kl, ku = 1 << 22, 63, 64
    ab = torch.rand((nb, kl + ku + 1), dtype=torch.float64, device=dev)       # memory == column-major (kl+ku+1) x nb band storage
    xb = torch.rand(nb, dtype=torch.float64, device=dev); yb = torch.zeros(nb, dtype=torch.float64, device=dev)
    for tr in "NT":
        ms = timed(lambda: g.call("dgbmv_", tr, nb, nb, kl, ku, 1.0, ab, kl + ku + 1, xb, 1, 0.0, yb, 1))
        gbs = 8.0 * nb * (kl + ku + 1) / ms / 1e6
        out["dgbmv_%s_2^22_band128" % tr] = {"gbs": gbs, "ms": ms, "frac_of_measured_hbm": gbs / hbm}
    del ab, xb, yb
    npk = 32768
    ap = torch.rand(npk * (npk + 1) // 2, dtype=torch.float64, device=dev) * 1e-5
    xp = torch.rand(npk, dtype=torch.float64, device=dev)
    for tr in "NT":
        ms = timed(lambda: g.call("dtpmv_", "U", tr, "N", npk, ap, xp, 1))
        gbs = 8.0 * npk * (npk + 1) / 2 / ms / 1e6
        out["dtpmv_U%s_32768" % tr] = {"gbs": gbs, "ms": ms, "frac_of_measured_hbm": gbs / hbm}
    # symmetric / Hermitian products in one pass over the stored triangle (level2_struct.cu: sympart_kernel), the packed solve
    try:
        yp = torch.zeros(npk, dtype=torch.float64, device=dev)
        for ul in "UL":
            ms = timed(lambda: g.call("dspmv_", ul, npk, 1.0, ap, xp, 1, 0.0, yp, 1))
            gbs = 8.0 * npk * (npk + 1) / 2 / ms / 1e6
            out["dspmv_%s_32768" % ul] = {"gbs": gbs, "ms": ms, "frac_of_measured_hbm": gbs / hbm}
        idx = torch.arange(npk, device=dev, dtype=torch.int64)
        ap[idx * (idx + 1) // 2 + idx] = 2.0                     # diagonal of the packed upper triangle: a well-conditioned solve
        xp.fill_(1.0)
        ms = timed(lambda: g.call("dtpsv_", "U", "T", "N", npk, ap, xp, 1), reps=3, warm=1)
        out["dtpsv_UT_32768"] = {"ms": ms, "gbs": 8.0 * npk * (npk + 1) / 2 / ms / 1e6, "note": "latency-bound: 1024 dependent 32-blocks"}
        del ap, xp, yp, idx
        Sf = torch.rand((npk, npk), dtype=torch.float64, device=dev) * (1.0 / npk); Sf.diagonal().fill_(2.0)
        sx = torch.rand(npk, dtype=torch.float64, device=dev); sy = torch.zeros(npk, dtype=torch.float64, device=dev)
        ms = timed(lambda: g.call("dsymv_", "U", npk, 1.0, Sf, npk, sx, 1, 0.0, sy, 1))
        gbs = 8.0 * npk * (npk + 1) / 2 / ms / 1e6
        out["dsymv_U_32768"] = {"gbs": gbs, "ms": ms, "frac_of_measured_hbm": gbs / hbm}
        sx.fill_(1.0)
        ms = timed(lambda: g.call("dtrsv_", "L", "N", "N", npk, Sf, npk, sx, 1), reps=3, warm=1)
        out["dtrsv_LN_32768"] = {"ms": ms, "gbs": 8.0 * npk * (npk + 1) / 2 / ms / 1e6, "note": "latency-bound: 1024 dependent 32-blocks (panel solver)"}
        del Sf, sx, sy
    except Exception as exc:
        out["level2_sym_error"] = repr(exc)
    return out


def cpu_ksample_for_budget(n, calls, budget_s):
    """Largest k (multiple of 1024, <= n) such that `calls` OpenBLAS dgemm_ calls of m=n=`n` fit in budget_s, from a k=256 probe."""
    r = cpu_reference_run(n, 256, 1, 1)
    if r is None:
        return min(n, 1024)
    rate = r["value"] * 1e12                      # flop/s on this host (a k=256 slice runs a little below the large-k rate)
    kmax = budget_s * rate / (2.0 * n * n * max(1, calls))
    return int(max(1024, min(n, int(kmax) // 1024 * 1024))) if n >= 1024 else n


def cpu_reference_run(n, ksample, steps, warmup):
    """OpenBLAS dgemm_ on all host cores: m=n=`n`, k=`ksample` slice of the workload."""
    import numpy as np
    from helpers import f77, load_openblas, splitmix_uniform
    cores = os.cpu_count() or 1
    os.environ.setdefault("OPENBLAS_NUM_THREADS", str(cores))
    ob = load_openblas()
    if ob is None:
        return None
    try:
        ob.openblas_set_num_threads(ctypes.c_int(cores))
    except Exception:
        pass
    A = splitmix_uniform(2, (n, ksample)); B = splitmix_uniform(3, (ksample, n)); C = np.zeros((n, n), order="F")
    for _ in range(warmup):
        f77(ob, "dgemm_", "N", "N", n, n, ksample, 1.0, A, n, B, ksample, 0.0, C, n)
    t0 = time.perf_counter()
    for _ in range(steps):
        f77(ob, "dgemm_", "N", "N", n, n, ksample, 1.0, A, n, B, ksample, 0.0, C, n)
    dt = (time.perf_counter() - t0) / steps
    host = {"nproc": cores}      # report the host the CPU number comes from
    try:
        host["cpu_model"] = next(l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name"))
    except Exception:
        pass
    try:
        ob.openblas_get_config.restype = ctypes.c_char_p
        host["openblas_config"] = ob.openblas_get_config().decode()
        host["openblas_coretype"] = os.environ.get("OPENBLAS_CORETYPE")
    except Exception:
        pass
    return {"value": 2.0 * n * n * ksample / dt / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": "reference",
            "sample": "OpenBLAS 0.3.15 dgemm_ NN m=n=%d, k=%d slice of the k=%d workload, %d threads, %.2f s/step" % (n, ksample, n, cores, dt),
            "host": host, "_seconds": dt}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--n", "--size", dest="n", type=int, default=16384)   # --size: torchrun's own parser finds a bare --n ambiguous
    ap.add_argument("--ksample", type=int, default=0, help="reference arm / cpu_baseline: k of the CPU DGEMM sample (0: the largest "
                                                             "multiple of 1024 <= n that keeps the run within --cpu-budget seconds)")
    ap.add_argument("--cpu-budget", type=float, default=240.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the SGEMM/ZGEMM/Level-1/2 lines of BASELINE.json's metric")
    ap.add_argument("--no-verify", action="store_true", help="skip the parity check of the timed result (outside the timed region): "
                                                             "N=1 128 rows of C against a float64 numpy product, N>1 the whole C against the 1-GPU kernel")
    ap.add_argument("--mode", default="symbol", help="N>1: symbol = one process drives all N GPUs from inside dgemm_ (devices=N); "
                                                     "ipc = the round-1 one-process-per-GPU CUDA-IPC push (comparison)")
    ap.add_argument("--kchunk", type=int, default=1024)
    ap.add_argument("--distribute", default=None, help="N>1: p2p_push (default on CUDA) | bcast")
    args = ap.parse_args()
    n = args.n
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = "dgemm NN m=n=k=%d f64 alpha=1 beta=0 (BASELINE.json configs[1])" % n
    flops = 2.0 * n * n * n

    if args.impl == "reference":
        if rank != 0:
            return 0
        # the full k = n call when the host finishes (steps + warmup) of them within the budget (16 cores: ~8 s each), otherwise
        # the largest k-slice that does; `same_config` says which, and steps / warmup printed are the ones actually run
        ks = args.ksample or cpu_ksample_for_budget(n, args.steps + args.warmup, args.cpu_budget)
        r = cpu_reference_run(n, ks, max(1, args.steps), max(0, args.warmup))
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "no CPU BLAS (OpenBLAS) found in this image"}))
            return 0
        sec = r.pop("_seconds")
        line = {"impl": "reference", "metric": "dgemm_tflops", "value": r["value"], "unit": "TFLOP/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "sample": r["sample"], "same_config": ks == n},
                "cpu_baseline": r, "e2e": {"value": r["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import libgpublas_b200 as g
    legacy = world > 1 and args.mode == "ipc"
    if world > 1:
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # In symbol mode rank 0's process drives ALL N GPUs.  An NCCL barrier on the waiting ranks is a kernel that spins on
        # their GPU until rank 0 arrives and time-slices that GPU with rank 0's DGEMM (measured: 272 ms/step instead of 123 at
        # N = 2).  So the waiting ranks block on the CPU: barriers and the max-over-ranks reduction go through a gloo group.
        cpu_group = dist.new_group(backend="gloo") if not legacy else None
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", torch.cuda.current_device())
    lib = g.load()
    g.use_torch_stream()
    g.set_sync(False)
    sampler = ClockSampler(torch.cuda.current_device())

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group) if cpu_group is not None else dist.barrier()

    verified = None
    e2e = None
    others = None
    probe = None
    cpu = None
    mg = None
    if legacy:
        # round-1 path kept for comparison (--mode ipc): one process per GPU, CUDA-IPC panel push by the home GPU
        from libgpublas_b200.multigpu import TiledGemm
        tg = TiledGemm(n, n, n, dev, rank, world, kchunk=args.kchunk, distribute=args.distribute)
        tg.make_inputs(seed=2)
        for _ in range(args.warmup):
            tg.run()
        sync_all()
        if rank == 0:
            sampler.start()
            time.sleep(0.3)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            tg.run()
        e1.record()
        sync_all()
        clocks = sampler.stop() if rank == 0 else {}
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = t.item() / args.steps
        value = flops / (ms_per_step * 1e-3) / 1e12
        kernel_ms = ms_per_step
        variant = g.last_variant()
        launches = args.steps * tg.kernels_per_step
        scaling, parallelism = "strong", tg.describe()
        if not args.no_verify and rank == 0:
            ref = torch.empty(n * n, dtype=torch.float64, device=dev)
            g.call("dgemm_", "N", "N", n, n, n, 1.0, tg.A, n, tg.B, n, 0.0, ref, n)
            torch.cuda.synchronize()
            got = tg.home_c()
            verified = {"max_abs_diff_vs_1gpu": float((got - ref).abs().max().item())}
            del ref, got
    else:
        # N = 1, and N > 1 "behind the symbol": rank 0 calls dgemm_ exactly as on one GPU; with devices=N the library
        # partitions the product over the N GPUs of the box itself (csrc/multi_gemm.cu).  The other ranks torchrun started
        # only take part in the barriers: the contract's launch, barrier and max-over-ranks clock stay, the data path has
        # no process boundary in it -- which is what an unmodified program under LD_PRELOAD gets.
        if world > 1:
            lib.b200blas_set_options(("devices=%d" % world).encode())
        launches = 0
        clocks = {}
        ms_local = 0.0
        if rank == 0:
            gen = torch.Generator(device=dev).manual_seed(2)
            A = torch.rand((n, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
            B = torch.rand((n, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
            C = torch.zeros((n, n), dtype=torch.float64, device=dev)

            def step():
                g.call("dgemm_", "N", "N", n, n, n, 1.0, A, n, B, n, 0.0, C, n)

            for _ in range(args.warmup):
                step()
        sync_all()
        if rank == 0:
            mg0 = mg_stats(lib)
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
            sampler.start()
            time.sleep(0.3)
            evs[0].record()
            for i in range(args.steps):
                step()
                evs[i + 1].record()
        sync_all()
        if rank == 0:
            clocks = sampler.stop()
            ms_local = evs[0].elapsed_time(evs[-1])
            per_launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
            mg1 = mg_stats(lib)
        if world > 1:
            t = torch.tensor([ms_local], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=cpu_group)
            ms_local = t.item()
        ms_per_step = ms_local / args.steps
        value = flops / (ms_per_step * 1e-3) / 1e12
        scaling = "strong"
        if rank == 0:
            variant = g.last_variant()
            kernel_ms = sum(per_launch_ms) / len(per_launch_ms)
            launches = args.steps * world
            if world > 1:
                calls = mg1[0] - mg0[0]
                assert calls == args.steps, "the timed calls did not take the partitioned path (%d of %d)" % (calls, args.steps)
                P, Q = (1, 2) if world == 2 else ((2, world // 2) if world % 2 == 0 else (1, world))
                mg = {"partitioned_calls": calls, "devices": mg1[1], "origin_gb_per_step": (mg1[2] - mg0[2]) / calls / 1e9,
                      "forwarded_gb_per_step": (mg1[3] - mg0[3]) / calls / 1e9, "hops_per_step": (mg1[4] - mg0[4]) // calls}
                parallelism = ("one process, devices=%d behind dgemm_: 2-D tiles on a device grid; A row-groups / B column bands leave the home GPU once and "
                               "are forwarded along chains of the devices that need them (copy engines over NVLink, a flag per piece); one flag-polling DMMA "
                               "launch per device; C tiles stored into the home allocation by the kernel epilogues" % world)
            else:
                parallelism = "single"
            if not args.no_verify:
                verified = verify_rows(torch, A, B, C, n)
                if world > 1:      # and bit-for-bit against the 1-GPU kernel on the same operands
                    lib.b200blas_set_options(b"devices=1")
                    ref = torch.empty((n, n), dtype=torch.float64, device=dev)
                    g.call("dgemm_", "N", "N", n, n, n, 1.0, A, n, B, n, 0.0, ref, n)
                    torch.cuda.synchronize()
                    verified["max_abs_diff_vs_1gpu"] = float((C - ref).abs().max().item())
                    del ref
                    lib.b200blas_set_options(("devices=%d" % world).encode())

        # ---- end-to-end through the C ABI with pinned host buffers: the library stages A, B host->device and C back inside the call ----
        if rank == 0 and not args.no_e2e:
            g.set_sync(True)
            hA = torch.empty((n, n), dtype=torch.float64).pin_memory(); hA.copy_(A.cpu())
            hB = torch.empty((n, n), dtype=torch.float64).pin_memory(); hB.copy_(B.cpu())
            hC = torch.empty((n, n), dtype=torch.float64).pin_memory()
            nA, nB, nC = hA.numpy(), hB.numpy(), hC.numpy()
            for _ in range(min(args.warmup, 2)):
                g.call("dgemm_", "N", "N", n, n, n, 1.0, nA, n, nB, n, 0.0, nC, n)
            torch.cuda.synchronize()
            s1 = g.stats()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                g.call("dgemm_", "N", "N", n, n, n, 1.0, nA, n, nB, n, 0.0, nC, n)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / args.steps
            s2 = g.stats()
            # the result read back must agree with the device-resident one.  At N = 1 the staging pipeline accumulates C over
            # k-chunks (another summation order than the single k loop): equal to rounding, |diff| <= 16*eps*k for U(-1,1)
            # data; the partitioned path keeps the single k loop: equal bit for bit.  Two corners (first and last panel).
            d0 = float((hC[:256, :256].double() - C[:256, :256].cpu().double()).abs().max())
            d1 = float((hC[n - 256:, n - 256:].double() - C[n - 256:, n - 256:].cpu().double()).abs().max())
            same = bool(max(d0, d1) <= 16 * 2.0 ** -53 * n)
            e2e = {"value": flops / dt / 1e12, "unit": "TFLOP/s", "ms_per_step": dt * 1e3,
                   "h2d_bytes_per_step": (s2["h2d_bytes"] - s1["h2d_bytes"]) // args.steps,
                   "d2h_bytes_per_step": (s2["d2h_bytes"] - s1["d2h_bytes"]) // args.steps,
                   "host_memory": "pinned", "matches_device_result": same, "max_abs_diff_vs_resident": max(d0, d1),
                   "path": ("dgemm_ on host pointers: chunked H2D of A/B and D2H of C panels overlapped with the DMMA kernel (csrc/staged_gemm.cuh)" if world == 1 else
                            "dgemm_ on host pointers, devices=%d: every GPU pulls its share of the A/B pieces over its own PCIe link, siblings "
                            "forward them over NVLink, C tiles return device->host per GPU (csrc/multi_gemm.cu)" % world)}
            # the reference's real miss path is plain malloc'd memory (runtime-mem.hpp:84-112), and its hit path a calloc'd block
            # the CPU has just filled: the same call on PAGEABLE host buffers, and on tracked managed buffers at first touch
            if world == 1:
                try:
                    e2e["pageable"] = e2e_pageable(g, n, nA, nB, C, flops)
                    e2e["managed_first_touch"] = e2e_managed_first_touch(g, lib, n, nA, nB, C, flops)
                except Exception as exc:  # noqa: BLE001
                    e2e["pageable_error"] = repr(exc)
            g.set_sync(False)
            del hA, hB, hC, nA, nB, nC

        if rank == 0:
            try:
                lib.b200blas_probe_fp64_tflops.restype = ctypes.c_double
                lib.b200blas_probe_fp64_tflops.argtypes = [ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
                burst = ctypes.c_double(0.0)
                probe = {"sustained": lib.b200blas_probe_fp64_tflops(1.5, ctypes.byref(burst)), "burst": burst.value,
                         "how": "register-resident mma.sync m16n8k16 f64 (DMMA) loop, 8 warps/SM, 1.5 s, in this process (csrc/probe.cu)"}
            except Exception as exc:  # noqa: BLE001
                probe = {"error": repr(exc)}
        if rank == 0 and not args.no_others:
            del A, B, C
            torch.cuda.empty_cache()
            others = {}
            try:      # the headline line must print whatever happens to a secondary measurement
                pk = measured_peaks()
                if probe and probe.get("sustained"):
                    pk["fp64_tflops_probe"] = probe["sustained"]
                if world == 1:
                    other_routines(g, torch, dev, pk, others)
                else:
                    other_routines_partitioned(g, lib, torch, dev, pk, others, world)
            except Exception as exc:  # noqa: BLE001
                others["error"] = repr(exc)
            if world == 1:
                try:
                    others["level1_from_c"] = l1_c_driver(measured_peaks())
                except BaseException as exc:  # noqa: BLE001  (pytest.skip raises outside Exception)
                    others["level1_from_c"] = {"error": repr(exc)}
        if rank == 0 and world == 1 and not args.no_cpu:
            cpu = cpu_reference_run(n, args.ksample or cpu_ksample_for_budget(n, 3, 20.0), 2, 1)
            if cpu:
                cpu.pop("_seconds", None)

    if rank == 0:
        peaks = measured_peaks()
        per_gpu = flops / (kernel_ms * 1e-3) / 1e12 if world == 1 else value / world
        line = {"metric": "dgemm_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": {"workload": workload, "parallelism": parallelism, "variant": variant,
                                                "l2": "no flush needed: A+B+C = %.1f GB >> 126 MB L2" % (3 * 8.0 * n * n / 1e9)},
                "roofline": {"bound": "tensor", "achieved": per_gpu, "peak": FP64_PEAK_NOMINAL, "unit": "TFLOP/s",
                             "frac": per_gpu / FP64_PEAK_NOMINAL, "traffic": None,
                             "peak_source": "FP64 tensor (DMMA) pipe: nominal 148 SM x 128 flop/clk x 1.965 GHz = 37.2 (per GPU); MEASURED_PEAKS.json has no "
                                            "FP64 entry (bf16 %.0f, HBM %.0f GB/s), so the ceiling is also measured in this process (peak_measured_in_process)"
                                            % (peaks.get("bf16_tflops", 0), peaks.get("hbm_gbs", 0)),
                             "kernel_ms": kernel_ms, "peak_measured_in_process": probe,
                             "frac_of_measured": (per_gpu / probe["sustained"]) if probe and probe.get("sustained") else None},
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "others": others}
        tr = traffic_from_profile()
        if tr and world == 1:
            line["roofline"]["traffic"] = tr["dram_bytes_per_launch"]
            line["roofline"]["traffic_source"] = tr["source"]
        if verified is not None:
            line["verified"] = verified
        if mg is not None:
            line["partitioned"] = mg
        print(json.dumps(line))
    if world > 1:
        sync_all()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
