#!/usr/bin/env python
"""Micro-benchmarks for BASELINE.json configs[4]: scan / projection sweeps over state size and sequence length,
batch 1 (latency) and large batch (throughput), each against the HBM roofline; plus the PCIe ceiling the end-to-end
number of bench.py runs against.

    python bench_micro.py [--out profiles/rNN_micro.json] [--quick]

Prints one JSON line per case and (with --out) writes the list.  Timing: CUDA events on the launching stream, median
of `reps` launches, a 256 MB buffer is written between launches so nothing is L2-resident at the start of a launch.
The CUDA extension is required (no fallback).
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        for k in ("hbm_gbs", "hbm_copy_gbs", "hbm_gbs_burst"):
            if k in d:
                return float(d[k]), f"measured (MEASURED_PEAKS.json {k})"
    except Exception:
        pass
    return 7700.0, "fallback (B200_PROFILING.md nominal)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--reps", type=int, default=9)
    ap.add_argument("--only", default=None, help="comma-separated kernel names (scan_c64, scan_i32, spmm_f32, pcie)")
    args = ap.parse_args()

    import torch

    from sparsernns_b200 import _abi

    lib = _abi.lib()
    if not torch.cuda.is_available():
        raise SystemExit("bench_micro.py needs a CUDA device")
    dev = torch.device("cuda:0")
    peak, peak_src = hbm_peak()
    rng = np.random.default_rng(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    results = []

    def timed(fn, reps):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    def emit(rec):
        results.append(rec)
        print(json.dumps(rec), flush=True)

    def lam_of(P):
        return np.exp(-rng.uniform(0.01, 0.1, P) + 1j * rng.uniform(-1, 1, P)).astype(np.complex64)

    # ---- complex scan: 16 bytes per state-step (8 in, 8 out)
    def scan_c64(B, L, P):
        by = 16.0 * B * L * P
        # small problems: several buffer sets (together larger than L2) so that launches can also be timed back to back
        nbuf = 1 if by > 1e9 else int(min(8, max(2, -(-400e6 // by))))
        Bu = [torch.randn(B, L, P, 2, device=dev) for _ in range(nbuf)]
        out = [torch.empty_like(Bu[0]) for _ in range(nbuf)]
        lam = torch.from_numpy(lam_of(P)).to(dev)
        wsb = lib.s5b200_scan_c64_workspace_bytes(B, L, P)
        ws = [torch.empty(wsb, dtype=torch.uint8, device=dev) for _ in range(nbuf)]

        def one(i):
            _abi.check(lib.s5b200_scan_c64(Bu[i].data_ptr(), lam.data_ptr(), out[i].data_ptr(), B, L, P, 0, ws[i].data_ptr(), wsb, st),
                       "scan_c64")

        ms = timed(lambda: one(0), args.reps)
        rec = {"kernel": "scan_c64", "B": B, "L": L, "P": P, "us": ms * 1e3, "alg_bytes": by, "gbs": by / ms / 1e6,
               "frac": by / ms / 1e6 / peak, "state_steps_per_s": B * L * P / ms * 1e3}
        if nbuf > 1:
            # the same launch repeated over `nbuf` distinct buffer sets inside ONE event pair: launch latency overlaps the
            # previous kernel, as it does inside the model; every launch still reads data that is not in L2
            msb = timed(lambda: [one(i) for i in range(nbuf)], args.reps) / nbuf
            rec.update({"us_back_to_back": msb * 1e3, "gbs_back_to_back": by / msb / 1e6, "frac_back_to_back": by / msb / 1e6 / peak,
                        "buffers": nbuf})
        emit(rec)

    # ---- integer scan (planar int32), 16 bytes per state-step
    def scan_i32(B, L, P):
        mk = lambda: torch.randint(-(1 << 14), 1 << 14, (B, L, P), dtype=torch.int32, device=dev)
        bre, bim = mk(), mk()
        xre, xim = torch.empty_like(bre), torch.empty_like(bim)
        lam = lam_of(P)
        are = torch.from_numpy(np.round(lam.real * (1 << 14)).astype(np.int32)).to(dev)
        aim = torch.from_numpy(np.round(lam.imag * (1 << 14)).astype(np.int32)).to(dev)

        def fn():
            _abi.check(lib.s5b200_scan_i32(bre.data_ptr(), bim.data_ptr(), are.data_ptr(), aim.data_ptr(), xre.data_ptr(),
                                           xim.data_ptr(), B, L, P, 14, 14, 10, 10, 10, 10, 0, st), "scan_i32")

        ms = timed(fn, max(3, args.reps // 3))
        by = 16.0 * B * L * P
        emit({"kernel": "scan_i32", "B": B, "L": L, "P": P, "us": ms * 1e3, "alg_bytes": by, "gbs": by / ms / 1e6,
              "frac": by / ms / 1e6 / peak, "state_steps_per_s": B * L * P / ms * 1e3})

    # ---- sparse projection, (K + N) * 4 bytes per row
    def spmm(rows, K, N, density, act_zero=0.0, dead=0.0, skip=0):
        """act_zero: fraction of activations zeroed element-wise (ReLU); dead: fraction of input channels that never fire."""
        W = rng.standard_normal((K, N)).astype(np.float32)
        W[rng.random((K, N)) > density] = 0.0
        h = C.c_void_p()
        _abi.check(lib.s5b200_sparse_create(W.ctypes.data, K, N, 0, 0, C.byref(h)), "sparse_create")
        X = torch.randn(rows, K, device=dev)
        if act_zero > 0:
            X = X * (torch.rand(rows, K, device=dev) >= act_zero)
        if dead > 0:
            X = X * (torch.rand(K, device=dev) >= dead)
        Y = torch.empty(rows, N, device=dev)

        def fn():
            _abi.check(lib.s5b200_spmm_f32(h, X.data_ptr(), None, Y.data_ptr(), rows, skip, st), "spmm_f32")

        ms = timed(fn, args.reps)
        lib.s5b200_sparse_destroy(h)
        by = 4.0 * rows * (K + N)
        emit({"kernel": "spmm_f32", "rows": rows, "K": K, "N": N, "density": density, "act_zero": act_zero, "dead_channels": dead,
              "skip_zero": skip, "us": ms * 1e3, "alg_bytes": by, "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peak})

    # ---- PCIe ceiling for the e2e number: pinned 123 MB each way, alone and concurrently
    def pcie():
        n = 32 * 3751 * 257
        h_in = torch.empty(n, dtype=torch.float32).pin_memory()
        h_out = torch.empty(n, dtype=torch.float32).pin_memory()
        d_in = torch.empty(n, dtype=torch.float32, device=dev)
        d_out = torch.zeros(n, dtype=torch.float32, device=dev)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def run(h2d, d2h, reps=5):
            ts = []
            for _ in range(reps + 1):
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                s1.wait_stream(torch.cuda.current_stream())
                s2.wait_stream(torch.cuda.current_stream())
                if h2d:
                    with torch.cuda.stream(s1):
                        d_in.copy_(h_in, non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_out, non_blocking=True)
                torch.cuda.current_stream().wait_stream(s1)
                torch.cuda.current_stream().wait_stream(s2)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return float(np.median(ts[1:]))

        by = n * 4.0
        a, b, c = run(True, False), run(False, True), run(True, True)
        emit({"kernel": "pcie", "bytes_each_way": by, "h2d_ms": a, "h2d_gbs": by / a / 1e6, "d2h_ms": b, "d2h_gbs": by / b / 1e6,
              "both_ms": c, "both_gbs_each_way": by / c / 1e6, "e2e_ceiling_frames_per_s": 32 * 3751 / c * 1e3})

    only = set(args.only.split(",")) if args.only else None
    if only is not None:   # run just the named sweeps
        keep = lambda f, name: f if name in only else (lambda *a, **k: None)
        pcie, scan_c64, scan_i32, spmm = keep(pcie, "pcie"), keep(scan_c64, "scan_c64"), keep(scan_i32, "scan_i32"), keep(spmm, "spmm_f32")
    emit({"hbm_peak_gbs": peak, "peak_source": peak_src, "device": torch.cuda.get_device_name(0)})
    pcie()
    if args.quick:
        cases = [(1, 4096, 64), (32, 3751, 64), (4096, 1024, 64)]
    else:
        cases = []
        for P in (16, 64, 128, 256):
            for L in (1024, 4096, 16384, 65536):
                cases.append((1, L, P))
            cases.append((4096, 1024, P))
        cases += [(4096, 4096, 64), (256, 16384, 64), (64, 65536, 64), (32, 3751, 64), (32, 3751, 128), (256, 3751, 64)]
    for c in cases:
        scan_c64(*c)
    for c in ([(32, 3751, 64)] if args.quick else [(1, 4096, 64), (32, 3751, 64), (256, 3751, 64), (4096, 1024, 64), (4096, 1024, 128)]):
        scan_i32(*c)
    for rows in ((120032,) if args.quick else (128, 4096, 120032, 4096 * 1024)):
        for (K, N, d) in ((96, 128, 0.1), (128, 96, 0.1), (257, 96, 0.1), (96, 128, 1.0)):
            spmm(rows, K, N, d)
    # zero-activation skipping (per input column x 32-row group): element-wise ReLU zeros alone rarely empty a group,
    # channels that never fire always do
    for (az, dead) in ((0.0, 0.0), (0.5, 0.0), (0.9, 0.0), (0.99, 0.0), (0.5, 0.5), (0.5, 0.9)):
        for skip in (0, 1):
            spmm(120032, 128, 96, 0.1, az, dead, skip)
    if args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
