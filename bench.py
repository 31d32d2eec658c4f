#!/usr/bin/env python
"""bench.py - S5 NDNS inference throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one forward pass of the NDNS model (encoder Dense -> 3 S5 blocks -> decoder Dense) over one
batch of synthetic STFT-magnitude clips per GPU.  Workload at every N: BASELINE.json configs[1] - the
ndns.json recipe at dim_scale 0.5 (H 96, P 64), 90 % magnitude-pruned, ReLU-ified, static w8a16 - with the
recipe's batch of 32 clips x 3751 frames per GPU (weak scaling: clips are independent, no data-path
collective; NCCL only reduces the metric checksum after the timed region).

  value      frames/s with inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same through the host-buffer entry point (pinned H2D of the inputs + D2H of the mask per step)
  roofline   dominant kernel: algorithmic bytes per launch / mean device time per launch (CUDA events
             recorded by the library around each launch) against the measured HBM peak
  cpu_baseline / --impl reference   the NumPy port of the reference (oracle/) on the box's host cores;
             JAX is not installable here, so the reference itself cannot run
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L_FRAMES = 3751
D_IO = 257


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips-per-gpu", type=int, default=32)
    ap.add_argument("--frames", type=int, default=L_FRAMES)
    ap.add_argument("--dim-scale", type=float, default=0.5, choices=[0.5, 1.0])
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary flavours and the CPU baseline")
    return ap.parse_args()


def load_fixture(dim_scale):
    from sparsernns_b200.checkpoint import load_npz

    path = os.path.join(ROOT, "tests", "golden", f"ndns_ds{dim_scale}_sparse_w8a16.npz")
    tree = load_npz(path)
    return tree["params"], tree["batch_stats"], tree["cal_stats"]


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.lines, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [s.strip() for s in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def synth_inputs(torch, n_batches, clips, frames, seed, device):
    """|STFT| of white noise - 0.0007, drawn directly from its Rayleigh law (sparsernns_b200.synth docstring)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    sigma = 0.05 * (512 / 2.0) ** 0.5 / 512
    out = []
    for _ in range(n_batches):
        re = torch.randn(clips, frames, D_IO, generator=g, device=device)
        im = torch.randn(clips, frames, D_IO, generator=g, device=device)
        out.append((sigma * torch.sqrt(re * re + im * im) - 0.0007).contiguous())
    return out


def algorithmic_bytes(rows, H, P):
    """Bytes that must cross HBM per launch of each kernel class ('Kernels')."""
    return {"enc": rows * (D_IO + H) * 4, "bproj": rows * (H * 4 + P * 8), "scan_c64": rows * P * 16,
            "glu": rows * (P * 8 + 2 * H * 4), "dec": rows * (H + D_IO) * 4,
            "glu_scan": rows * (P * 8 + 2 * H * 4),   # reads Bu instead of the states; the states stay on the SM
            "block": rows * 2 * H * 4,                # one kernel per S5 block: block input in, block output out
            "fxp_enc": rows * (D_IO + H) * 4, "fxp_bn": rows * H * 4, "fxp_bproj": rows * (H * 4 + P * 8),
            "scan_i32": rows * P * 16, "fxp_glu": rows * (P * 8 + 2 * H * 4), "fxp_resid": rows * H * 12,
            "fxp_dec": rows * (H + D_IO) * 4}


def cpu_port_forward_fn(params, batch_stats, dims):
    """The NumPy port of the reference's static-w8a16 forward (oracle/): cpu_baseline and --impl reference only."""
    from oracle import model as OM

    cfg = OM.ModelConfig(n_layers=dims.n_layers, d_model=dims.d_model, ssm_size=dims.ssm_size, relufication=True,
                         static_quant=True, associative_scan=False)
    var = {"params": params, "batch_stats": batch_stats}
    return lambda x: OM.model_forward(x, var, cfg)


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
        return int(n)
    except Exception:
        return os.cpu_count() or 1


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU path.  JAX/Flax are not installable in this image (no wheel, no
    network), so the reference is represented by its NumPy port under oracle/ (kind 'port')."""
    if rank != 0:
        return
    import numpy as np
    from sparsernns_b200 import synth

    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs alone on rank 0 and may use every host core
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass
    dims = synth.ndns_dims(args.dim_scale)
    params, batch_stats, _ = load_fixture(args.dim_scale)
    fwd = cpu_port_forward_fn(params, batch_stats, dims)
    probe = synth.synthetic_stft_batch(1, args.frames, seed=1)
    t0 = time.perf_counter(); fwd(probe); t1 = time.perf_counter() - t0
    budget = 150.0  # seconds for the whole run
    clips = max(1, min(args.clips_per_gpu, int(budget / max(t1, 1e-3) / max(1, args.steps + args.warmup))))
    x = synth.synthetic_stft_batch(clips, args.frames, seed=2)
    sub = 8   # the port is fastest on sub-batches of about 8 clips (working set; same sub-batch as the cpu_baseline leg): 52 k against 31 k frames/s

    def step():
        for s0 in range(0, clips, sub):
            fwd(x[s0:s0 + sub])

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = clips * args.frames * args.steps / dt
    cores = host_threads()
    line = {
        "impl": "reference", "metric": "S5 NDNS inference frames/sec (sparse static w8a16, dim_scale %.1f)" % args.dim_scale,
        "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ndns.json dim_scale %.1f, 90%% magnitude-pruned, relufication, static w8a16 forward" % args.dim_scale,
                   "clips_per_step": clips, "frames_per_clip": args.frames, "H": dims.d_model, "P": dims.ssm_size,
                   "note": "NumPy port of the reference (oracle/); the JAX reference is not installable offline"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "%d clip(s) x %d frames per step in sub-batches of %d, %d steps" % (clips, args.frames, sub, args.steps)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_REAL_STDOUT = None


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL's version banner, torchrun's OMP notice) write to fd 1
    behind Python's back, so fd 1 is pointed at stderr for the whole run and the result goes to a saved copy."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    args = parse_args()
    _claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import numpy as np
    import torch

    from sparsernns_b200 import _abi, synth
    from sparsernns_b200.build import build_library
    from sparsernns_b200.model import BoundModel, S5Config

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    build_library()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the single JSON line
        dist.init_process_group("nccl", device_id=dev)

    from sparsernns_b200.dist import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local_rank)   # pinned host buffers of the e2e leg land on the GPU's NUMA node

    dims = synth.ndns_dims(args.dim_scale)
    params, batch_stats, cal_stats = load_fixture(args.dim_scale)
    cfg = S5Config(n_layers=dims.n_layers, d_model=dims.d_model, ssm_size=dims.ssm_size, relufication=True,
                   quantization="w8a16", use_sparse=True)
    model = BoundModel(cfg, {"params": params, "batch_stats": batch_stats})
    lib = _abi.lib()
    B, L, K, W = args.clips_per_gpu, args.frames, args.steps, args.warmup
    rows = B * L
    n_rot = 4  # rotate over 4 input batches (4 x 123 MB at B=32) so that no step finds its input in the 126 MB L2
    xs = synth_inputs(torch, n_rot, B, L, 1000 + rank, dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    y = None

    def step(i):
        nonlocal y
        y = model(xs[i % n_rot])

    for i in range(max(W, 3)):
        step(i)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    lib.s5b200_reset_launch_count()
    ms_total = timed(step, K)
    launches = int(lib.s5b200_launch_count())
    clock_info = clocks.stop() if rank == 0 else None
    fps = world * rows * K / (ms_total / 1e3)

    # metric checksum: the only collective of the run (NCCL), after the timed region
    chk = torch.stack([y.double().sum(), y.double().abs().max()])
    if dist is not None:
        from sparsernns_b200.dist import reduce_metrics
        chk = reduce_metrics(chk, "sum")

    # end to end through the host-buffer entry point
    xh = [t.cpu().pin_memory() for t in xs[:2]]
    yh = torch.empty((B, L, D_IO), dtype=torch.float32).pin_memory()
    # Host-synchronous calls: every step copies its 123 MB of input from pinned host memory, runs the forward and
    # copies its 123 MB of output back before the call returns.
    yh2 = [yh, torch.empty_like(yh).pin_memory()]
    model.forward_host(xh[0], out=yh)
    barrier()
    ke = max(3, min(K, 10))
    t0 = time.perf_counter()
    for i in range(ke):
        model.forward_host(xh[i % 2], out=yh2[i % 2])
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    # the stream-ordered form of the same call (several batches in flight, one synchronize at the end), for information
    barrier()
    t0 = time.perf_counter()
    for i in range(ke):
        model.forward_host(xh[i % 2], out=yh2[i % 2], sync=False)
    torch.cuda.synchronize()
    e2e_async_s = torch.tensor([time.perf_counter() - t0], device=dev)
    # the same call with TWO batches in flight: consecutive steps go to two alternating caller streams, so the H2D copies of
    # step i + 1 start while step i is still computing / copying back (on ONE caller stream a step is ordered after the
    # previous step's last D2H copy).  Every step still moves its 123 MB in and its 123 MB out inside the timed region;
    # one synchronize at the end, like the device-timed value above.
    y_sync = [t.clone() for t in yh2]
    pipe_streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for i in range(2):
        with torch.cuda.stream(pipe_streams[i]):
            model.forward_host(xh[i], out=yh2[i], sync=False)
    torch.cuda.synchronize()
    barrier()
    kp = 2 * ke
    t0 = time.perf_counter()
    for i in range(kp):
        with torch.cuda.stream(pipe_streams[i % 2]):
            model.forward_host(xh[i % 2], out=yh2[i % 2], sync=False)
    torch.cuda.synchronize()
    e2e_pipe_s = torch.tensor([time.perf_counter() - t0], device=dev)
    pipe_equal = all(torch.equal(a, b) for a, b in zip(y_sync, yh2))
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_async_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_pipe_s, op=dist.ReduceOp.MAX)
    e2e_sync_fps = world * rows * ke / float(e2e_s.item())
    e2e_async_fps = world * rows * ke / float(e2e_async_s.item())
    e2e_pipe_fps = world * rows * kp / float(e2e_pipe_s.item())
    e2e_fps = e2e_pipe_fps

    # informational: the application's own end to end (the forward of the reference's NDNS eval step, train_helpers.py:1570-1614):
    # noisy AUDIO in pinned host memory -> H2D -> STFT -> model -> mask apply + iSTFT -> D2H of the cleaned audio.  A quarter of
    # the PCIe bytes of the tensor-level e2e above, because spectra never cross the bus.  Only when L is the STFT of whole hops.
    e2e_audio = None
    if not args.no_extras and L >= 2:
        from sparsernns_b200 import ndns
        n_samp = (L - 1) * 128
        ah = [(0.05 * torch.randn(B, n_samp)).pin_memory() for _ in range(2)]
        oh = [torch.empty(B, n_samp).pin_memory() for _ in range(2)]
        nch = 4 if B % 4 == 0 else 1                                       # clip chunks, one stream each: copies overlap kernels
        nst = 4 if nch == 4 else 1
        streams = [torch.cuda.Stream() for _ in range(nst)]
        # one handle per stream: a BoundModel owns ONE workspace (look-back flags included), so forwards in flight need their own
        models = [model] + [BoundModel(cfg, {"params": params, "batch_stats": batch_stats}) for _ in range(nst - 1)]

        def audio_step(i):
            for c in range(nch):
                with torch.cuda.stream(streams[c % nst]):
                    sl = slice(c * B // nch, (c + 1) * B // nch)
                    a = ah[i % 2][sl].to(dev, non_blocking=True)
                    mag, ph, xin = ndns.stft_splitter(a, blf=True, model_input=True)
                    cleaned = ndns.stft_mixer(mag, ph, blf=True, mask=models[c % nst](xin))
                    oh[i % 2][sl].copy_(cleaned, non_blocking=True)

        for i in range(2):
            audio_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(ke):
            audio_step(i)
        torch.cuda.synchronize()
        ea_s = torch.tensor([time.perf_counter() - t0], device=dev)
        oh_loop = [t.clone() for t in oh]
        if dist is not None:
            dist.all_reduce(ea_s, op=dist.ReduceOp.MAX)
        loop_fps = world * rows * ke / float(ea_s.item())
        # the same path as ONE library call per batch (s5b200_ndns_denoise_host_async: the clip-chunk pipeline of the tensor-level call with
        # STFT / iSTFT either side of the forward), two batches in flight on alternating caller streams like the e2e leg above
        for i in range(2):
            with torch.cuda.stream(pipe_streams[i]):
                ndns.denoise_host(model, ah[i], out=oh[i], sync=False)
        torch.cuda.synchronize()
        lib_equal = all(torch.equal(oh[i], oh_loop[i]) for i in range(2))
        barrier()
        t0 = time.perf_counter()
        for i in range(kp):
            with torch.cuda.stream(pipe_streams[i % 2]):
                ndns.denoise_host(model, ah[i % 2], out=oh[i % 2], sync=False)
        torch.cuda.synchronize()
        el_s = torch.tensor([time.perf_counter() - t0], device=dev)
        if dist is not None:
            dist.all_reduce(el_s, op=dist.ReduceOp.MAX)
        e2e_audio = {"value": world * rows * kp / float(el_s.item()), "unit": "frames/s", "h2d_bytes_per_step": B * n_samp * 4,
                     "d2h_bytes_per_step": B * n_samp * 4, "steps": kp,
                     "what": "noisy audio (pinned host) -> H2D -> STFT -> model -> mask apply + iSTFT -> D2H of the cleaned audio: one "
                             "s5b200_ndns_denoise_host_async call per batch, two batches in flight",
                     "outputs_equal_python_loop": lib_equal,
                     "python_loop_value": loop_fps,
                     "python_loop": "%d clip chunks on %d streams over sparsernns_b200.ndns" % (nch, nst)}

    # per-kernel device time (events recorded by the library around each launch); the end-to-end leg above leaves copies
    # on other streams behind: drain them and re-warm before timing single launches
    torch.cuda.synchronize()
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    lib.s5b200_profile_begin()
    for i in range(K):
        step(i)
    ms_cls = (C.c_double * len(_abi.KERNEL_CLASSES))()
    n_cls = (C.c_longlong * len(_abi.KERNEL_CLASSES))()
    _abi.check(lib.s5b200_profile_end(ms_cls, n_cls))
    peak, peak_src = measured_peak()
    alg = algorithmic_bytes(rows, dims.d_model, dims.ssm_size)
    kernels = {}
    for i, name in enumerate(_abi.KERNEL_CLASSES):
        if n_cls[i]:
            per = ms_cls[i] / n_cls[i]
            ab = alg.get(name)
            if name == "block" and ab:
                # all blocks may share one launch: algorithmic bytes of a launch = blocks per launch x (input + output)
                ab = ab * dims.n_layers / max(1.0, n_cls[i] / K)
            gbs = ab / (per * 1e-3) / 1e9 if ab else None
            kernels[name] = {"launches_per_step": n_cls[i] / K, "ms_per_launch": per, "ms_per_step": ms_cls[i] / K,
                             "alg_bytes_per_launch": ab, "achieved_gbs": gbs, "frac": gbs / peak if gbs else None}
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(dom)
    except Exception:
        pass
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                "share_of_step": kernels[dom]["ms_per_step"] / sum(k["ms_per_step"] for k in kernels.values())}
    if dom in ("block", "glu_scan"):
        roofline["note"] = ("fused kernel (all S5 blocks in one launch): Bu and the recurrence states never leave the SM, so its "
                            "algorithmic bytes are only each block's input and output; it is bound by issue slots (CUDA-core "
                            "fake-quant arithmetic around the tcgen05 MMAs, profiles/), not by HBM")
    step_bytes = sum(k["alg_bytes_per_launch"] * k["launches_per_step"] for k in kernels.values() if k["alg_bytes_per_launch"])
    roofline["step_alg_bytes"] = step_bytes
    roofline["step_achieved_gbs"] = step_bytes / (ms_total / K * 1e-3) / 1e9

    extras = {}
    cpu_baseline = None
    scan_rf = None
    tput = None
    if rank == 0 and not args.no_extras:
        try:
            extras = secondary_flavours(torch, model, params, batch_stats, cal_stats, dims, xs, B, L, dev)
        except Exception as e:  # secondary numbers must never take the headline down
            extras = {"error": repr(e)}
        try:
            scan_rf = scan_roofline(torch, B, L, dims.ssm_size, dev, peak)
        except Exception as e:
            scan_rf = {"error": repr(e)}
        if world == 1:
            try:
                tput = throughput_batch(torch, lib, _abi, model, dims, L, dev, peak)
            except Exception as e:
                tput = {"error": repr(e)}
        if world == 1:
            # bounded sample of the same workload: 8-clip slices of the step's batch, repeated until about 10 s of CPU work
            fwd = cpu_port_forward_fn(params, batch_stats, dims)
            nclip = min(B, 8)
            xc = xs[0][:nclip].cpu().numpy()
            yc = fwd(xc)                                             # warm-up (BLAS threads, page faults)
            rel = float(np.abs(model(xs[0])[:nclip].cpu().numpy() - yc).max() / np.abs(yc).max())
            runs, t0 = 0, time.perf_counter()
            while True:
                fwd(xs[0][(runs * nclip) % max(1, B - nclip + 1):][:nclip].cpu().numpy() if B > nclip else xc)
                runs += 1
                dt = time.perf_counter() - t0
                if dt >= 10.0 or runs >= 64:
                    break
            cpu_baseline = {"value": runs * nclip * L / dt, "unit": "frames/s", "cores": host_threads(), "kind": "port",
                            "sample": "%d runs of %d of the %d clips of one step, %.1f s (NumPy port of the reference, oracle/)"
                                      % (runs, nclip, B, dt),
                            "gpu_vs_port_rel_err": rel}

    if rank == 0:
        line = {
            "metric": "S5 NDNS inference frames/sec (sparse static w8a16, dim_scale %.1f)" % args.dim_scale,
            "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": max(W, 3), "ms_per_step": ms_total / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: ndns.json dim_scale %.1f, 90%% iterative-magnitude-pruned (ERK), "
                                   "relufication, static w8a16 forward" % args.dim_scale,
                       "clips_per_gpu": B, "frames_per_clip": L, "H": dims.d_model, "P": dims.ssm_size, "n_layers": dims.n_layers,
                       "parallelism": "clips sharded over %d GPU(s), no data-path collective" % world,
                       "l2": "inputs rotate over %d batches (%.0f MB) > 126 MB L2" % (n_rot, n_rot * rows * D_IO * 4 / 1e6),
                       "weights": "random-init recipe shapes, scales from a one-batch calibration (tests/golden fixture)"},
            "clocks": clock_info,
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": rows * D_IO * 4, "d2h_bytes_per_step": rows * D_IO * 4,
                    "steps": kp, "mode": "s5b200_model_forward_host_async, pinned host buffers, two batches in flight (consecutive steps "
                                         "on two alternating caller streams), one synchronize after the last step",
                    "outputs_equal_host_synchronous": pipe_equal,
                    "host_synchronous_value": e2e_sync_fps, "stream_ordered_value": e2e_async_fps, "numa": numa,
                    "audio_to_audio": e2e_audio},
            "gpu_launches": launches,
            "roofline": roofline,
            "kernels": kernels,
            "cpu_baseline": cpu_baseline,
            "checksum": [float(chk[0].item()), float(chk[1].item())],
        }
        if extras:
            line["flavours"] = extras
        if scan_rf:
            line["scan_roofline"] = scan_rf
        if tput:
            line["throughput_batch"] = tput
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


def kernel_breakdown(fn, reps=3):
    """ms per step of every kernel class the library launched while running fn() (CUDA events per launch)."""
    from sparsernns_b200 import _abi

    lib = _abi.lib()
    lib.s5b200_profile_begin()
    for _ in range(reps):
        fn()
    ms = (C.c_double * len(_abi.KERNEL_CLASSES))()
    n = (C.c_longlong * len(_abi.KERNEL_CLASSES))()
    _abi.check(lib.s5b200_profile_end(ms, n))
    return {name: round(ms[i] / reps, 4) for i, name in enumerate(_abi.KERNEL_CLASSES) if n[i]}


def throughput_batch(torch, lib, _abi, model, dims, L, dev, peak, clips=256, steps=5):
    """The same forward at the throughput batch SURVEY section 8d names for weak scaling (256 clips per GPU): the 40-60 us launches of
    the 32-clip step are dominated by ramp and tail, here the memory-bound kernels show their steady-state fraction of the roofline."""
    import ctypes as C

    free, _ = torch.cuda.mem_get_info()
    per_clip = L * D_IO * 4
    if free < 6 * clips * per_clip:
        return {"skipped": "not enough free device memory"}
    xs = synth_inputs(torch, 2, clips, L, 4242, dev)              # 2 x 987 MB: every step reads an input that is not in L2
    rows = clips * L
    for i in range(2):
        y = model(xs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        y = model(xs[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    lib.s5b200_profile_begin()
    for i in range(steps):
        y = model(xs[i % 2])
    ms_cls = (C.c_double * len(_abi.KERNEL_CLASSES))()
    n_cls = (C.c_longlong * len(_abi.KERNEL_CLASSES))()
    _abi.check(lib.s5b200_profile_end(ms_cls, n_cls))
    alg = algorithmic_bytes(rows, dims.d_model, dims.ssm_size)
    kern = {}
    for i, name in enumerate(_abi.KERNEL_CLASSES):
        if n_cls[i]:
            per = ms_cls[i] / n_cls[i]
            ab = alg.get(name)
            if name == "block" and ab:
                ab = ab * dims.n_layers / max(1.0, n_cls[i] / steps)
            kern[name] = {"ms_per_launch": per, "launches_per_step": n_cls[i] / steps,
                          "frac": (ab / (per * 1e-3) / 1e9 / peak) if ab else None}
    del xs, y
    torch.cuda.empty_cache()
    return {"clips": clips, "frames_per_s": rows / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "kernels": kern,
            "l2": "inputs alternate over 2 batches (%.0f MB each) > 126 MB L2" % (clips * per_clip / 1e6)}


def scan_roofline(torch, B, L, P, dev, peak):
    """The second half of BASELINE.json's metric: the stand-alone diagonal recurrence (s5b200_scan_c64, the kernel that replaces
    lax.associative_scan) against the HBM roofline, 16 bytes per state-step, on (a) the step's own shape and (b) the throughput
    batch of BASELINE configs[4].  CUDA events; inputs rotate over buffer sets larger than L2 where one set is not."""
    from sparsernns_b200 import _abi

    lib = _abi.lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {}
    for tag, (b, l) in (("step_shape", (B, L)), ("batch_4096", (4096, 1024))):
        by = 16.0 * b * l * P
        nbuf = 1 if by > 1e9 else int(min(8, max(2, -(-400e6 // by))))
        bu = [torch.randn(b, l, P, 2, device=dev) for _ in range(nbuf)]
        xo = [torch.empty_like(bu[0]) for _ in range(nbuf)]
        lam = torch.view_as_real(torch.polar(torch.rand(P, device=dev) * 0.09 + 0.9, torch.rand(P, device=dev) * 2 - 1)).contiguous()
        wsb = lib.s5b200_scan_c64_workspace_bytes(b, l, P)
        ws = [torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev) for _ in range(nbuf)]

        def run(i):
            _abi.check(lib.s5b200_scan_c64(bu[i].data_ptr(), lam.data_ptr(), xo[i].data_ptr(), b, l, P, 0, ws[i].data_ptr(), wsb, st), "scan_c64")

        for i in range(3 * nbuf):
            run(i % nbuf)
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for r in range(reps):
            for i in range(nbuf):
                run(i)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * nbuf)
        out[tag] = {"B": b, "L": l, "P": P, "us": ms * 1e3, "alg_bytes": by, "achieved_gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peak,
                    "buffers": nbuf}
        del bu, xo, ws
    return out


def secondary_flavours(torch, model, params, batch_stats, cal_stats, dims, xs, B, L, dev):
    """Device-timed frames/s of the other two numerics flavours on the same batch (informational)."""
    from sparsernns_b200 import fxputils
    from sparsernns_b200.fxpmodel import FxpRegressionModel
    from sparsernns_b200.model import BoundModel, S5Config

    out = {}

    def t(fn, n=5):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for i in range(n):
            fn(i)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    md = fxputils.load_modeldict(params, cal_stats)
    q = fxputils.create_fxp_qconfig(md, "w8a16", shared_exp=True)
    fx = FxpRegressionModel(md, q, dims.n_layers)
    ms = t(lambda i=0: fx.forward_float(xs[i % len(xs)]))
    out["fixed_point_w8a16"] = {"frames_per_s": B * L / (ms / 1e3), "ms_per_step": ms, "kernels_ms_per_step": kernel_breakdown(lambda: fx.forward_float(xs[0]))}
    cfg = S5Config(n_layers=dims.n_layers, d_model=dims.d_model, ssm_size=dims.ssm_size, relufication=False,
                   quantization="none", use_sparse=False)
    fm = BoundModel(cfg, {"params": params, "batch_stats": batch_stats})
    ms = t(lambda i=0: fm(xs[i % len(xs)]))
    out["dense_fp32_gelu"] = {"frames_per_s": B * L / (ms / 1e3), "ms_per_step": ms, "kernels_ms_per_step": kernel_breakdown(lambda: fm(xs[0]))}
    return out


if __name__ == "__main__":
    main()
