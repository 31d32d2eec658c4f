"""End-to-end leg with two batches in flight (bench.py's e2e mode): chunk schedules of s5b200_model_forward_host_async swept
in one process (S5B200_E2E_CHUNKS is read per call), plus the copy-only ceiling of the same loop with plain cudaMemcpyAsync."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from sparsernns_b200 import synth
from sparsernns_b200.model import BoundModel, S5Config

dims = synth.ndns_dims(0.5)
params, batch_stats, _ = bench.load_fixture(0.5)
cfg = S5Config(n_layers=dims.n_layers, d_model=dims.d_model, ssm_size=dims.ssm_size, relufication=True, quantization="w8a16")
model = BoundModel(cfg, {"params": params, "batch_stats": batch_stats})
B, L = 32, 3751
xh = [torch.randn(B, L, 257).abs().mul_(1e-3).pin_memory() for _ in range(2)]
yh = [torch.empty_like(xh[0]).pin_memory() for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
K = 20


def run(n_streams):
    for i in range(2):
        with torch.cuda.stream(streams[i % n_streams]):
            model.forward_host(xh[i], out=yh[i], sync=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        with torch.cuda.stream(streams[i % n_streams]):
            model.forward_host(xh[i % 2], out=yh[i % 2], sync=False)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / K * 1e3


out = []
scheds = [None, "1", "2", "3", "4", "6", "8", "16", "8,8,8,8", "4,4,8,8,8", "2,6,8,8,8", "1,1,2,4,8,8,8", "6,6,6,6,4,4", "4,7,7,7,7", "2,2,4,8,8,8",
          "11,11,10", "16,16"]
for s in scheds:
    if s is None:
        os.environ.pop("S5B200_E2E_CHUNKS", None)
    else:
        os.environ["S5B200_E2E_CHUNKS"] = s
    for ns in (2, 1):
        ms = min(run(ns) for _ in range(2))
        rec = {"chunks": s or "default (1,1,2,4,6,6,6,6)", "caller_streams": ns, "ms_per_step": round(ms, 3), "frames_per_s": round(B * L / ms * 1e3)}
        print(json.dumps(rec), flush=True)
        out.append(rec)
os.environ.pop("S5B200_E2E_CHUNKS", None)

# copy-only ceiling: the same bytes with plain copies on two streams, both directions busy
d_in = [torch.empty(B, L, 257, device="cuda") for _ in range(2)]
d_out = [torch.randn(B, L, 257, device="cuda") for _ in range(2)]
for big in (True, False):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        if big:
            with torch.cuda.stream(streams[0]):
                d_in[i % 2].copy_(xh[i % 2], non_blocking=True)
            with torch.cuda.stream(streams[1]):
                yh[i % 2].copy_(d_out[i % 2], non_blocking=True)
        else:
            for c in range(4):
                sl = slice(8 * c, 8 * c + 8)
                with torch.cuda.stream(streams[0]):
                    d_in[i % 2][sl].copy_(xh[i % 2][sl], non_blocking=True)
                with torch.cuda.stream(streams[1]):
                    yh[i % 2][sl].copy_(d_out[i % 2][sl], non_blocking=True)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / K * 1e3
    rec = {"copy_only": "one copy per direction and step" if big else "four 8-clip copies per direction and step", "ms_per_step": round(ms, 3),
           "frames_per_s": round(B * L / ms * 1e3), "gbs_each_way": round(B * L * 257 * 4 / ms / 1e6, 1)}
    print(json.dumps(rec), flush=True)
    out.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/e2e_pipe_sweep.json", "w"), indent=1)
