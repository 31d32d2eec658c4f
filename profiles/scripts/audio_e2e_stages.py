"""Audio-to-audio end to end (bench.py's e2e.audio_to_audio): device time of each stage and the chunk / stream sweep of the loop."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from sparsernns_b200 import ndns, synth
from sparsernns_b200.model import BoundModel, S5Config

dims = synth.ndns_dims(0.5)
params, batch_stats, _ = bench.load_fixture(0.5)
cfg = S5Config(n_layers=dims.n_layers, d_model=dims.d_model, ssm_size=dims.ssm_size, relufication=True, quantization="w8a16")
B, L = 32, 3751
n_samp = (L - 1) * 128
dev = torch.device("cuda:0")
models = [BoundModel(cfg, {"params": params, "batch_stats": batch_stats}) for _ in range(4)]
ah = [(0.05 * torch.randn(B, n_samp)).pin_memory() for _ in range(2)]
oh = [torch.empty(B, n_samp).pin_memory() for _ in range(2)]
out = []


def ev_time(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for nb in (32, 8):
    a = ah[0][:nb].to(dev)
    mag, ph, xin = ndns.stft_splitter(a, blf=True, model_input=True)
    mask = models[0](xin)
    rec = {"clips": nb,
           "h2d_ms": round(ev_time(lambda: a.copy_(ah[0][:nb], non_blocking=True)), 4),
           "stft_ms": round(ev_time(lambda: ndns.stft_splitter(a, blf=True, model_input=True)), 4),
           "model_ms": round(ev_time(lambda: models[0](xin)), 4),
           "mixer_ms": round(ev_time(lambda: ndns.stft_mixer(mag, ph, blf=True, mask=mask)), 4)}
    cl = ndns.stft_mixer(mag, ph, blf=True, mask=mask)
    rec["d2h_ms"] = round(ev_time(lambda: oh[0][:nb].copy_(cl, non_blocking=True)), 4)
    print(json.dumps(rec), flush=True)
    out.append(rec)
    del a, mag, ph, xin, mask, cl

K = 10
for nch, nst in ((4, 2), (2, 2), (8, 2), (4, 4), (8, 4), (16, 4), (8, 3), (1, 1)):
    streams = [torch.cuda.Stream() for _ in range(nst)]

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
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        audio_step(i)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / K * 1e3
    rec = {"chunks": nch, "streams": nst, "ms_per_step": round(ms, 3), "frames_per_s": round(B * L / ms * 1e3)}
    print(json.dumps(rec), flush=True)
    out.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/audio_e2e_stages.json", "w"), indent=1)
