"""BASELINE config 5 (no reference counterpart): continuous multi-array streaming at 48 kHz.
A arrays x 3 mics, blocks of B ticks pushed through the device front end (at_stream_push), captured frames
localized (at_localize_device), gated, averaged (at_average_device) and mapped (at_heatmap_device).
Reports array-ticks/s, the real-time factor at 48 kHz, and per-block latency (push -> results on the host)."""
import argparse, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audio_triangulation_b200 as at

ap = argparse.ArgumentParser()
ap.add_argument("--arrays", type=int, default=10000)
ap.add_argument("--block", type=int, default=256)
ap.add_argument("--blocks", type=int, default=40)
args = ap.parse_args()
A, B = args.arrays, args.block
loc = at.Localizer(max_shift=44, sample_rate_hz=48000.0)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
# synthetic continuous input: quiet noise, every array gets a short burst roughly every 6 blocks
def make_block(k):
    x = 128 + torch.randn((A, B, 3), device=dev, generator=g) * 1.5
    hit = (torch.arange(A, device=dev) + k) % 6 == 0
    env = torch.hann_window(B, device=dev).view(1, B, 1)
    burst = torch.randn((A, B, 1), device=dev, generator=g) * 60 * env
    x = x + burst * hit.view(A, 1, 1)
    return x.clamp(0, 255).round().to(torch.uint8).contiguous()
blocks = [make_block(k) for k in range(8)]
st = at.Stream(loc, A)
NL = loc.n_lags
est = torch.zeros((A, 3, NL), dtype=torch.int64, device=dev)
est_best = torch.zeros((A, 3), dtype=torch.int32, device=dev)
est_time = torch.zeros((A, 3), dtype=torch.int64, device=dev)
fresh = torch.zeros((A, 3, NL), dtype=torch.int64, device=dev)
gate_all = torch.zeros(A, dtype=torch.uint8, device=dev)
out = {}
lat, n_events = [], 0
def step(k):
    global n_events
    r = st.push(blocks[k % len(blocks)], out=out)
    idx = torch.nonzero(r["fired"] > 0).flatten()
    cells = None
    if idx.numel():
        res = loc.localize_device(r["frames"][idx].contiguous(), r["heads"][idx].contiguous(), want=("lags", "corr", "gate"))
        fresh[idx] = res["corr"]
        gate_all.zero_(); gate_all[idx] = res["gate"]
        loc.average_device(est, est_best, est_time, fresh, gate_all, now_us=int(1e6 * (k + 1) * B / 48000.0))
        cells = loc.heatmap_device(est[idx].contiguous(), want=("cell",))["cell"].cpu()
        n_events += int(idx.numel())
    return cells
for k in range(8): step(k)
torch.cuda.synchronize(); n_events = 0
t0 = time.perf_counter()
for k in range(args.blocks):
    t1 = time.perf_counter()
    step(8 + k)
    torch.cuda.synchronize()
    lat.append(time.perf_counter() - t1)
dt = time.perf_counter() - t0
ticks = A * B * args.blocks
print(json.dumps({"config": "48 kHz streaming, %d arrays x 3 mics, %d-tick blocks" % (A, B),
                  "array_ticks_per_s": ticks / dt, "realtime_factor_at_48kHz": ticks / dt / (A * 48000.0),
                  "arrays_sustainable_in_real_time": ticks / dt / 48000.0,
                  "block_latency_ms_p50": 1e3 * float(np.percentile(lat, 50)), "block_latency_ms_p99": 1e3 * float(np.percentile(lat, 99)),
                  "block_duration_ms_at_48kHz": 1e3 * B / 48000.0, "events_localized": n_events,
                  "kernel_launches": loc.kernel_launches()}))
