#!/usr/bin/env python3
"""bench.py -- localized frames/s of the per-frame localization path (BASELINE.json metric).

  python bench.py [--gpus N --steps K --warmup W]          our arm (CUDA through the C ABI)
  python bench.py --impl reference [...]                   the reference's own C path on host cores

Workload (N=1): BASELINE.json configs[1] -- 2^20 synthetic frames, reference geometry (3 mics x
1024 samples, +-46 lags, 50 kHz), fixed-point direct cross-correlation, outputs = 3 TDOA lags +
likelihood-map cell + plane coordinates per frame.  N>1: every rank holds its own 2^20-frame
slice of a batch N times larger (weak scaling, contiguous frame ranges, no data-path collective);
every rank's kernel stores its 24 B/frame of results straight into rank 0's result arrays
(at_shared_alloc / at_shared_open: CUDA-IPC mapped peer memory, the only bytes that cross NVLink),
inside the timed region.
A step = one pass of the hot path over the whole batch.  The input (3.2 GB per GPU) is far larger
than L2 (126 MB), so no explicit L2 flush is needed between timed iterations.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAC_PER_FRAME = 279210          # 3 pairs x sum_{s=-46..46} (1024 - |s|)
BYTES_IN_PER_FRAME = 3 * 1024   # uint8 ADC bytes
FRAMES_DEFAULT = 1 << 20
WANT = ("lags", "cell", "xy")
BYTES_OUT_PER_FRAME = 3 * 4 + 4 + 8
SYNTH_KATS, SYNTH_RANDOM_HEADS, SYNTH_MAX_NOISE, SYNTH_WHITE = 4, 2, 8, 16
WORKLOAD = ("BASELINE configs[1]: 2^20 synthetic frames per GPU, reference geometry (3 mics x 1024 samples, "
            "+-46 lags, 50 kHz), fixed-point direct xcorr, outputs lags+cell+xy")


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler(threading.Thread):
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.mask, self.max_mhz = index, [], 0, None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while self.nv and not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                fn = getattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                    self.nv.nvmlDeviceGetCurrentClocksThrottleReasons
                self.mask |= int(fn(self.h))
            except Exception:
                pass
            time.sleep(0.01)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=1.0)
        reasons = [name for bit, name in self.REASONS.items() if self.mask & bit]
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.samples)}


# --------------------------------------------------------------------------- reference arm / cpu baseline
def reference_lib():
    """The CPU checker: the reference's own objects (oracle/_ref) when they were built, else the restated oracle.
    Never touches the product library."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_bindings import Oracle, load_ref
    ref = load_ref(fast=True)
    port = Oracle()
    return ("reference" if ref is not None else "port"), ref, port


def reference_run(kind, ref, port, adc, nthreads):
    """Steps a9-a15 of SURVEY 8a for every frame of `adc` on `nthreads` host threads. Returns (seconds, lags)."""
    lags = np.zeros((adc.shape[0], 3), np.int32)
    t0 = time.perf_counter()
    if kind == "reference":
        ref.ref_localize_frames(adc.reshape(-1), adc.shape[0], lags.ctypes.data, None, nthreads, 0)
    else:
        lags = port.localize(adc, want_corr=False, want_cell=False, nthreads=nthreads)["lags"]
    return time.perf_counter() - t0, lags


def run_reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    kind, ref, port = reference_lib()
    cores = os.cpu_count() or 1
    F = args.frames
    # The same batch as our arm -- frames [0, F) of the generator -- built on the host by oracle/synth_host.cpp
    # (the generator header, compiled into the checker), so this arm maps nothing but oracle/ and oracle/_ref.
    t0 = time.perf_counter()
    adc, _, _ = port.synth(F, flags=SYNTH_KATS, first_frame=0, nthreads=cores)
    gen_s = time.perf_counter() - t0
    for _ in range(min(args.warmup, 1)):      # one untimed pass warms the page cache / thread pool; more buys nothing on a CPU
        reference_run(kind, ref, port, adc, cores)
    t = 0.0
    for _ in range(args.steps):
        dt, _ = reference_run(kind, ref, port, adc, cores)
        t += dt
    value = F * args.steps / t
    line = {"impl": "reference", "metric": "localized frames/sec", "value": value, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16 x int16 -> int64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": F, "global_frames": F,
                       "note": "the reference's own sample_compute steps (write_out, <<8, window, 3 x correlations_init) on "
                               "all host threads; one step = the whole 2^20-frame batch"},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind,
                             "sample": "frames [0, %d) of the bench batch per step (generated on the host in %.1f s by "
                                       "oracle/synth_host.cpp); reference objects -O3 x86-64-v3, one pthread per core"
                                       % (F, gen_s)},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------- our arm
def time_device(loc, torch, stream, adc, heads, want, steps, out=None, struct_corr=False):
    """Device-resident throughput of one output selection: frames/s over `steps` launches (3 warm-up launches)."""
    out = {} if out is None else out
    for _ in range(3):
        loc.localize_device(adc, heads, want=want, out=out, struct_corr=struct_corr)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(steps):
        loc.localize_device(adc, heads, want=want, out=out, struct_corr=struct_corr)
    b.record(stream)
    torch.cuda.synchronize()
    return adc.shape[0] * steps / (a.elapsed_time(b) * 1e-3)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import audio_triangulation_b200 as at

    rank, local_rank, world = dist_env()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    F = args.frames
    loc = at.Localizer(device=local_rank, kernel=args.kernel)
    stream = torch.cuda.current_stream(dev)

    # synthetic batch, resident in HBM: this rank's contiguous frame range of the global batch
    from audio_triangulation_b200.sharding import frame_range
    lo, hi = frame_range(rank, world, world * F)
    assert hi - lo == F
    adc, _, _ = loc.synth_device(F, flags=SYNTH_KATS if rank == 0 else 0, first_frame=lo)

    # Result arrays.  N > 1: rank 0 owns arrays for the GLOBAL batch (at_shared_alloc) and every other rank maps them
    # (at_shared_open, CUDA IPC); each rank's kernel stores its slice there directly, so the "gather" is the epilogue's
    # own 24 B/frame of stores over NVLink -- no collective, no copy kernel on the SMs.
    shapes = {"lags": ((F, 3), torch.int32, 12), "cell": ((F,), torch.int32, 4), "xy": ((F, 2), torch.float32, 8)}
    shared = None
    if world > 1:
        nbytes = world * F * BYTES_OUT_PER_FRAME
        handle = [None]
        if rank == 0:
            shared = loc.shared_alloc(nbytes)
            handle = [shared.handle]
        dist.broadcast_object_list(handle, src=0)
        if rank != 0:
            shared = loc.shared_open(handle[0], nbytes)
        off, base = {}, 0
        for k, (_, _, bpf) in shapes.items():          # [lags of all ranks][cell of all ranks][xy of all ranks]
            off[k] = base
            base += world * F * bpf
        remote = {k: shared.view(off[k] + rank * F * bpf, F * bpf) for k, (_, _, bpf) in shapes.items()}
        out = remote
        if rank == 0:
            whole = shared.tensor()
            glob = {k: whole[off[k]: off[k] + world * F * bpf].view(d).view((world,) + sh) for k, (sh, d, bpf) in shapes.items()}
            out = {k: glob[k][0] for k in shapes}
    else:
        out = {k: torch.empty(sh, dtype=d, device=dev) for k, (sh, d, _) in shapes.items()}
    # --gather copy: the kernel stores into local double buffers and a side stream moves each step's 24 MB slice to rank 0
    # with the copy engine while the next step computes (no small store packets on NVLink)
    staged = world > 1 and rank != 0 and args.gather == "copy"
    if staged:
        bufs = [{k: torch.empty(sh, dtype=d, device=dev) for k, (sh, d, _) in shapes.items()} for _ in range(2)]
        side = torch.cuda.Stream(dev)
        done = [torch.cuda.Event(), torch.cuda.Event()]
    nstep = [0]

    def step():
        if not staged:
            loc.localize_device(adc, None, want=WANT, out=out)
            return
        b = nstep[0] & 1
        nstep[0] += 1
        stream.wait_event(done[b])                      # the copy that last read this buffer
        loc.localize_device(adc, None, want=WANT, out=bufs[b])
        ready = torch.cuda.Event()
        ready.record(stream)
        side.wait_event(ready)
        for k, (_, _, bpf) in shapes.items():
            loc.copy_async(remote[k], bufs[b][k], F * bpf, side)
        done[b].record(side)

    def barrier():
        if staged:
            stream.wait_stream(side)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = loc.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    if staged:
        stream.wait_stream(side)                        # the last slice has arrived at rank 0 inside the timed region
    ev1.record(stream)
    barrier()
    launches = loc.kernel_launches() - launches0
    clocks = sampler.finish()
    total_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        mine = torch.tensor([total_ms / args.steps, clocks["sm_mhz"] or 0.0, float("sw_power_cap" in clocks["reasons"])],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [round(x[0].item(), 4) for x in allr], "sm_mhz": [x[1].item() for x in allr],
                    "sw_power_cap": [bool(x[2].item()) for x in allr]}
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = t.item()
    kern_ms = total_ms / args.steps          # the step IS the kernel (certified pass + exact pass over its short list)
    value = world * F * args.steps / (total_ms * 1e-3)

    # ---- (untimed) are the sharded results byte-identical to a single-GPU computation of the same frame ranges?
    sharded_ok = None
    if world > 1:
        barrier()
        if rank == 0:
            sharded_ok = True
            chk = {}
            for r in range(world):
                rlo, _ = frame_range(r, world, world * F)
                radc, _, _ = loc.synth_device(F, flags=SYNTH_KATS if r == 0 else 0, first_frame=rlo)
                loc.localize_device(radc, None, want=WANT, out=chk)
                torch.cuda.synchronize(dev)
                for k in WANT:
                    sharded_ok = sharded_ok and bool(torch.equal(chk[k].view(torch.uint8), glob[k][r].view(torch.uint8)))
                del radc
        barrier()

    # ---- how the frames of this batch were settled (untimed extra pass)
    search = None
    try:
        st = loc.localize_device(adc, None, want=WANT + ("stats",))["stats"]
        torch.cuda.synchronize(dev)
        st = st.cpu().numpy().astype(float)
        if st[:4].sum() > 0:
            tot = st[:4].sum()
            search = {"peak_tuple_lookup": st[3] / tot, "first_box": st[0] / tot, "widened_box": st[1] / tot,
                      "full_scan": st[2] / tot, "lags_certified_without_ll_product": st[4] / tot}
    except Exception:
        pass

    # ---- end to end through the host API: pinned host frames -> H2D -> kernels -> D2H results, every step
    pinned = torch.empty((F, 3, 1024), dtype=torch.uint8).pin_memory()
    pinned.copy_(adc, non_blocking=False)
    hout = {"lags": torch.empty((F, 3), dtype=torch.int32).pin_memory(),
            "cell": torch.empty((F,), dtype=torch.int32).pin_memory(),
            "xy": torch.empty((F, 2), dtype=torch.float32).pin_memory()}
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(2):
        loc.localize_host(pinned, want=WANT, out=hout)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        loc.localize_host(pinned, want=WANT, out=hout)   # returns when the results are in host memory
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * F * e2e_steps / te.item()
    e2e_ok = None
    if rank == 0:
        e2e_ok = all(bool((hout[k].numpy() == out[k].cpu().numpy()).all()) for k in WANT)
    h2d_gbs = world * F * BYTES_IN_PER_FRAME * e2e_steps / te.item() / 1e9
    # the same with the reference's complete per-frame product out: whole correlations_t structs (post-Gaussian curves,
    # lags, time stamp; 2 280 B/frame D2H next to the 3 072 B/frame H2D) -- what the reference arm computes per frame
    e2e_struct = None
    if world == 1 and not args.no_extras:
        Fs = min(F, 1 << 18)
        hs = {"lags": torch.empty((Fs, 3), dtype=torch.int32).pin_memory(),
              "corr": torch.empty((Fs, 3, 95), dtype=torch.int64).pin_memory()}
        loc.localize_host(pinned[:Fs], want=("lags", "corr"), out=hs, struct_corr=True)
        t0 = time.perf_counter()
        for _ in range(3):
            loc.localize_host(pinned[:Fs], want=("lags", "corr"), out=hs, struct_corr=True)
        dt = time.perf_counter() - t0
        e2e_struct = {"value": Fs * 3 / dt, "unit": "frames/s", "frames": Fs, "h2d_bytes_per_step": Fs * BYTES_IN_PER_FRAME,
                      "d2h_bytes_per_step": Fs * (3 * 95 * 8 + 12), "lags_match": bool((hs["lags"].numpy() == out["lags"][:Fs].cpu().numpy()).all())}
        del hs
    del pinned

    line = None
    if rank == 0:
        # ---- the whole truth about the path (N = 1): other output selections, other inputs (fewer steps each)
        extra = {}
        if world == 1 and not args.no_extras:
            k = max(2, min(args.steps, args.extra_steps))
            modes = {"lags_cell_xy": value,
                     "lags_only": time_device(loc, torch, stream, adc, None, ("lags",), k)}
            Fc = min(F, 1 << 18)     # whole correlations_t structs: 2 280 B/frame of output
            modes["full_corr_struct"] = time_device(loc, torch, stream, adc[:Fc], None, ("lags", "corr"), k, struct_corr=True)
            modes["full_corr_struct_frames"] = Fc
            extra["modes"] = modes
            extra["e2e_full_corr_struct"] = e2e_struct
            worst = {}
            for name, flags in (("white_noise", SYNTH_WHITE), ("lowest_snr", SYNTH_MAX_NOISE)):
                wloc = at.Localizer(device=local_rank, kernel=args.kernel)    # its own context: its own launch history
                wadc, _, _ = wloc.synth_device(F, flags=flags, first_frame=0)
                torch.cuda.synchronize(dev)
                worst[name] = time_device(wloc, torch, stream, wadc, None, WANT, k)
                st = wloc.localize_device(wadc, None, want=WANT + ("stats",))["stats"].cpu().numpy().astype(float)
                worst[name + "_certified_frac"] = float(st[4] / F)
                del wadc
                wloc.close()
            extra["worst_case"] = worst
            hadc, hheads, _ = loc.synth_device(F, flags=SYNTH_KATS | SYNTH_RANDOM_HEADS, first_frame=0)
            torch.cuda.synchronize(dev)
            extra["random_ring_heads"] = time_device(loc, torch, stream, hadc, hheads, WANT, k)
            del hadc, hheads
            extra["certified_frac"] = (search or {}).get("lags_certified_without_ll_product")
            # BASELINE configs[3] and [4] (no reference counterpart; parity against the generalised oracle only)
            try:
                loc8 = at.Localizer(device=local_rank, n_mics=8, n_bits=12, max_shift=46, points=at.hemisphere_points(72, 12, 2.0))
                F8 = 1 << 14                       # 512 MB of input: larger than L2
                adc8, _, _ = loc8.synth_device(F8)
                torch.cuda.synchronize(dev)
                extra["config4"] = {"workload": "8 mics x 4096 samples, 28 pairs, +-46 lags, direct fixed-point xcorr on tcgen05; "
                                                "position = arg-max over 72 x 12 hemisphere directions",
                                    "frames": F8,
                                    "lags_frames_per_s": time_device(loc8, torch, stream, adc8, None, ("lags",), k),
                                    "lags_position_frames_per_s": time_device(loc8, torch, stream, adc8, None, ("lags", "cell", "xy"), k),
                                    "int16_tmac_per_s": None}
                extra["config4"]["int16_tmac_per_s"] = extra["config4"]["lags_frames_per_s"] * 28 * (93 * 4096 - 2162) / 1e12
                # the hand-written FFT / GCC-PHAT variant on the same batch (another statistic: lags agree, curves do not):
                # forward radix-8 FFTs + one tcgen05 fp16 contraction over the +-46 lags; its cost does not depend on the lag
                # range up to +-63, the direct form's does (one 128-row tile covers 93 + 16 lags)
                g8 = loc8.gccphat_device(adc8)
                torch.cuda.synchronize(dev)
                t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0e.record()
                for _ in range(k):
                    loc8.gccphat_device(adc8)
                t1e.record()
                torch.cuda.synchronize(dev)
                gps = F8 * k / (t0e.elapsed_time(t1e) * 1e-3)
                d8 = loc8.localize_device(adc8, None, want=("lags",))["lags"]
                extra["config4"]["gccphat_frames_per_s"] = gps
                extra["config4"]["gccphat_lag_agreement_with_direct"] = float((g8 == d8).float().mean().item())
                extra["config4"]["direct_over_gccphat"] = extra["config4"]["lags_frames_per_s"] / gps
                del adc8
                loc8.close()
            except Exception as e:   # pragma: no cover
                extra["config4"] = {"error": str(e)}
            try:
                import subprocess
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stream_bench.py"), "--blocks", "20"],
                                   capture_output=True, text=True, timeout=240)
                extra["config5"] = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception as e:   # pragma: no cover
                extra["config5"] = {"error": str(e)}

        # ---- roofline of the dominant kernel
        ubench = {}
        for name in ("umma_i8", "umma_frame", "imma_s8", "imad_wide", "imad", "lds"):
            try:
                g, _ = loc.microbench(name)
                ubench[name] = g
            except Exception as e:   # pragma: no cover
                ubench[name] = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_ach = (BYTES_IN_PER_FRAME + BYTES_OUT_PER_FRAME) * F / (kern_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("dram_bytes_per_frame") * F
        except Exception:
            pass
        tc = args.kernel in ("auto", "umma")
        # auto / umma: the 279 210 int16 MACs of a frame are 4 x 279 210 int8 MACs on the tcgen05 tensor cores (polyphase
        # Hankel MMAs); peak = dense tcgen05 kind::i8 rate measured live on this GPU (MEASURED_PEAKS.json has no int8
        # figure).  imma: legacy mma.sync int8 rate; imad: IMAD.WIDE chain rate.
        peak_name = "umma_i8" if tc else ("imma_s8" if args.kernel == "imma" else "imad_wide")
        peak = (ubench.get(peak_name) or 0.0) / 1e3
        per_frame = MAC_PER_FRAME * (1 if args.kernel == "imad" else 4)
        achieved = per_frame * F / (kern_ms * 1e-3) / 1e12
        certified = (search or {}).get("lags_certified_without_ll_product") or 0.0
        # MACs put on the tensor core per frame: the certified pass issues N = 64 + 32 + 32 + 16 columns per K-step (two
        # K-steps, M = 128, K = 32); the frames it cannot settle get the exact pass (N = 192 per K-step) on top
        issued = 2 * 144 * 128 * 32 + (1.0 - certified) * 2 * 192 * 128 * 32
        seq = ubench.get("umma_frame")
        roof = {"bound": "tensor" if args.kernel != "imad" else "int-pipe",
                "achieved": achieved, "peak": peak,
                "unit": "T int8-MAC/s (4 per int16 MAC, useful lags only)" if args.kernel != "imad" else "T int16-MAC/s",
                "frac": achieved / peak if peak else None, "traffic": traffic,
                "peak_source": "measured live: at_microbench(%s) on this GPU" % peak_name,
                # what the formulation can draw from the tensor core: a frame is eight M128 x N<=64 x K32 MMAs whose
                # Hankel A operand (4 KB each, re-fetched from shared memory per MMA) bounds them, not the MAC array
                "tensor_sequence_frames_per_s": seq * 1e9 if seq else None,
                "frac_of_tensor_sequence": (F / (kern_ms * 1e-3)) / (seq * 1e9) if seq else None,
                "issued_int8_mac_per_frame": issued if tc else None,
                "int16_tmac_per_s": MAC_PER_FRAME * F / (kern_ms * 1e-3) / 1e12,
                "kernel_ms": kern_ms, "algorithmic_mac_per_frame": MAC_PER_FRAME,
                "hbm_achieved_gbs": hbm_ach, "hbm_peak_gbs": hbm_peak, "hbm_frac": hbm_ach / hbm_peak,
                "hbm_peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                "microbench_gops": ubench}
        # the north-star's other denominators: the integer pipe (one IMAD = one int16 MAC) -- what the best CUDA-core
        # kernel could reach -- and, from the committed ncu capture, the pipe that actually bounds the tcgen05 kernel
        if ubench.get("imad"):
            roof["int_pipe_peak_tmac_per_s"] = ubench["imad"] / 1e3
            roof["frac_of_int_pipe_roofline"] = roof["int16_tmac_per_s"] / (ubench["imad"] / 1e3)
        if tc:
            try:
                cap = json.load(open(os.path.join(ROOT, "profiles", "r2_umma_full.json")))[0]
                tcw = cap["l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]["value"]
                lsw = cap["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]["value"]
                roof["limiter"] = {"pipe": "shared-memory data pipe (tensor-core operand fetch + LSU)", "busy_frac": (tcw + lsw) / 100.0,
                                   "tensor_operand_fetch_frac": tcw / 100.0, "lsu_frac": lsw / 100.0,
                                   "source": "profiles/r2_umma_full.json (ncu --set full of this kernel, not measured in this run)"}
            except Exception:
                pass
        line = {"metric": "localized frames/sec", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int16 x int16 -> int64, computed as 4 x (int8 x int8 -> int32) on tensor cores (u8 ADC in)", "data": "synthetic",
                "config": {"workload": WORKLOAD, "frames_per_gpu": F, "global_frames": world * F,
                           "kernel": args.kernel, "likelihood_search": search, "l2": "inputs (3.2 GB/GPU) larger than L2, no flush",
                           "sharding": (("contiguous frame ranges; every rank's kernel stores its 24 B/frame straight into rank 0's "
                                         "arrays (CUDA-IPC peer memory over NVLink)") if args.gather == "stores" else
                                        ("contiguous frame ranges; every rank's 24 B/frame go to rank 0's arrays (CUDA-IPC peer memory) "
                                         "by copy-engine peer copies overlapped with the next step")) if world > 1 else "single GPU"},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": F * BYTES_IN_PER_FRAME,
                        "d2h_bytes_per_step": F * BYTES_OUT_PER_FRAME, "steps": e2e_steps, "matches_device_path": e2e_ok,
                        "aggregate_h2d_gbs": h2d_gbs},
                "roofline": roof}
        if world > 1:
            line["sharded_bytes_identical"] = sharded_ok
            line["per_rank"] = per_rank
        line.update(extra)

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference's own objects on the host cores
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind, ref, port = reference_lib()
        cores = os.cpu_count() or 1
        probe = adc[:2048].cpu().numpy()
        dt, _ = reference_run(kind, ref, port, probe, cores)
        n = int(max(2048, min(F, args.cpu_seconds * 2048 / max(dt, 1e-6))))
        sample = adc[:n].cpu().numpy()
        dt, ref_lags = reference_run(kind, ref, port, sample, cores)
        mism = int((ref_lags != out["lags"][:n].cpu().numpy()).any(1).sum())
        dt1, _ = reference_run(kind, ref, port, sample[: max(1024, n // cores)], 1)
        line["cpu_baseline"] = {"value": n / dt, "unit": "frames/s", "cores": cores, "kind": kind,
                                "sample": "first %d frames of the same batch, all %d host threads; single-thread: %.0f frames/s"
                                          % (n, cores, max(1024, n // cores) / dt1),
                                "lag_mismatches_vs_gpu": mism}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        if rank != 0:
            shared.close()
        dist.barrier()
        if rank == 0:
            del out, glob, whole
            shared.close()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The one JSON line, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries (NCCL prints "NCCL version ..." to stdout) must not pollute the one-line contract: route fd 1 to
    # stderr for the whole run and keep the original stdout for emit().
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_DEFAULT, help="frames per GPU per step")
    ap.add_argument("--kernel", default="auto", choices=["auto", "imad", "imma", "umma"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--extra-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--gather", default="stores", choices=["stores", "copy"],
                    help="N > 1: kernels store results straight into rank 0's arrays, or local buffers + copy-engine peer copies")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
