#!/usr/bin/env python3
"""bench.py -- localized frames/s of the per-frame localization path (BASELINE.json metric).

  python bench.py [--gpus N --steps K --warmup W]          our arm (CUDA through the C ABI)
  python bench.py --impl reference [...]                   the reference's own C path on host cores

Workload (N=1): BASELINE.json configs[1] -- 2^20 synthetic frames, reference geometry (3 mics x
1024 samples, +-46 lags, 50 kHz), fixed-point direct cross-correlation, outputs = 3 TDOA lags +
likelihood-map cell + plane coordinates per frame.  N>1: every rank holds its own 2^20-frame
slice of a batch N times larger (weak scaling, contiguous frame ranges, no data-path collective)
and the per-frame results are gathered to rank 0 over NCCL inside the timed region.
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


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers of the
    end-to-end leg are allocated next to the GPU's PCIe root (matters from 4 ranks up).  Best effort; returns the node."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) > 4:
            bus = bus[-12:]                                   # nvml reports an 8-digit PCI domain, sysfs uses 4
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


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
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_bindings import Oracle, load_ref
    ref = load_ref(fast=True)
    if ref is not None:
        return "reference", ref, None
    return "port", None, Oracle()


def reference_run(kind, ref, port, adc, nthreads):
    """Steps a9-a15 of SURVEY 8a for every frame of `adc` on `nthreads` host threads. Returns (seconds, lags)."""
    lags = np.zeros((adc.shape[0], 3), np.int32)
    t0 = time.perf_counter()
    if kind == "reference":
        ref.ref_localize_frames(adc.reshape(-1), adc.shape[0], lags.ctypes.data, None, nthreads, 0)
    else:
        lags = port.localize(adc, want_corr=False, want_cell=False, nthreads=nthreads)["lags"]
    return time.perf_counter() - t0, lags


def sample_frames(n):
    """Frames of the bench workload for the CPU legs: the product's generator when a GPU is there
    (identical bytes to the GPU arm), numpy bursts otherwise."""
    try:
        import torch
        if torch.cuda.is_available():
            import audio_triangulation_b200 as at
            loc = at.Localizer(device=0)
            adc, _, _ = loc.synth_device(n, flags=4)
            torch.cuda.synchronize()
            out = adc.cpu().numpy()
            loc.close()
            return out, "at_synth_device frames [0,%d) of the bench batch" % n
    except Exception:
        pass
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from frames import burst_frames
    return burst_frames(n, seed=1)[0], "numpy burst frames (no GPU for the product generator)"


def run_reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    kind, ref, port = reference_lib()
    cores = os.cpu_count() or 1
    # size one step to roughly 2 s of all-core work
    probe, how = sample_frames(2048)
    dt, _ = reference_run(kind, ref, port, probe, cores)
    per_step = int(max(2048, min(1 << 18, 2.0 * 2048 / max(dt, 1e-6))))
    adc, how = sample_frames(per_step)
    for _ in range(args.warmup):
        reference_run(kind, ref, port, adc, cores)
    t = 0.0
    for _ in range(args.steps):
        dt, _ = reference_run(kind, ref, port, adc, cores)
        t += dt
    value = per_step * args.steps / t
    line = {"impl": "reference", "metric": "localized frames/sec", "value": value, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16 x int16 -> int64",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: reference geometry 3 mics x 1024 samples, +-46 lags; "
                                   "bounded sample of %d frames per step" % per_step, "frames_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind,
                             "sample": "%d frames/step, %s; reference objects -O3 x86-64-v3, one pthread per core" % (per_step, how)},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import audio_triangulation_b200 as at

    rank, local_rank, world = dist_env()
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
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
    adc, _, _ = loc.synth_device(F, flags=4 if rank == 0 else 0, first_frame=lo)
    outs = [{}, {}]            # double-buffered results: the gather of step i overlaps the kernel of step i+1
    out = outs[0]
    gathered = None
    comm = torch.cuda.Stream(dev) if world > 1 else None
    if world > 1 and rank == 0:
        gathered = [torch.empty((F, 3), dtype=torch.int32, device=dev) for _ in range(world)]

    def step(i):
        o = outs[i & 1]
        loc.localize_device(adc, None, want=WANT, out=o)
        if world > 1:   # the only bytes that cross NVLink: 12 B of lags per frame to rank 0, on a side stream
            ready = torch.cuda.Event()
            ready.record(stream)
            comm.wait_event(ready)
            with torch.cuda.stream(comm):
                dist.gather(o["lags"], gathered, dst=0)

    def barrier():
        if comm is not None:
            stream.wait_stream(comm)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = loc.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0.record(stream)
    for i in range(args.steps):
        kev[i][0].record(stream)
        step(i)
        kev[i][1].record(stream)          # main stream: brackets the localization kernel only
    if comm is not None:
        stream.wait_stream(comm)          # the last gather is inside the timed region
    ev1.record(stream)
    barrier()
    launches = loc.kernel_launches() - launches0
    clocks = sampler.finish()
    total_ms = ev0.elapsed_time(ev1)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    t = torch.tensor([total_ms, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms = t.tolist()
    value = world * F * args.steps / (total_ms * 1e-3)

    # ---- how the (exact) bounded likelihood search resolved the frames of this batch (untimed extra pass)
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
    e2e_ok = bool((hout["lags"].numpy() == out["lags"].cpu().numpy()).all())

    # ---- roofline of the dominant (only) kernel
    kernel_used = args.kernel
    ubench = {}
    if rank == 0:
        for name in ("imma_s8", "imad_wide", "imad", "dp2a", "lds"):
            try:
                g, mhz = loc.microbench(name)
                ubench[name] = {"gops": g}
            except Exception as e:   # pragma: no cover
                ubench[name] = {"error": str(e)}
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

    line = None
    if rank == 0:
        # The auto / imma kernel computes the 279,210 int16 MACs of a frame as 4 x 279,210 int8 MACs on the
        # tensor cores (byte-split Toeplitz x Hankel tiles); padded to whole 16x8x32 tiles that is 396 IMMA =
        # 1,622,016 int8 MACs -- of which the l.l digit product (99 IMMA) is issued only for the frames whose
        # arg-max the other nine products cannot certify.  Peak = legacy mma.sync int8 rate measured live on this GPU
        # (MEASURED_PEAKS.json carries no int8 figure; its bf16 number is the tcgen05 path this kernel cannot use,
        # see DESIGN.md).  The imad kernel is measured against the IMAD.WIDE chain rate instead.
        imma = kernel_used in ("auto", "imma", "imma_lm")
        peak_name = "imma_s8" if imma else "imad_wide"
        peak = ubench.get(peak_name, {}).get("gops", 0.0) / 1e3
        per_frame = MAC_PER_FRAME * (4 if imma else 1)
        certified = (search or {}).get("lags_certified_without_ll_product", 0.0) if kernel_used in ("auto", "imma") else 0.0
        issued_per_frame = 4096 * 33 * (9 + 3 * (1.0 - certified))          # int8 MACs on the tensor pipe per frame
        achieved = per_frame * F / (kern_ms * 1e-3) / 1e12
        roof = {"bound": "tensor" if imma else "int-pipe",
                "achieved": achieved, "peak": peak,
                "unit": "T int8-MAC/s (4 per int16 MAC, useful lags only)" if imma else "T int16-MAC/s",
                "frac": achieved / peak if peak else None, "traffic": traffic,
                "peak_source": "measured live: at_microbench(%s) on this GPU" % peak_name,
                "issued_frac": (achieved * issued_per_frame / per_frame / peak) if (imma and peak) else None,
                "issued_int8_mac_per_frame": issued_per_frame if imma else None,
                # the same work expressed against the INTEGER-pipe roofline the direct form would have:
                # 279,210 int16 MAC per frame vs the measured one-instruction-per-MAC rates of this GPU
                "int16_tmac_per_s": MAC_PER_FRAME * F / (kern_ms * 1e-3) / 1e12,
                "vs_int_pipe_roofline": {
                    "imad_32bit_peak": (MAC_PER_FRAME * F / (kern_ms * 1e-3) / 1e9) / ubench["imad"]["gops"] if ubench.get("imad", {}).get("gops") else None,
                    "mad_wide_chain_peak": (MAC_PER_FRAME * F / (kern_ms * 1e-3) / 1e9) / ubench["imad_wide"]["gops"] if ubench.get("imad_wide", {}).get("gops") else None},
                "kernel_ms": kern_ms, "algorithmic_mac_per_frame": MAC_PER_FRAME,
                "hbm_achieved_gbs": hbm_ach, "hbm_peak_gbs": hbm_peak, "hbm_frac": hbm_ach / hbm_peak,
                "hbm_peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                "microbench_gops": {k: v.get("gops") for k, v in ubench.items()}}
        line = {"metric": "localized frames/sec", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int16 x int16 -> int64, computed as 4 x (int8 x int8 -> int32) on tensor cores (u8 ADC in)", "data": "synthetic",
                "config": {"workload": "BASELINE configs[1]: 2^20 synthetic frames per GPU, reference geometry "
                                       "(3 mics x 1024 samples, +-46 lags, 50 kHz), fixed-point direct xcorr, "
                                       "outputs lags+cell+xy", "frames_per_gpu": F, "global_frames": world * F,
                           "kernel": kernel_used, "likelihood_search": search, "l2": "inputs (3.2 GB/GPU) larger than L2, no flush",
                           "sharding": "contiguous frame ranges, lags gathered to rank 0 over NCCL" if world > 1 else "single GPU"},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": F * BYTES_IN_PER_FRAME,
                        "d2h_bytes_per_step": F * BYTES_OUT_PER_FRAME, "steps": e2e_steps, "matches_device_path": e2e_ok,
                        "rank0_numa_node": numa_node},
                "roofline": roof}

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
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
