"""GPU, every visible device: frame-sharded runs must be byte-identical to a single-GPU run of the same batch
(test pyramid item 3).  With one visible GPU the in-process test shards over two contexts of that GPU
and the multi-process test runs its two ranks on that GPU (same CUDA-IPC path)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
WANT = ("lags", "cell", "xy")


def _torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def test_host_sharded_over_all_devices_equals_one_device():
    """at_localize_host_sharded with one context per visible GPU against at_localize_host on GPU 0 alone."""
    torch = _torch()
    import audio_triangulation_b200 as at
    from audio_triangulation_b200 import _lib
    n = torch.cuda.device_count()
    devs = list(range(n)) if n > 1 else [0, 0]
    locs = [at.Localizer(device=d) for d in devs]
    F = 40_003                                   # ragged over any device count
    adc, heads, _ = locs[0].synth_device(F, flags=2 | 4)
    torch.cuda.synchronize()
    adc_h, heads_h = adc.cpu().numpy(), heads.cpu().numpy()
    one = locs[0].localize_host(adc_h, heads_h, want=WANT + ("gate",))
    lags = np.zeros((F, 3), np.int32); cell = np.zeros(F, np.int32); xy = np.zeros((F, 2), np.float32); gate = np.zeros(F, np.uint8)
    o = _lib.AtOutputs(); o.lags = lags.ctypes.data; o.cell = cell.ctypes.data; o.xy = xy.ctypes.data; o.gate = gate.ctypes.data
    ctxs = (C.c_void_p * len(locs))(*[l.ctx for l in locs])
    _lib.check(locs[0].lib.at_localize_host_sharded(ctxs, len(locs), adc_h.ctypes.data, heads_h.ctypes.data, F, C.byref(o)))
    assert (lags == one["lags"]).all() and (cell == one["cell"]).all() and (gate == one["gate"]).all()
    assert (xy.view(np.uint32) == one["xy"].view(np.uint32)).all()


def _rank(rank, world, port, F, q):
    """One process per GPU, as bench.py runs it: every rank's kernel stores its slice straight into rank 0's arrays
    (CUDA-IPC mapped), rank 0 then recomputes every slice on its own GPU."""
    import torch
    import torch.distributed as dist
    import audio_triangulation_b200 as at
    from audio_triangulation_b200.sharding import frame_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    device = rank % torch.cuda.device_count()          # one GPU: both ranks share it (the IPC path is the same)
    torch.cuda.set_device(device)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # only the handle exchange and barriers: CPU collectives
    dev = torch.device("cuda", device)
    loc = at.Localizer(device=device)
    lo, hi = frame_range(rank, world, world * F)
    adc, heads, _ = loc.synth_device(F, flags=2, first_frame=lo)
    shapes = {"lags": ((F, 3), torch.int32, 12), "cell": ((F,), torch.int32, 4), "xy": ((F, 2), torch.float32, 8)}
    nbytes = world * F * 24
    handle = [None]
    if rank == 0:
        shared = loc.shared_alloc(nbytes)
        handle = [shared.handle]
    dist.broadcast_object_list(handle, src=0)
    if rank != 0:
        shared = loc.shared_open(handle[0], nbytes)
    off, base = {}, 0
    for k, (_, _, bpf) in shapes.items():
        off[k] = base
        base += world * F * bpf
    out = {k: shared.view(off[k] + rank * F * bpf, F * bpf) for k, (_, _, bpf) in shapes.items()}
    glob = None
    if rank == 0:
        whole = shared.tensor()
        glob = {k: whole[off[k]: off[k] + world * F * bpf].view(d).view((world,) + sh) for k, (sh, d, bpf) in shapes.items()}
    loc.localize_device(adc, heads, want=WANT, out=out)
    torch.cuda.synchronize(dev)
    dist.barrier()
    if rank == 0:
        ok = True
        for r in range(world):
            rlo, _ = frame_range(r, world, world * F)
            radc, rheads, _ = loc.synth_device(F, flags=2, first_frame=rlo)
            chk = loc.localize_device(radc, rheads, want=WANT)
            torch.cuda.synchronize(dev)
            ok = ok and all(bool(torch.equal(chk[k].view(torch.uint8), glob[k][r].view(torch.uint8))) for k in WANT)
        q.put(ok)
    dist.barrier()
    if rank != 0:
        shared.close()
    dist.barrier()
    if rank == 0:
        del glob
        shared.close()
    dist.destroy_process_group()


def test_peer_stores_into_rank0_equal_single_gpu():
    torch = _torch()
    world = max(2, min(torch.cuda.device_count(), 8))
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank, args=(r, world, port, 30_011, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok
