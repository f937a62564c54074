"""CPU, world size 2 over gloo: the frame-sharding host logic (contiguous ranges, ragged gather to rank 0).
The per-rank compute is stood in by the oracle (test infrastructure) -- what is under test is the partition and
the gather, which the GPU path shares (bench.py, NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audio_triangulation_b200.sharding import frame_range, gather_to_rank0


def test_frame_ranges_cover_the_batch():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 1000, (1 << 20) + 5):
            spans = [frame_range(r, world, n) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from frames import burst_frames
    from oracle_bindings import Oracle
    adc, _ = burst_frames(n_frames, seed=321)          # every rank derives the same global batch
    lo, hi = frame_range(rank, world, n_frames)
    local = Oracle().localize(adc[lo:hi], want_corr=False)
    lags = gather_to_rank0(torch.from_numpy(local["lags"]), n_frames, dist)
    cell = gather_to_rank0(torch.from_numpy(local["cell"]), n_frames, dist)
    if rank == 0:
        full = Oracle().localize(adc, want_corr=False)
        q.put(bool((lags.numpy() == full["lags"]).all() and (cell.numpy() == full["cell"]).all()
                   and lags.shape == (n_frames, 3)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [37, 64])
def test_two_rank_shard_and_gather_equals_single_process(n_frames):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
