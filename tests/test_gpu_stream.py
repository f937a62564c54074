"""GPU: the device streaming front end (at_stream_*) against the oracle's capture loop
(ato_capture = rolling_buffer_push + onset gate of sample_compute.h:55-99, pinned to the reference
objects by tests/test_oracle_golden.py::test_capture_gate)."""
import numpy as np
import pytest

from oracle_bindings import N, Oracle

pytestmark = pytest.mark.gpu


def make_streams(n_arrays, n_ticks, seed):
    rng = np.random.default_rng(seed)
    out = np.empty((n_arrays, n_ticks, 3), np.uint8)
    for a in range(n_arrays):
        x = 128 + rng.normal(0, 1.2 + (a % 3), (n_ticks, 3))
        t = 1200 + int(rng.integers(0, 600))
        while t + 800 < n_ticks:                      # bursts at random times, different per-mic delays
            burst = np.convolve(rng.normal(0, 40 + 10 * (a % 5), 600), np.ones(3) / 3, "same") * np.hanning(600)
            for m, d in enumerate(rng.integers(-12, 13, 3)):
                x[t + d:t + d + 600, m] += burst
            t += int(rng.integers(1500, 4000))
        out[a] = np.clip(np.round(x), 0, 255).astype(np.uint8)
    return out


def oracle_onsets(oracle, stream):
    """All onsets of one array's recorded stream: list of (global 1-based tick, head, ring[3][N])."""
    res, pos = [], 0
    while pos < stream.shape[0]:
        rings = np.zeros((3, N), np.int16)
        head = np.zeros(1, np.int32)
        fired = oracle.lib.ato_capture(np.ascontiguousarray(stream[pos:]).reshape(-1), stream.shape[0] - pos, 3, 10,
                                       rings.reshape(-1), head)
        if fired < 0:
            break
        res.append((pos + fired, int(head[0]), rings.astype(np.uint8)))
        pos += fired
    return res


@pytest.mark.parametrize("block", [256, 1024, 16])
def test_stream_push_matches_capture_loop(loc, oracle, block):
    import torch
    import audio_triangulation_b200 as at
    A, T = 24, 8192 if block >= 256 else 4096
    streams = make_streams(A, T, seed=block)
    st = at.Stream(loc, A)
    got = [[] for _ in range(A)]
    d_streams = torch.from_numpy(streams).cuda()
    for b0 in range(0, T, block):
        r = st.push(d_streams[:, b0:b0 + block].contiguous())
        torch.cuda.synchronize()
        fired = r["fired"].cpu().numpy()
        for a in np.nonzero(fired > 0)[0]:
            got[a].append((b0 + int(fired[a]), int(r["heads"][a].item()), r["frames"][a].cpu().numpy()))
    n_onsets = 0
    for a in range(A):
        exp = oracle_onsets(oracle, streams[a])
        assert [g[0] for g in got[a]] == [e[0] for e in exp], (a, [g[0] for g in got[a]], [e[0] for e in exp])
        for g, e in zip(got[a], exp):
            assert g[1] == e[1] and (g[2] == e[2]).all()
        n_onsets += len(exp)
    assert n_onsets >= A          # the streams do contain events


def test_stream_to_localization_pipeline(loc, oracle):
    """push -> captured rings -> at_localize_device (ring order + heads) == oracle on the same rings."""
    import torch
    import audio_triangulation_b200 as at
    A, T, block = 32, 6144, 512
    streams = make_streams(A, T, seed=7)
    st = at.Stream(loc, A)
    d_streams = torch.from_numpy(streams).cuda()
    checked = 0
    for b0 in range(0, T, block):
        r = st.push(d_streams[:, b0:b0 + block].contiguous())
        idx = torch.nonzero(r["fired"] > 0).flatten()
        if idx.numel() == 0:
            continue
        frames, heads = r["frames"][idx].contiguous(), r["heads"][idx].contiguous()
        res = loc.localize_device(frames, heads, want=("lags", "cell", "gate"))
        torch.cuda.synchronize()
        o = oracle.localize(frames.cpu().numpy(), heads=heads.cpu().numpy(), want_corr=False)
        assert (res["lags"].cpu().numpy() == o["lags"]).all() and (res["cell"].cpu().numpy() == o["cell"]).all()
        checked += idx.numel()
    assert checked >= A // 2


def test_stream_argument_checks(loc):
    import torch
    import audio_triangulation_b200 as at
    st = at.Stream(loc, 4)
    with pytest.raises(at.AtError):
        st.push(torch.zeros((4, 24, 3), dtype=torch.uint8, device="cuda"))      # not a multiple of 16
    with pytest.raises(at.AtError):
        st.push(torch.zeros((4, 2048, 3), dtype=torch.uint8, device="cuda"))    # more than one frame length
    r = st.push(torch.full((4, 1024, 3), 128, dtype=torch.uint8, device="cuda"))
    torch.cuda.synchronize()
    assert (r["fired"].cpu().numpy() == -1).all()                                # silence never fires
