"""GPU: the hand-written FFT / GCC-PHAT variant against a float64 numpy restatement of the same statistic.
No reference counterpart exists (the reference correlates directly in integers), so this is an agreement test of
arg-max lags, not a bit-exact parity test: float32 vs float64 may disagree on near-ties."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def numpy_gccphat(windowed, L):
    F, M, N = windowed.shape
    X = np.fft.rfft(windowed.astype(np.float64), 2 * N, axis=-1)
    out = []
    for i in range(M):
        for j in range(i + 1, M):
            G = np.conj(X[:, i]) * X[:, j]
            mag = np.abs(G)
            G = np.where(mag > 1e-20, G / np.maximum(mag, 1e-300), 0)
            r = np.fft.irfft(G, 2 * N, axis=-1)
            cur = np.concatenate([r[:, -L:], r[:, :L + 1]], axis=1)        # lags -L..L
            out.append(np.argmax(cur, axis=1) - L)                         # first maximum
    return np.stack(out, 1).astype(np.int32)


@pytest.mark.parametrize("inverse", ["dft", "fft"])       # tcgen05 contraction over the wanted lags / inverse FFTs
@pytest.mark.parametrize("shape", [(3, 10, 128), (8, 12, 24), (8, 10, 64), (4, 10, 37), (8, 12, 300)])
def test_gccphat_lags_agree_with_float64(shape, inverse, monkeypatch):
    import torch
    import audio_triangulation_b200 as at
    monkeypatch.setenv("AT_GCC_INVERSE", inverse)
    M, nb, F = shape
    loc = at.Localizer(n_mics=M, n_bits=nb)
    adc, heads, _ = loc.synth_device(F, flags=1 | 2, seed=99)             # integer delays, random ring heads
    res = loc.localize_device(adc, heads, want=("windowed", "lags"))
    got = loc.gccphat_device(adc, heads)
    torch.cuda.synchronize()
    exp = numpy_gccphat(res["windowed"].cpu().numpy(), loc.n_lags // 2)
    got = got.cpu().numpy()
    agree = (got == exp).mean()
    assert agree >= 0.98, agree                                            # float32 vs float64 near-ties
    assert np.abs(got - exp).max() <= 2 or agree >= 0.995
    # PHAT and the direct integer correlation find the same TDOA on clean integer-delay bursts most of the time
    direct = res["lags"].cpu().numpy()
    assert (np.abs(got - direct) <= 1).mean() >= 0.85


def test_gccphat_inverse_forms_agree(monkeypatch):
    """The two inverse forms see the same whitened spectra: their arg-max lags differ only at float near-ties."""
    import torch
    import audio_triangulation_b200 as at
    loc = at.Localizer(n_mics=8, n_bits=12)
    adc, heads, _ = loc.synth_device(515, flags=2, seed=5)                # ragged: not a multiple of the 8-frame column group
    out = {}
    for inverse in ("dft", "fft"):
        monkeypatch.setenv("AT_GCC_INVERSE", inverse)
        out[inverse] = loc.gccphat_device(adc, heads).cpu().numpy()
    torch.cuda.synchronize()
    assert (out["dft"] == out["fft"]).mean() >= 0.995
    loc.close()
