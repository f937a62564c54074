"""Deterministic numpy test frames (CPU side; the product's own generator is at_synth_*)."""
import numpy as np

N = 1024


def kat_frames():
    """Known-answer frames of SURVEY section 4.  Returns (names, adc uint8 [K,3,1024])."""
    names, out = [], []

    def add(name, a, b, c):
        names.append(name)
        out.append(np.stack([np.clip(x, 0, 255).astype(np.uint8) for x in (a, b, c)]))

    base = np.full(N, 128, np.int64)
    add("silence", np.full(N, 128), np.full(N, 131), np.full(N, 126))
    a, b, c = base.copy(), base.copy(), base.copy()
    a[500] += 100; b[511] += 100; c[480] += 100
    add("impulse", a, b, c)
    sq = np.where((np.arange(N) >> 3) & 1, 255, 0)
    add("fullscale_square", sq, np.roll(sq, 3), np.roll(sq, -5))
    rng = np.random.default_rng(1234)
    src = np.convolve(rng.normal(0, 60, 2 * N), np.ones(3) / 3, "same") * np.exp(-0.5 * ((np.arange(2 * N) - 1100) / 200.0) ** 2)

    def delayed(d):
        return 128 + np.round(src[500 - d:500 - d + N])
    add("burst_dB7_dC-12", delayed(0), delayed(7), delayed(-12))
    add("out_of_window_dB30_dC-40", delayed(0), delayed(30), delayed(-40))
    add("all_zero", np.zeros(N), np.zeros(N), np.zeros(N))
    add("all_255", np.full(N, 255), np.full(N, 255), np.full(N, 255))
    ramp = np.arange(N) >> 2
    add("ramp", ramp, 255 - ramp, ramp)
    return names, np.stack(out)


def burst_frames(count, seed, n_mics=3, n_samples=N, max_delay=28):
    rng = np.random.default_rng(seed)
    adc = np.empty((count, n_mics, n_samples), np.uint8)
    delays = rng.integers(-max_delay, max_delay + 1, (count, n_mics))
    delays[:, 0] = 0
    t = np.arange(3 * n_samples)
    for f in range(count):
        kind = f % 8
        if kind == 7:   # plain white noise, full range
            adc[f] = rng.integers(0, 256, (n_mics, n_samples), dtype=np.uint8)
            continue
        src = np.convolve(rng.normal(0, 50 + 10 * kind, t.size), np.ones(2 + kind % 3) / (2 + kind % 3), "same")
        src *= np.exp(-0.5 * ((t - 1.5 * n_samples - 80) / (0.2 * n_samples)) ** 2)
        for m in range(n_mics):
            d = int(delays[f, m])
            seg = src[n_samples - d:2 * n_samples - d]
            dc = rng.integers(-8, 9)
            noise = rng.normal(0, [0, 2, 6][f % 3], n_samples)
            adc[f, m] = np.clip(np.round(128 + dc + seg + noise), 0, 255).astype(np.uint8)
    return adc, delays
