// at_synth.h -- integer-only, counter-based synthetic frame generator shared by the host
// harness path (at_synth_host) and the CUDA generator kernel (at_synth_device).  Because every
// step is integer arithmetic on a splitmix64 counter hash, both emit identical bytes.
//
// It stands in for the Pico's ADC/DMA capture (ref: components/dma_sampler.c:3-56): what the
// compute path sees is uint8 triples in channel order A,B,C (ref: sample_compute.h:67-69).
//
// Model per frame f: a source at a heat-map cell -> per-mic propagation delay (Q8
// samples, table built on the host in double at at_create); source signal = band-limited noise
// (8-tap box over a 4x oversampled +-128 hash sequence, linear interpolation at the fractional
// delay) under a squared-triangle envelope centred on sample 600; per-channel DC offset in
// [-8, 8]; additive noise with scale {0, 2, 6}/8 of a +-30 four-nibble sum; rounded, clipped.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AT_HD __host__ __device__ __forceinline__
#else
#define AT_HD static inline
#endif

#define AT_SYNTH_F_INTEGER_DELAYS 1u
#define AT_SYNTH_F_RANDOM_HEADS 2u
#define AT_SYNTH_F_KATS 4u
#define AT_SYNTH_F_MAX_NOISE 8u    /* every frame at the lowest SNR of the model (noise scale 6) */
#define AT_SYNTH_F_WHITE 16u       /* no source at all: independent uniform bytes (worst case for every shortcut) */
#define AT_SYNTH_KAT_WHITE 100
#define AT_SYNTH_N_KATS 4

AT_HD uint64_t at_mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// hash(seed, a, b, c) split so the (seed, a, b) prefix is computed once per frame/channel
AT_HD uint64_t at_hash_prefix(uint64_t seed, uint64_t a, uint64_t b)
{
    const uint64_t h = at_mix64(seed ^ (a * 0xD6E8FEB86659FD93ull));
    return at_mix64(h ^ (b * 0xA0761D6478BD642Full));
}
AT_HD uint64_t at_hash_final(uint64_t prefix, uint64_t c)
{
    return at_mix64(prefix ^ (c * 0xE7037ED1A0B428DBull));
}
AT_HD uint64_t at_hash3(uint64_t seed, uint64_t a, uint64_t b, uint64_t c)
{
    return at_hash_final(at_hash_prefix(seed, a, b), c);
}

struct at_synth_frame {
    int32_t cell;      // source cell (row-major y*W + x)
    int32_t head;      // ring head the frame is stored with
    int32_t noise_mul; // 0, 2 or 6
    int32_t kat;       // -1 or known-answer frame id
};

AT_HD at_synth_frame at_synth_frame_params(uint64_t seed, uint32_t flags, uint64_t f, int n_cells, int n_bits)
{
    at_synth_frame p;
    const uint64_t h = at_hash3(seed, f, 0xF00Dull, 0);
    p.cell = (int32_t)((h & 0xFFFFFFull) % (uint64_t)n_cells);
    p.head = (flags & AT_SYNTH_F_RANDOM_HEADS) ? (int32_t)((h >> 24) & ((1u << n_bits) - 1)) : 0;
    const uint32_t sel = (uint32_t)((h >> 40) % 3u);
    p.noise_mul = (flags & AT_SYNTH_F_MAX_NOISE) ? 6 : (sel == 0 ? 0 : (sel == 1 ? 2 : 6));
    p.kat = ((flags & AT_SYNTH_F_KATS) && f < AT_SYNTH_N_KATS) ? (int32_t)f : ((flags & AT_SYNTH_F_WHITE) ? AT_SYNTH_KAT_WHITE : -1);
    return p;
}

// white +-128 hash sequence at oversampled (4x) index k; the source is its 8-tap box sum
AT_HD int32_t at_synth_r(uint64_t src_prefix, int64_t k)
{
    return (int32_t)(at_hash_final(src_prefix, (uint64_t)k) & 0xFF) - 128;
}

// Known-answer frames: 0 silence, 1 impulse pair, 2 full-scale square, 3 ramp+impulse.
AT_HD int32_t at_synth_kat(int kat, int mic, int i)
{
    switch (kat) {
    case 0: return mic == 0 ? 128 : (mic == 1 ? 131 : 126);
    case 1: { const int at = mic == 0 ? 500 : (mic == 1 ? 511 : 480); return i == at ? 228 : 128; }
    case 2: return ((i + 3 * mic) >> 3) & 1 ? 255 : 0;
    default: { const int at = 300 + 37 * mic; return i == at ? 255 : (i >> 2) & 0xFF; }
    }
}

// One output byte: frame f, mic m, chronological sample i.  delay_q8 = that mic's delay.
AT_HD uint8_t at_synth_sample(uint64_t seed, uint64_t f, const at_synth_frame &p, int mic, int i,
                              int32_t delay_q8)
{
    if (p.kat == AT_SYNTH_KAT_WHITE) return (uint8_t)(at_hash_final(at_hash_prefix(seed, f, 2 + (uint64_t)mic), (uint64_t)i) & 0xFF);
    if (p.kat >= 0) return (uint8_t)at_synth_kat(p.kat, mic, i);
    const uint64_t hm = at_hash3(seed, f, 0xDCull, (uint64_t)mic);
    const int32_t dc = (int32_t)(hm % 17u) - 8;
    const int64_t pos = (int64_t)i * 256 - delay_q8;           // Q8 source time
    const int64_t k = pos >> 6;                                // 4x oversampled index (floor)
    const int32_t frac = (int32_t)(pos & 63);
    int32_t s0 = 0;                                            // box sum over k-7..k   (+-1024)
    const uint64_t src = at_hash_prefix(seed, f, 1);
    for (int j = 0; j < 8; j++) s0 += at_synth_r(src, k - j);
    const int32_t s1 = s0 - at_synth_r(src, k - 7) + at_synth_r(src, k + 1); // k-6..k+1
    const int32_t val = s0 * (64 - frac) + s1 * frac;          // +-65536
    const int32_t n = (int32_t)(pos >> 8);
    int32_t tri = 400 - (n > 600 ? n - 600 : 600 - n);
    if (tri < 0) tri = 0;
    const int32_t env = (tri * tri) >> 9;                      // 0..312
    const int32_t sig = (val * env) >> 17;
    const uint64_t hn = at_hash_final(at_hash_prefix(seed, f, 2 + (uint64_t)mic), (uint64_t)i);
    const int32_t nz = (int32_t)((hn & 15) + ((hn >> 4) & 15) + ((hn >> 8) & 15) + ((hn >> 12) & 15)) - 30;
    int32_t v = 128 + dc + sig + ((nz * p.noise_mul) >> 3);
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    return (uint8_t)v;
}
