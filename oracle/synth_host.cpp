/* oracle/synth_host.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The bench workload's synthetic frames on the host, WITHOUT the product library: bench.py's `--impl reference` arm
 * builds its inputs here so that the only shared objects it maps are the reference's own (oracle/_ref) and this
 * checker.  The per-byte arithmetic is the generator header the product also compiles
 * (audio_triangulation_b200/csrc/at_synth.h: integer-only, counter-based, so every build emits the same bytes); the
 * per-cell propagation-delay table is restated from at_create (at_api.cu, "synthetic-source propagation delays") for
 * the reference geometry, with the microphone coordinates of microphones.c:9-61 as restated in at_oracle.c.
 * tests/test_gpu_parity.py::test_oracle_synth_equals_product_synth pins the two generators to each other.
 */
#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "../audio_triangulation_b200/csrc/at_synth.h"

extern "C" void ato_mics_triangle(float d_ab, float d_bc, float d_ca, int mirror, int rotate, float *xy);

static bool g_plain = false;
namespace {
struct Job {
    uint64_t seed; uint32_t flags; size_t first, lo, hi; int n_cells;
    const int32_t *delay_q8; uint8_t *adc; int32_t *heads; int32_t *cell;
};
// at_synth_sample with the source's white sequence r(k) (shared by the three channels of a frame) tabulated once per
// frame instead of re-hashed ten times per byte; same integer arithmetic, same bytes (pinned by the tests).
constexpr int KLO = -256, KHI = 4096 + 256;
inline uint8_t sample_cached(uint64_t seed, uint64_t f, const at_synth_frame &p, int mic, int i, int32_t delay_q8,
                             const int32_t *r /* r[k - KLO] */, int32_t dc, uint64_t noise_prefix)
{
    const int64_t pos = (int64_t)i * 256 - delay_q8;
    const int64_t k = pos >> 6;
    if (k - 7 < KLO || k + 1 >= KHI) return at_synth_sample(seed, f, p, mic, i, delay_q8);
    const int32_t frac = (int32_t)(pos & 63);
    int32_t s0 = 0;
    for (int j = 0; j < 8; j++) s0 += r[k - j - KLO];
    const int32_t s1 = s0 - r[k - 7 - KLO] + r[k + 1 - KLO];
    const int32_t val = s0 * (64 - frac) + s1 * frac;
    const int32_t n = (int32_t)(pos >> 8);
    int32_t tri = 400 - (n > 600 ? n - 600 : 600 - n);
    if (tri < 0) tri = 0;
    const int32_t env = (tri * tri) >> 9;
    const int32_t sig = (val * env) >> 17;
    const uint64_t hn = at_hash_final(noise_prefix, (uint64_t)i);
    const int32_t nz = (int32_t)((hn & 15) + ((hn >> 4) & 15) + ((hn >> 8) & 15) + ((hn >> 12) & 15)) - 30;
    int32_t v = 128 + dc + sig + ((nz * p.noise_mul) >> 3);
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    return (uint8_t)v;
}
void *worker(void *arg)
{
    const Job &j = *(const Job *)arg;
    const int M = 3, N = 1024;
    std::vector<int32_t> r(KHI - KLO);
    for (size_t fl = j.lo; fl < j.hi; fl++) {
        const uint64_t f = j.first + fl;
        const at_synth_frame fp = at_synth_frame_params(j.seed, j.flags, f, j.n_cells, 10);
        if (fp.kat < 0 && !g_plain) {
            const uint64_t src = at_hash_prefix(j.seed, f, 1);
            for (int k = KLO; k < KHI; k++) r[k - KLO] = at_synth_r(src, k);
        }
        for (int m = 0; m < M; m++) {
            int32_t dq = j.delay_q8[(size_t)fp.cell * M + m];
            if (j.flags & AT_SYNTH_F_INTEGER_DELAYS) dq = (dq + 128) & ~255;
            uint8_t *dst = j.adc + (fl * M + m) * (size_t)N;
            if (fp.kat >= 0 || g_plain) {
                for (int i = 0; i < N; i++) dst[(fp.head + i) & (N - 1)] = at_synth_sample(j.seed, f, fp, m, i, dq);
            } else {
                const int32_t dc = (int32_t)(at_hash3(j.seed, f, 0xDCull, (uint64_t)m) % 17u) - 8;
                const uint64_t np = at_hash_prefix(j.seed, f, 2 + (uint64_t)m);
                for (int i = 0; i < N; i++)
                    dst[(fp.head + i) & (N - 1)] = sample_cached(j.seed, f, fp, m, i, dq, r.data(), dc, np);
            }
        }
        if (j.heads) j.heads[fl] = fp.head;
        if (j.cell) j.cell[fl] = fp.cell;
    }
    return nullptr;
}
}  // namespace

/* (ato_synth_frames_plain: every byte through at_synth_sample itself, for the self-check in tests/test_oracle_golden.py) */
extern "C" void ato_synth_plain(int on) { g_plain = on != 0; }

/* Frames [first, first + n) of the reference-geometry workload (3 mics x 1024 samples, 101 x 101 cells at 24 px/m on
 * the 1.2 m sphere, 50 kHz, 343 m/s), ring order with the generator's heads.  adc [n][3][1024]; heads / cell may be NULL. */
extern "C" void ato_synth_frames(uint64_t seed, uint32_t flags, size_t first, size_t n, uint8_t *adc, int32_t *heads,
                                 int32_t *cell, int nthreads)
{
    const int M = 3, half_w = 50, half_h = 50, W = 2 * half_w + 1, n_cells = W * (2 * half_h + 1);
    const double px_per_m = 24.0f, height = 1.2f, speed = 343.0f, rate = 50000.f;
    float mic[6];
    ato_mics_triangle(0.132f, 0.15f, 0.20f, 1, 0, mic);
    std::vector<int32_t> dq((size_t)n_cells * M);
    for (int c = 0; c < n_cells; c++) {
        double x = (c % W - half_w) / px_per_m, y = (half_h - c / W) / px_per_m, z = height;
        const double k = height / sqrt(x * x + y * y + z * z);
        x *= k; y *= k; z *= k;
        for (int m = 0; m < M; m++) {
            const double dx = x - mic[2 * m], dy = y - mic[2 * m + 1];
            dq[(size_t)c * M + m] = (int32_t)llround(256.0 * (sqrt(dx * dx + dy * dy + z * z) - height) / speed * rate);
        }
    }
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    std::vector<pthread_t> th(nthreads);
    std::vector<Job> jobs(nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = Job{seed, flags, first, n * t / nthreads, n * (t + 1) / nthreads, n_cells, dq.data(), adc, heads, cell};
        pthread_create(&th[t], nullptr, worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], nullptr);
}
