/* oracle/at_oracle.c -- TEST INFRASTRUCTURE ONLY.  See at_oracle.h.
 *
 * Every function cites the reference lines (relative to /root/reference/src) it restates.
 * Integer steps are exact; the float steps use the same operand types, operation order and
 * conversions as the reference compiled with -O2 -fno-fast-math -ffp-contract=off on x86-64.
 */
#include "at_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* components/rolling_buffer.c:43-71 -- un-rotate the ring from `head`, subtract
 * floor(mean) truncated to int16, return sum of squares of the result. */
int64_t ato_dc_remove(const int16_t *ring, int n_bits, int head, int16_t *out)
{
    const int n = 1 << n_bits;
    int64_t total = 0;
    for (int i = 0; i < n; i++) {
        int16_t v = ring[(head + i) & (n - 1)];
        out[i] = v;
        total += v;
    }
    const int16_t mean = (int16_t)(total >> n_bits);   /* :64, arithmetic shift = floor */
    int64_t power = 0;
    for (int i = 0; i < n; i++) {
        out[i] = (int16_t)(out[i] - mean);              /* :66 */
        power += (int64_t)out[i] * out[i];              /* :70 */
    }
    return power;
}

/* components/buffer.c:15-16 (the live part of buffer_normalize_range; :20-48 is dead code
 * behind the early return at :18).  int16 <<= 8 wraps. */
void ato_shift8(int16_t *x, int n)
{
    for (int i = 0; i < n; i++)
        x[i] = (int16_t)(uint16_t)((uint32_t)(int32_t)x[i] << 8);
}

/* components/buffer.c:4-11 -- Q15 multiply, arithmetic shift, truncate; shorter frames
 * decimate the table (:8). */
void ato_window(int16_t *x, int n_bits, const int16_t *table, int table_bits)
{
    const int n = 1 << n_bits, step = table_bits - n_bits;
    for (int i = 0; i < n; i++) {
        int32_t p = (int32_t)x[i] * (int32_t)table[i << step];
        x[i] = (int16_t)(p >> 15);
    }
}

/* components/correlations.c:7-24 -- lagged dot products over [-L, L], int64 accumulate,
 * strict '>' arg-max scanning lags upward (ties resolve to the most negative lag). */
void ato_xcorr(const int16_t *a, const int16_t *b, int n, int L, int64_t *c, int32_t *best)
{
    int64_t top = INT64_MIN;
    int32_t arg = 0;
    for (int s = -L; s <= L; s++) {
        const int lo = s < 0 ? -s : 0;          /* first index into a */
        const int hi = s < 0 ? n : n - s;       /* one past last index into a */
        int64_t acc = 0;
        for (int j = lo; j < hi; j++)
            acc += (int32_t)a[j] * (int32_t)b[j + s];
        c[s + L] = acc;
        if (acc > top) { top = acc; arg = s; }
    }
    *best = arg;
}

/* components/correlations.c:26-33 -- c[s] = (int64)((float)c[s] * (float)exp(-(s-best)^2 / 36.f)) */
void ato_gauss(int64_t *c, int L, int best)
{
    for (int s = -L; s <= L; s++) {
        int d = s - best;
        d *= d;
        const float scale = (float)exp((double)((float)(-d) / 36.f));
        c[s + L] = (int64_t)((float)c[s + L] * scale);
    }
}

/* components/correlations.c:38-63 -- exponential moving average with time constant 0.5 s,
 * float arithmetic on int64 data, re-arg-max, stamp. */
void ato_average(int64_t *est, int32_t *est_best, uint64_t *est_time,
                 const int64_t *fresh, int L, uint64_t now_us)
{
    const float dt = (float)(now_us - *est_time) / 1e6f;
    const float decay = (float)(1.0 - exp((double)(-dt / 0.5f)));
    int64_t top = INT64_MIN;
    for (int i = 0; i < 2 * L + 1; i++) {
        const float step = (float)(fresh[i] - est[i]) * decay;
        est[i] = (int64_t)((float)est[i] + step);
        if (est[i] > top) { top = est[i]; *est_best = i - L; }
    }
    *est_time = now_us;
}

/* components/rolling_buffer.c:3-14 */
void ato_ring_init(ato_ring *r, int16_t *storage, int n_bits)
{
    memset(r, 0, sizeof *r);
    r->n_bits = n_bits;
    r->buf = storage;
    memset(storage, 0, sizeof(int16_t) << n_bits);
}

/* components/rolling_buffer.c:16-41 -- the sample half a ring back leaves the "incoming"
 * (newer) half-window sums and enters the "outgoing" (older) ones; the slot being
 * overwritten leaves the outgoing sums. */
void ato_ring_push(ato_ring *r, int16_t s)
{
    const int n = 1 << r->n_bits, half = n >> 1;
    const int16_t mid = r->buf[(r->head + half) & (n - 1)];
    const int16_t old = r->buf[r->head];
    r->out_tot += (int64_t)mid - old;
    r->out_pow += (int64_t)mid * mid - (int64_t)old * old;
    r->in_tot += (int64_t)s - mid;
    r->in_pow += (int64_t)s * s - (int64_t)mid * mid;
    r->buf[r->head] = s;
    if (++r->head >= n) { r->head = 0; r->full = 1; }
}

/* components/rolling_buffer.c:73-85 -- (N/2)*sum(x^2) - (sum x)^2 */
int64_t ato_ring_incoming(const ato_ring *r)
{
    return (int64_t)((uint64_t)r->in_pow << (r->n_bits - 1)) - r->in_tot * r->in_tot;
}
int64_t ato_ring_outgoing(const ato_ring *r)
{
    return (int64_t)((uint64_t)r->out_pow << (r->n_bits - 1)) - r->out_tot * r->out_tot;
}

/* sample_compute.h:55-99 -- capture loop with the onset gate (:75-91); threshold :21. */
long ato_capture(const uint8_t *stream, size_t n, int n_mics, int n_bits,
                 int16_t *ring_storage, int32_t *head_out)
{
    ato_ring *r = calloc((size_t)n_mics, sizeof *r);
    const int64_t threshold = (int64_t)2 << (2 * (n_bits - 1));
    long fired = -1;
    for (int m = 0; m < n_mics; m++) ato_ring_init(&r[m], ring_storage + ((size_t)m << n_bits), n_bits);
    for (size_t t = 0; t < n && fired < 0; t++) {
        int all_full = 1;
        for (int m = 0; m < n_mics; m++) {
            ato_ring_push(&r[m], (int16_t)stream[t * n_mics + m]);
            all_full &= r[m].full;
        }
        if (all_full) {
            int64_t out = 0, in = 0;
            for (int m = 0; m < n_mics; m++) { out += ato_ring_outgoing(&r[m]); in += ato_ring_incoming(&r[m]); }
            if (out > threshold + in) fired = (long)(t + 1);
        }
    }
    if (head_out) *head_out = r[0].head;
    free(r);
    return fired;
}

/* components/microphones.c:9-61 -- triangle from side lengths, centroid at origin,
 * optional y mirror / rotation of A onto +x.  float32 throughout. */
void ato_mics_triangle(float d_ab, float d_bc, float d_ca, int mirror, int rotate, float *xy)
{
    const float xc = (d_ab * d_ab + d_ca * d_ca - d_bc * d_bc) / (2.0f * d_ab);
    const float yc = sqrtf(fmaxf(0.0f, d_ca * d_ca - xc * xc));
    float px[3] = {0.0f, d_ab, xc};
    float py[3] = {0.0f, 0.0f, yc * (mirror ? -1.0f : 1.0f)};
    const float cx = (px[0] + px[1] + px[2]) / 3.0f;
    const float cy = (py[0] + py[1] + py[2]) / 3.0f;
    for (int m = 0; m < 3; m++) { xy[2 * m] = px[m] - cx; xy[2 * m + 1] = py[m] - cy; }
    if (rotate) {
        const float th = atan2f(xy[1], xy[0]);
        const float c = cosf(-th), s = sinf(-th);
        for (int m = 0; m < 3; m++) {
            const float x = xy[2 * m], y = xy[2 * m + 1];
            xy[2 * m] = x * c - y * s;
            xy[2 * m + 1] = x * s + y * c;
        }
    }
}

/* components/vga/vga_heatmap.h:11-13 */
static float norm3(float x, float y, float z) { return sqrtf(x * x + y * y + z * z); }

/* components/vga/vga_heatmap.h:50-92 -- per cell: plane point -> sphere of radius `height`
 * -> mic distances -> expected lag per pair, roundf, clamp to +-L, store lag+L.
 * Pairs are enumerated (0,1),(0,2)..(0,M-1),(1,2).. which is ab, ac, bc for M=3. */
void ato_lut_build(const float *mic_xy, int n_mics, int L, float rate_hz, float speed,
                   int half_w, int half_h, float px_per_m, float height, uint8_t *idx)
{
    const int W = 2 * half_w + 1, H = 2 * half_h + 1;
    float *dist = malloc(sizeof(float) * (size_t)n_mics);
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            float xm = (float)(x - half_w) / px_per_m;
            float ym = (float)(half_h - y) / px_per_m;
            float zm = height;
            const float k = height / norm3(zm, xm, ym);
            xm *= k; ym *= k; zm *= k;
            for (int m = 0; m < n_mics; m++)
                dist[m] = norm3(zm, xm - mic_xy[2 * m], ym - mic_xy[2 * m + 1]);
            int p = 0;
            for (int i = 0; i < n_mics; i++)
                for (int j = i + 1; j < n_mics; j++, p++) {
                    const float dt = (dist[j] - dist[i]) / speed;
                    int s = (int)roundf(dt * rate_hz);
                    if (s < -L) s = -L; else if (s > L) s = L;
                    idx[((size_t)p * H + y) * W + x] = (uint8_t)(s + L);
                }
        }
    }
    free(dist);
}

/* vga_heatmap.h:63-90 with the candidate position given (3-D candidate sets: hemisphere of directions, volume grid) */
void ato_lut_build_points(const float *mic_xy, int n_mics, int L, float rate_hz, float speed,
                          const float *points, int n_points, uint8_t *idx)
{
    float *dist = malloc(sizeof(float) * (size_t)n_mics);
    for (int c = 0; c < n_points; c++) {
        const float xm = points[3 * c], ym = points[3 * c + 1], zm = points[3 * c + 2];
        for (int m = 0; m < n_mics; m++)
            dist[m] = norm3(zm, xm - mic_xy[2 * m], ym - mic_xy[2 * m + 1]);
        int p = 0;
        for (int i = 0; i < n_mics; i++)
            for (int j = i + 1; j < n_mics; j++, p++) {
                const float dt = (dist[j] - dist[i]) / speed;
                int s = (int)roundf(dt * rate_hz);
                if (s < -L) s = -L; else if (s > L) s = L;
                idx[(size_t)p * n_points + c] = (uint8_t)(s + L);
            }
    }
    free(dist);
}

/* components/vga/vga_heatmap.h:96-126 -- L(cell) = sum over pairs of corr[pair][lut[pair][cell]];
 * its maximum (strict '>' in row-major order; we also return the first cell reaching it);
 * thresholds 63/64, 31/32, 15/16, 7/8 via multiply + arithmetic shift (:111-114);
 * class per cell with the reference's colour codes WHITE=15 GREEN=3 RED=8 BLUE=5 BLACK=0
 * (lib/vga/vga16_graphics.h:31-34). */
void ato_heatmap(const int64_t *corr, const uint8_t *idx, int n_pairs, int n_cells, int L,
                 int64_t *highest, int32_t *first_cell, uint8_t *classes)
{
    const int nl = 2 * L + 1;
    int64_t top = INT64_MIN;
    int32_t arg = -1;
    for (int cidx = 0; cidx < n_cells; cidx++) {
        int64_t like = 0;
        for (int p = 0; p < n_pairs; p++) like += corr[(size_t)p * nl + idx[(size_t)p * n_cells + cidx]];
        if (like > top) { top = like; arg = cidx; }
    }
    if (highest) *highest = top;
    if (first_cell) *first_cell = arg;
    if (!classes) return;
    const int64_t t_white = (top * 63) >> 6, t_green = (top * 31) >> 5;
    const int64_t t_red = (top * 15) >> 4, t_blue = (top * 7) >> 3;
    for (int cidx = 0; cidx < n_cells; cidx++) {
        int64_t like = 0;
        for (int p = 0; p < n_pairs; p++) like += corr[(size_t)p * nl + idx[(size_t)p * n_cells + cidx]];
        classes[cidx] = like >= t_white ? 15 : like >= t_green ? 3 : like >= t_red ? 8 : like >= t_blue ? 5 : 0;
    }
}

/* ---------------- batch driver: sample_compute.h:104-122 for every frame ---------------- */
struct span {
    const ato_config *cfg; const uint8_t *adc; const int32_t *heads; size_t lo, hi;
    int32_t *lags; int64_t *corr; int64_t *raw; int32_t *cell; int64_t *highest;
};

static void *span_run(void *arg)
{
    struct span *sp = arg;
    const ato_config *g = sp->cfg;
    const int M = g->n_mics, n = 1 << g->n_bits, L = g->max_shift, nl = 2 * L + 1;
    const int P = M * (M - 1) / 2;
    int16_t *ring = malloc(sizeof(int16_t) * (size_t)n);
    int16_t *sig = malloc(sizeof(int16_t) * (size_t)n * M);
    int64_t *cur = malloc(sizeof(int64_t) * (size_t)nl * P);
    for (size_t f = sp->lo; f < sp->hi; f++) {
        const uint8_t *frame = sp->adc + f * (size_t)M * n;
        const int head = sp->heads ? sp->heads[f] : 0;
        for (int m = 0; m < M; m++) {
            for (int i = 0; i < n; i++) ring[i] = (int16_t)frame[(size_t)m * n + i];
            ato_dc_remove(ring, g->n_bits, head, sig + (size_t)m * n);
            ato_shift8(sig + (size_t)m * n, n);
            ato_window(sig + (size_t)m * n, g->n_bits, g->window, g->window_bits);
        }
        int p = 0;
        for (int i = 0; i < M; i++)
            for (int j = i + 1; j < M; j++, p++) {
                int32_t best;
                ato_xcorr(sig + (size_t)i * n, sig + (size_t)j * n, n, L, cur + (size_t)p * nl, &best);
                if (sp->raw) memcpy(sp->raw + (f * P + p) * nl, cur + (size_t)p * nl, sizeof(int64_t) * nl);
                if (sp->lags) sp->lags[f * P + p] = best;
                if (sp->corr || sp->cell || sp->highest) ato_gauss(cur + (size_t)p * nl, L, best);
                if (sp->corr) memcpy(sp->corr + (f * P + p) * nl, cur + (size_t)p * nl, sizeof(int64_t) * nl);
            }
        if ((sp->cell || sp->highest) && g->lut)
            ato_heatmap(cur, g->lut, P, g->n_cells, L, sp->highest ? sp->highest + f : NULL,
                        sp->cell ? sp->cell + f : NULL, NULL);
    }
    free(ring); free(sig); free(cur);
    return NULL;
}

void ato_localize(const ato_config *cfg, const uint8_t *adc, const int32_t *heads, size_t n_frames,
                  int32_t *lags, int64_t *corr, int64_t *raw, int32_t *cell, int64_t *highest,
                  int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 512) nthreads = 512;
    pthread_t th[512];
    struct span sp[512];
    for (int t = 0; t < nthreads; t++) {
        sp[t] = (struct span){cfg, adc, heads, n_frames * t / nthreads, n_frames * (t + 1) / nthreads,
                              lags, corr, raw, cell, highest};
        if (nthreads == 1) span_run(&sp[t]);
        else pthread_create(&th[t], NULL, span_run, &sp[t]);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}
