/* oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin driver around the reference's OWN compute-path objects (buffer.c,
 * rolling_buffer.c, correlations.c, microphones.c compiled unmodified from
 * /root/reference/src by oracle/Makefile).  It supplies the one SDK symbol those
 * objects need (get_absolute_time, used at correlations.c:35 and :40), and walks
 * frames through the reference functions in the order sample_compute.h:104-122
 * prescribes, so tests and the CPU-baseline leg of bench.py can call the real
 * reference from Python (ctypes).
 */
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <components/buffer.h>
#include <components/correlations.h>
#include <components/microphones.h>
#include <components/rolling_buffer.h>

/* ---- injected clock (per thread so the threaded timing leg stays race-free) ---- */
static __thread uint64_t t_now_us = 0;
absolute_time_t get_absolute_time(void) { return t_now_us; }
void ref_set_time(uint64_t now_us) { t_now_us = now_us; }

/* ---- layout facts, for the drop-in struct-compat test ---- */
void ref_layout(int64_t out[12])
{
    out[0] = sizeof(struct buffer_t);
    out[1] = offsetof(struct buffer_t, power);
    out[2] = sizeof(struct rolling_buffer_t);
    out[3] = offsetof(struct rolling_buffer_t, incoming_power);
    out[4] = offsetof(struct rolling_buffer_t, is_full);
    out[5] = offsetof(struct rolling_buffer_t, buffer);
    out[6] = sizeof(struct correlations_t);
    out[7] = offsetof(struct correlations_t, best_shift);
    out[8] = offsetof(struct correlations_t, last_update);
    out[9] = MAX_SHIFT_SAMPLES;
    out[10] = CORRELATION_BUFFER_SIZE;
    out[11] = BUFFER_SIZE;
}

void ref_mics(float out[6])
{
    microphones_init();
    out[0] = mic_a_location.x; out[1] = mic_a_location.y;
    out[2] = mic_b_location.x; out[3] = mic_b_location.y;
    out[4] = mic_c_location.x; out[5] = mic_c_location.y;
}

/* A ring in the state the capture loop leaves it in: `head` arbitrary, samples stored
 * in ring order (chronological sample i lives at (head+i) mod N). */
static void ring_from_frame(struct rolling_buffer_t *rb, const uint8_t *chrono, int head)
{
    rolling_buffer_init(rb);
    rb->head = head;
    rb->is_full = true;
    for (int i = 0; i < BUFFER_SIZE; i++)
        rb->buffer[(head + i) & (BUFFER_SIZE - 1)] = (sample_t)chrono[i];
}

/* One frame, every stage exposed (sample_compute.h:105-122).  adc = [3][1024] uint8,
 * chronological.  Any output pointer may be NULL. */
void ref_frame_stages(const uint8_t *adc, int head,
                      int16_t *after_dc /*[3][1024]*/, int64_t *power /*[3]*/,
                      int16_t *after_shift /*[3][1024]*/, int16_t *after_window /*[3][1024]*/,
                      struct correlations_t *corr /*[3] ab, ac, bc*/)
{
    struct rolling_buffer_t rb;
    struct buffer_t buf[3];
    for (int m = 0; m < 3; m++) {
        ring_from_frame(&rb, adc + (size_t)m * BUFFER_SIZE, head);
        rolling_buffer_write_out(&rb, &buf[m]);
        if (after_dc) memcpy(after_dc + (size_t)m * BUFFER_SIZE, buf[m].buffer, sizeof buf[m].buffer);
        if (power) power[m] = buf[m].power;
        buffer_normalize_range(&buf[m]);
        if (after_shift) memcpy(after_shift + (size_t)m * BUFFER_SIZE, buf[m].buffer, sizeof buf[m].buffer);
        buffer_window(&buf[m]);
        if (after_window) memcpy(after_window + (size_t)m * BUFFER_SIZE, buf[m].buffer, sizeof buf[m].buffer);
    }
    if (corr) {
        correlations_init(&corr[0], &buf[0], &buf[1]);
        correlations_init(&corr[1], &buf[0], &buf[2]);
        correlations_init(&corr[2], &buf[1], &buf[2]);
    }
}

struct job {
    const uint8_t *adc; size_t lo, hi; int32_t *lags; struct correlations_t *corr; uint64_t now;
};

static void *worker(void *p)
{
    struct job *j = p;
    struct correlations_t c[3];
    t_now_us = j->now;
    for (size_t f = j->lo; f < j->hi; f++) {
        struct correlations_t *dst = j->corr ? j->corr + 3 * f : c;
        ref_frame_stages(j->adc + f * 3 * BUFFER_SIZE, 0, NULL, NULL, NULL, NULL, dst);
        if (j->lags) {
            j->lags[3 * f + 0] = dst[0].best_shift;
            j->lags[3 * f + 1] = dst[1].best_shift;
            j->lags[3 * f + 2] = dst[2].best_shift;
        }
    }
    return NULL;
}

/* Batch of chronological frames adc[n][3][1024] through steps a9-a15 of SURVEY 8a,
 * disjoint frame ranges on `nthreads` pthreads.  corr may be NULL. */
void ref_localize_frames(const uint8_t *adc, size_t n, int32_t *lags,
                         struct correlations_t *corr, int nthreads, uint64_t now_us)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 512) nthreads = 512;
    pthread_t th[512];
    struct job jobs[512];
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (struct job){adc, n * t / nthreads, n * (t + 1) / nthreads, lags, corr, now_us};
        if (nthreads == 1) worker(&jobs[t]);
        else pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

/* The capture loop of sample_compute.h:55-99 driven by a recorded triple stream
 * stream[n][3] (A,B,C bytes per 20 us tick).  Returns the number of ticks consumed when
 * the onset gate fires (:89-90), or -1 if the stream ends first.  rings[3] receive the
 * final ring state. */
long ref_capture(const uint8_t *stream, size_t n, struct rolling_buffer_t *rings /*[3]*/)
{
    const power_t threshold = ((power_t)2) << (2 * BUFFER_HALF_SIZE_BITS); /* sample_compute.h:21 */
    for (int m = 0; m < 3; m++) rolling_buffer_init(&rings[m]);
    for (size_t t = 0; t < n; t++) {
        for (int m = 0; m < 3; m++) rolling_buffer_push(&rings[m], (sample_t)stream[3 * t + m]);
        if (rings[0].is_full && rings[1].is_full && rings[2].is_full) {
            power_t out = 0, in = 0;
            for (int m = 0; m < 3; m++) {
                out += rolling_buffer_get_outgoing_power(&rings[m]);
                in += rolling_buffer_get_incoming_power(&rings[m]);
            }
            if (out > threshold + in) return (long)(t + 1);
        }
    }
    return -1;
}
