/* at_harness.c -- plain-C host harness for libat_b200.so.
 *
 * Replaces the Pico's ADC/DMA capture (ref: components/dma_sampler.c) and the VGA debug output
 * (ref: vga_debug.h) with synthetic multi-channel frames in and numbers out, and walks the
 * reference's own call sequence (ref: sample_compute.h:104-139) through the drop-in symbols before
 * running the batched path.
 *
 *   ./at_harness [n_frames] [device]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/at_b200.h"

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

#define CHECK(call) do { if ((call) != AT_OK) { fprintf(stderr, "%s: %s\n", #call, at_last_error()); return 1; } } while (0)

int main(int argc, char **argv)
{
    const size_t F = argc > 1 ? strtoull(argv[1], NULL, 10) : 65536;
    at_config cfg;
    at_config_reference(&cfg);
    if (argc > 2) cfg.device = atoi(argv[2]);
    at_context *ctx;
    CHECK(at_create(&cfg, &ctx));

    /* --- one frame through the reference's call sequence, function by function (drop-in symbols) --- */
    static uint8_t one[3][BUFFER_SIZE];
    int32_t head, src_cell;
    CHECK(at_synth_host(ctx, 0xA7D10, AT_SYNTH_INTEGER_DELAYS, 1000, 1, &one[0][0], &head, &src_cell));
    static struct rolling_buffer_t rb[3];
    static struct buffer_t buf[3];
    static struct correlations_t fresh[3], est[3];
    microphones_init();                                               /* ref: main.c:57 */
    for (int m = 0; m < 3; m++) {
        rolling_buffer_init(&rb[m]);                                  /* ref: sample_compute.h:55-57 */
        for (int i = 0; i < BUFFER_SIZE; i++) rolling_buffer_push(&rb[m], one[m][i]);   /* :71-73 */
        rolling_buffer_write_out(&rb[m], &buf[m]);                    /* :105-107 */
        buffer_normalize_range(&buf[m]);                              /* :110-112 */
        buffer_window(&buf[m]);                                       /* :115-117 */
    }
    correlations_init(&fresh[0], &buf[0], &buf[1]);                   /* :120 */
    correlations_init(&fresh[1], &buf[0], &buf[2]);                   /* :121 */
    correlations_init(&fresh[2], &buf[1], &buf[2]);                   /* :122 */
    int tot = 0;
    for (int p = 0; p < 3; p++) tot += fresh[p].best_shift * fresh[p].best_shift;
    if (tot > 4)                                                      /* :134 */
        for (int p = 0; p < 3; p++) correlations_average(&est[p], &fresh[p]);   /* :137-139 */
    printf("drop-in path : mics A(%.4f,%.4f) B(%.4f,%.4f) C(%.4f,%.4f); lags ab/ac/bc = %d %d %d (source cell %d)\n",
           mic_a_location.x, mic_a_location.y, mic_b_location.x, mic_b_location.y, mic_c_location.x, mic_c_location.y,
           fresh[0].best_shift, fresh[1].best_shift, fresh[2].best_shift, src_cell);

    /* --- the same frame through the batched path must agree --- */
    int32_t lags1[3], cell1; float xy1[2];
    at_outputs o1; memset(&o1, 0, sizeof o1);
    o1.lags = lags1; o1.cell = &cell1; o1.xy = xy1;
    CHECK(at_localize_host(ctx, &one[0][0], NULL, 1, &o1));
    printf("batched path : lags = %d %d %d, cell %d -> (%.3f, %.3f) m   %s\n", lags1[0], lags1[1], lags1[2], cell1, xy1[0], xy1[1],
           (lags1[0] == fresh[0].best_shift && lags1[1] == fresh[1].best_shift && lags1[2] == fresh[2].best_shift) ? "[agree]" : "[MISMATCH]");

    /* --- a batch of synthetic frames, host buffers in, results out --- */
    uint8_t *adc = NULL;                                      /* capture and result buffers: page-locked, copied at the link rate */
    int32_t *lags = NULL, *cell = NULL, *truth = malloc(F * sizeof *truth);
    CHECK(at_host_alloc(ctx, F * 3 * BUFFER_SIZE, (void **)&adc));
    CHECK(at_host_alloc(ctx, F * 3 * sizeof *lags, (void **)&lags));
    CHECK(at_host_alloc(ctx, F * sizeof *cell, (void **)&cell));
    if (!truth) return 2;
    double t0 = now_s();
    CHECK(at_synth_host(ctx, 0xA7D10, 0, 0, F, adc, NULL, truth));
    double t1 = now_s();
    at_outputs o; memset(&o, 0, sizeof o);
    o.lags = lags; o.cell = cell;
    CHECK(at_localize_host(ctx, adc, NULL, F, &o));           /* warm-up (allocations) */
    double t2 = now_s();
    CHECK(at_localize_host(ctx, adc, NULL, F, &o));
    double t3 = now_s();
    size_t near = 0;
    for (size_t f = 0; f < F; f++) {
        const int dx = cell[f] % 101 - truth[f] % 101, dy = cell[f] / 101 - truth[f] / 101;
        near += (dx * dx + dy * dy) <= 25;
    }
    printf("batch        : %zu frames generated in %.2f s; localized in %.3f s (%.2f M frames/s from page-locked host memory); "
           "%.1f %% within 5 cells of the source; %llu kernel launches\n",
           F, t1 - t0, t3 - t2, F / (t3 - t2) / 1e6, 100.0 * near / F, (unsigned long long)at_kernel_launches());
    at_host_free(ctx, adc); at_host_free(ctx, lags); at_host_free(ctx, cell); free(truth);
    at_destroy(ctx);
    return 0;
}
