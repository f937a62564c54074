/* oracle/sc_host.c -- TEST INFRASTRUCTURE.
 * Runs the reference's own protothread_sample_and_compute (src/sample_compute.h, included unmodified below) on
 * the host, fed by a recorded ADC triple stream.  Built twice by oracle/Makefile:
 *   _ref/sc_ref   linked with the reference's buffer.c / rolling_buffer.c / correlations.c
 *   _ref/sc_b200  linked with libat_b200.so (the drop-in symbols) -- nothing else differs.
 * Both print one line per gated frame; tests/test_gpu_dropin_protothread.py requires identical output.
 *
 *   usage: sc_xxx stream.bin      (stream.bin = uint8 triples A,B,C, one per 20 us tick)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sc_shim/sc_sdk.h"

volatile uint8_t dma_sample_array[3];      /* ref: components/dma_sampler.c:3 (the capture side we replace) */
static uint8_t *g_stream;
static size_t g_ticks, g_pos;
static uint64_t g_now_us;

absolute_time_t get_absolute_time(void) { return g_now_us; }
uint64_t time_us_64(void) { return g_now_us; }

static void next_triple(void)
{
    if (g_pos >= g_ticks) { printf("END ticks=%zu\n", g_pos); exit(0); }
    dma_sample_array[0] = g_stream[3 * g_pos]; dma_sample_array[1] = g_stream[3 * g_pos + 1]; dma_sample_array[2] = g_stream[3 * g_pos + 2];
    g_pos++;
}
/* the capture loop sleeps here once per sample (sample_compute.h:98): time advances, the next triple appears */
void busy_wait_until(absolute_time_t t) { g_now_us = t; next_triple(); }

#include <sample_compute.h>                 /* the reference's orchestration code and its static state */

static unsigned long long checksum(const struct correlations_t *c)
{
    unsigned long long s = 0;
    for (int i = 0; i < CORRELATION_BUFFER_SIZE; i++) s = s * 31u + (unsigned long long)c->correlations[i];
    return s;
}

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s stream.bin\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    g_stream = malloc((size_t)n); g_ticks = (size_t)n / 3;
    if (fread(g_stream, 1, (size_t)n, f) != (size_t)n) return 2;
    fclose(f);
    g_now_us = 1000000;
    next_triple();
    PT_SEM_INIT(&vga_semaphore, 0);         /* ref: main.c:67-68 */
    PT_SEM_INIT(&load_audio_semaphore, 1);
    static struct pt pt;
    PT_INIT(&pt);
    int event = 0;
    for (;;) {
        protothread_sample_and_compute(&pt);    /* returns when it waits on load_audio_semaphore (sample_compute.h:145) */
        if (vga_semaphore.count > 0) {          /* we stand in for protothread_vga_debug (vga_debug.h:22-33) */
            vga_semaphore.count--;
            printf("event %d tick %zu new %d %d %d avg %d %d %d sum_new %llu %llu %llu sum_avg %llu %llu %llu\n", event++, g_pos,
                   new_corr_ab.best_shift, new_corr_ac.best_shift, new_corr_bc.best_shift,
                   corr_ab.best_shift, corr_ac.best_shift, corr_bc.best_shift,
                   checksum(&new_corr_ab), checksum(&new_corr_ac), checksum(&new_corr_bc),
                   checksum(&corr_ab), checksum(&corr_ac), checksum(&corr_bc));
            g_now_us += 30000;                  /* the display thread's drawing time */
            load_audio_semaphore.count++;       /* PT_SEM_SIGNAL(load_audio_semaphore), vga_debug.h:32 */
        }
    }
}
