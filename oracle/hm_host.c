/* oracle/hm_host.c -- TEST INFRASTRUCTURE.
 * The reference's presentation layer on the host: src/components/vga/vga.h is included UNMODIFIED below (it pulls in
 * sample_compute.h and the four vga_*.h views), behind the same SDK shim as sc_host.c and with the VGA primitives of
 * lib/vga/vga16_graphics.h defined as no-ops (fillRect records what it is asked to paint).  That makes
 * vga_init_heatmap (vga_heatmap.h:48-93: the lag look-up table) and vga_draw_heatmap (vga_heatmap.h:95-135: likelihood
 * map, thresholds, colour classes) callable as they are, which pins SURVEY section 8 rows a19 / a20 to the reference
 * itself.  Built by oracle/Makefile into _ref/libat_ref_hm.so; tests/golden/make_heatmap_golden.py turns its outputs
 * into the committed fixture tests/golden/ref_heatmap.npz.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sc_shim/sc_sdk.h"

volatile uint8_t dma_sample_array[3];      /* ref: components/dma_sampler.c:3 */
absolute_time_t get_absolute_time(void) { return 0; }
uint64_t time_us_64(void) { return 0; }
void busy_wait_until(absolute_time_t t) { (void)t; }

#include <components/vga/vga.h>             /* the reference's own code, all of L4 and sample_compute.h */

/* ---- VGA primitives (lib/vga/vga16_graphics.h:38-62): nothing is drawn; fillRect keeps a log of the cells it paints */
static int g_fill_calls;
static signed char g_fill_color[HEATMAP_HEIGHT][HEATMAP_WIDTH];
void fillRect(short x, short y, short w, short h, char color)
{
    const int cx = ((x - POS_ORIG_X) >> MAP_SCALE_BITS) + POS_HALF_W, cy = POS_HALF_H - ((y - POS_ORIG_Y) >> MAP_SCALE_BITS);
    (void)w; (void)h;
    g_fill_calls++;
    if (cx >= 0 && cx < HEATMAP_WIDTH && cy >= 0 && cy < HEATMAP_HEIGHT) g_fill_color[cy][cx] = color;
}
void initVGA(void) {}
void drawPixel(short x, short y, char c) { (void)x; (void)y; (void)c; }
void drawVLine(short x, short y, short h, char c) { (void)x; (void)y; (void)h; (void)c; }
void drawHLine(short x, short y, short w, char c) { (void)x; (void)y; (void)w; (void)c; }
void drawLine(short x0, short y0, short x1, short y1, char c) { (void)x0; (void)y0; (void)x1; (void)y1; (void)c; }
void drawRect(short x, short y, short w, short h, char c) { (void)x; (void)y; (void)w; (void)h; (void)c; }
void drawCircle(short x0, short y0, short r, char c) { (void)x0; (void)y0; (void)r; (void)c; }
void fillCircle(short x0, short y0, short r, char c) { (void)x0; (void)y0; (void)r; (void)c; }
void drawRoundRect(short x, short y, short w, short h, short r, char c) { (void)x; (void)y; (void)w; (void)h; (void)r; (void)c; }
void fillRoundRect(short x, short y, short w, short h, short r, char c) { (void)x; (void)y; (void)w; (void)h; (void)r; (void)c; }
void drawChar(short x, short y, unsigned char ch, char c, char bg, unsigned char size) { (void)x; (void)y; (void)ch; (void)c; (void)bg; (void)size; }
void setCursor(short x, short y) { (void)x; (void)y; }
void setTextColor(char c) { (void)c; }
void setTextColor2(char c, char bg) { (void)c; (void)bg; }
void setTextSize(unsigned char s) { (void)s; }
void setTextWrap(char w) { (void)w; }
void tft_write(unsigned char c) { (void)c; }
void writeString(char *str) { (void)str; }
void drawCharBig(short x, short y, unsigned char ch, char c, char bg) { (void)x; (void)y; (void)ch; (void)c; (void)bg; }
void writeStringBig(char *str) { (void)str; }
void setTextColorBig(char a, char b) { (void)a; (void)b; }

/* ---- exported to the golden-vector generator */
void hm_dims(int *w, int *h, int *n_lags, int *colors /* white, green, red, blue, black */)
{
    *w = HEATMAP_WIDTH; *h = HEATMAP_HEIGHT; *n_lags = CORRELATION_BUFFER_SIZE;
    colors[0] = WHITE; colors[1] = GREEN; colors[2] = RED; colors[3] = BLUE; colors[4] = BLACK;
}
void hm_init(float *mic_xy /* [3][2] */, uint8_t *lut /* [3][H][W]: ab, ac, bc */)
{
    microphones_init();                     /* components/microphones.c, as main.c:57 does before anything else */
    vga_init_heatmap();                     /* vga_heatmap.h:48-93 */
    mic_xy[0] = mic_a_location.x; mic_xy[1] = mic_a_location.y; mic_xy[2] = mic_b_location.x; mic_xy[3] = mic_b_location.y;
    mic_xy[4] = mic_c_location.x; mic_xy[5] = mic_c_location.y;
    memcpy(lut, heat_idx_ab, sizeof heat_idx_ab);
    memcpy(lut + sizeof heat_idx_ab, heat_idx_ac, sizeof heat_idx_ac);
    memcpy(lut + 2 * sizeof heat_idx_ab, heat_idx_bc, sizeof heat_idx_bc);
}
/* one call of vga_draw_heatmap on the given (averaged, re-weighted) curves; colors = the colour of every cell after it.
 * Returns 0 when the fillRect log agrees with heat_colors on every cell that was painted. */
int hm_draw(const int64_t *curves /* [3][93]: ab, ac, bc */, uint8_t *colors /* [H][W] */)
{
    memcpy(corr_ab.correlations, curves, sizeof corr_ab.correlations);
    memcpy(corr_ac.correlations, curves + CORRELATION_BUFFER_SIZE, sizeof corr_ac.correlations);
    memcpy(corr_bc.correlations, curves + 2 * CORRELATION_BUFFER_SIZE, sizeof corr_bc.correlations);
    memset(heat_colors, -1, sizeof heat_colors);            /* no colour: every cell is repainted */
    memset(g_fill_color, -2, sizeof g_fill_color);
    g_fill_calls = 0;
    vga_draw_heatmap();                     /* vga_heatmap.h:95-158 */
    int bad = g_fill_calls != HEATMAP_WIDTH * HEATMAP_HEIGHT;
    for (int y = 0; y < HEATMAP_HEIGHT; y++)
        for (int x = 0; x < HEATMAP_WIDTH; x++) {
            colors[y * HEATMAP_WIDTH + x] = (uint8_t)heat_colors[y][x];
            if (g_fill_color[y][x] != heat_colors[y][x]) bad++;
        }
    return bad;
}
