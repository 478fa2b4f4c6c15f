/* TEST INFRASTRUCTURE (oracle).  Luma that liblqr's LQR_ER_LUMA reader hands to the energy callback.
 *
 * liblqr (module lqr-1, pinned ">= 0.5.0" by /root/reference/configure.in:64-67) is NOT in the reference tree
 * and not installed here, so this restates its published behaviour [liblqr 0.4.x lqr_energy.c, from memory]:
 * every 8-bit channel is normalised to [0,1] as v/255.0 in double, RGB is combined with the Rec.709
 * weights 0.2126/0.7152/0.0722, and the result is multiplied by alpha/255 when an alpha channel exists
 * (2 = grey+alpha, 4 = RGBA).  The call site it feeds is /root/reference/src/render.c:150
 * (lqr_rwindow_read(rw, ii, jj, 0) with reader type LQR_ER_LUMA registered at render.c:314-315).
 * Parity at this boundary is UNPINNED by the reference (it has no tests); see DESIGN.md.
 */
#ifndef DCTC_ORACLE_LUMA_H
#define DCTC_ORACLE_LUMA_H
#include <stddef.h>
#include <stdint.h>

static inline double dctc_oracle_luma_px(const uint8_t *p, int channels)
{
    double v;
    if (channels >= 3) {
        double r = (double) p[0] / 255.0, g = (double) p[1] / 255.0, b = (double) p[2] / 255.0;
        v = 0.2126 * r + 0.7152 * g + 0.0722 * b;
    } else {
        v = (double) p[0] / 255.0;
    }
    if (channels == 2 || channels == 4) v *= (double) p[channels - 1] / 255.0;
    return v;
}

static inline void dctc_oracle_luma_plane(const uint8_t *img, int w, int h, int channels, size_t pitch,
                                          double *luma)
{
    int x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++)
            luma[(size_t) y * w + x] = dctc_oracle_luma_px(img + (size_t) y * pitch + (size_t) x * channels, channels);
}
#endif
