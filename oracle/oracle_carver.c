/* ORACLE — TEST INFRASTRUCTURE.  Independent, deliberately naive restatement of the seam loop that
 * lqr_carver_resize runs for dct-carver (/root/reference/src/render.c:313 delta_x=1 rigidity=0, :377 resize):
 * for every seam the energy map is recomputed FROM SCRATCH with the oracle operator, the cumulative map is
 * rebuilt from scratch, the seam ends at the leftmost minimum of the last row and is back-tracked through the
 * first strict minimum among the (up to) three parents.  [liblqr 0.4.x semantics, from memory — liblqr is not in
 * the reference tree: "parity unpinned" at this boundary.]  The product's incremental carver
 * (dct_carver_b200/host/dctc_lqr.c + K2) must reproduce these seams exactly when fed the same energies. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int dctc_oracle_energy_image(const uint8_t *img, int w, int h, int channels, size_t pitch, int blocksize,
                             float edges, float textures, float *out, uint8_t *cls_or_null, int nthreads);

/* energy_override: when non-NULL it is called instead of the oracle operator (lets tests feed GPU energies) */
typedef int (*oracle_energy_fn)(const uint8_t *img, int w, int h, int ch, float *out, void *user);

int dctc_oracle_retarget_width(const uint8_t *img, int w, int h, int ch, int blocksize, float edges, float textures,
                               int n_seams, int *seams_out /* n_seams*h */, uint8_t *img_out /* (w-n_seams)*h*ch */,
                               oracle_energy_fn energy_override, void *user, int nthreads)
{
    uint8_t *cur = (uint8_t *) malloc((size_t) w * h * ch);
    float *en = (float *) malloc(sizeof(float) * (size_t) w * h);
    float *m = (float *) malloc(sizeof(float) * (size_t) w * h);
    int s, x, y, cw = w, rc = -1;
    if (!cur || !en || !m || n_seams >= w) goto done;
    memcpy(cur, img, (size_t) w * h * ch);
    for (s = 0; s < n_seams; s++) {
        int *seam = seams_out + (size_t) s * h;
        if (energy_override) rc = energy_override(cur, cw, h, ch, en, user);
        else rc = dctc_oracle_energy_image(cur, cw, h, ch, (size_t) cw * ch, blocksize, edges, textures, en, NULL, nthreads);
        if (rc) goto done;
        for (x = 0; x < cw; x++) m[x] = en[x];
        for (y = 1; y < h; y++)
            for (x = 0; x < cw; x++) {
                int lo = x > 0 ? x - 1 : 0, hi = x < cw - 1 ? x + 1 : cw - 1, x1;
                float best = m[(size_t) (y - 1) * cw + lo];
                for (x1 = lo + 1; x1 <= hi; x1++)
                    if (m[(size_t) (y - 1) * cw + x1] < best) best = m[(size_t) (y - 1) * cw + x1];
                m[(size_t) y * cw + x] = en[(size_t) y * cw + x] + best;
            }
        x = 0;
        for (y = 1; y < cw; y++)
            if (m[(size_t) (h - 1) * cw + y] < m[(size_t) (h - 1) * cw + x]) x = y;
        seam[h - 1] = x;
        for (y = h - 1; y > 0; y--) {
            int c = seam[y], lo = c > 0 ? c - 1 : 0, hi = c < cw - 1 ? c + 1 : cw - 1, x1, arg = lo;
            for (x1 = lo + 1; x1 <= hi; x1++)
                if (m[(size_t) (y - 1) * cw + x1] < m[(size_t) (y - 1) * cw + arg]) arg = x1;
            seam[y - 1] = arg;
        }
        for (y = 0; y < h; y++) {   /* carve: compact to pitch cw-1 */
            memmove(cur + (size_t) y * (cw - 1) * ch, cur + (size_t) y * cw * ch, (size_t) seam[y] * ch);
            memmove(cur + ((size_t) y * (cw - 1) + seam[y]) * ch, cur + ((size_t) y * cw + seam[y] + 1) * ch,
                    (size_t) (cw - 1 - seam[y]) * ch);
        }
        cw--;
    }
    if (img_out) memcpy(img_out, cur, (size_t) cw * h * ch);
    rc = 0;
done:
    free(cur); free(en); free(m);
    return rc;
}
