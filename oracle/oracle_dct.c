/* ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A from-scratch CPU restatement (closed form, double precision) of dct-carver's energy hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it;
 * the product (dct_carver_b200/csrc) never links or calls anything in oracle/.
 *
 * Parity pin: the reference has no tests, golden vectors or fixtures of its own.  This file is
 * therefore pinned against the reference ITSELF, compiled unmodified into oracle/_ref/libdctc_ref.so by
 * oracle/Makefile, in tests/test_oracle.py (bit-exact float32 for b=4,8,16; <= 1 float ulp for b=2), and
 * against the survey's known-answer sums.  The liblqr side (luma, seam DP) is "parity unpinned":
 * liblqr is not in the reference tree.
 *
 * What is restated, with the reference lines each piece follows:
 *   window gather + edge replication   /root/reference/src/render.c:122-132,146-152
 *   block-size dispatch / normalisation /root/reference/src/dct.c:77-94;
 *                                       b=8,16 orthonormal  src/fft2d/shrtdct.c:23-28,189-194;
 *                                       b=2,4 unnormalised  src/fft2d/fftsg2d.c:207-211
 *   last-arg-max scan + weighting       /root/reference/src/dct.c:96-110 (edge atoms: dct.c:10-43,56-73)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include "oracle_luma.h"

#define ORACLE_MAX_B 16

/* 1-D DCT-II basis.  b in {8,16}: sqrt(2/b)*s(k)*cos(pi*(j+1/2)*k/b), s(0)=1/sqrt(2) (shrtdct.c:23-28);
 * b in {2,4}: cos(pi*(j+1/2)*k/b) with no scaling (fftsg2d.c:207-211). */
void dctc_oracle_basis(int b, double *B /* b*b, row k */)
{
    const double pi = 3.14159265358979323846264338327950288;
    int k, j;
    for (k = 0; k < b; k++) {
        for (j = 0; j < b; j++) {
            double c = cos(pi * (j + 0.5) * k / b);
            if (b >= 8) c *= sqrt(2.0 / b) * (k == 0 ? sqrt(0.5) : 1.0);
            B[k * b + j] = c;
        }
    }
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* One pixel.  first_off = offset of the first window sample relative to (x,y):
 * carver path: -b/2+1 (render.c:146-147); preview path: -(C-1), C=(b-1)/2 (render.c:43-44, dct.h:8-9). */
static double pixel_energy(const double *luma, int w, int h, int x, int y, int b, const double *B,
                           double edges, double textures, int first_off, int *is_edge)
{
    double D[ORACLE_MAX_B][ORACLE_MAX_B], U[ORACLE_MAX_B][ORACLE_MAX_B];
    double best = 0.0;
    int a, c, k1, k2, bk1 = 0, bk2 = 0;
    /* D[a][c]: first index walks x, second walks y (render.c:150 stores data[i][j] with i the x offset) */
    for (a = 0; a < b; a++) {
        int xx = clampi(x + first_off + a, 0, w - 1);
        for (c = 0; c < b; c++) {
            int yy = clampi(y + first_off + c, 0, h - 1);
            D[a][c] = luma[(size_t) yy * w + xx];
        }
    }
    /* pass 1 along the first index, pass 2 along the second (shrtdct.c:62-89 then :90-117) */
    for (k1 = 0; k1 < b; k1++)
        for (c = 0; c < b; c++) {
            double s = 0.0;
            for (a = 0; a < b; a++) s += B[k1 * b + a] * D[a][c];
            U[k1][c] = s;
        }
    for (k1 = 0; k1 < b; k1++)
        for (k2 = 0; k2 < b; k2++) {
            double s = 0.0, v;
            for (c = 0; c < b; c++) s += U[k1][c] * B[k2 * b + c];
            v = fabs(s);
            if (best <= v && (k1 || k2)) { best = v; bk1 = k1; bk2 = k2; }   /* last max wins, dct.c:103 */
        }
    *is_edge = (bk1 == 0 && bk2 == 1) || (bk1 == 1 && bk2 == 0);             /* dct.c:10-43 */
    return *is_edge ? best * edges : best * textures;                         /* dct.c:109 */
}

typedef struct {
    const double *luma;
    int w, h, b, y0, y1, first_off;
    float edges, textures;
    float *out;
    double *out_d;
    uint8_t *cls;
} Job;

static void *worker(void *arg)
{
    Job *j = (Job *) arg;
    double B[ORACLE_MAX_B * ORACLE_MAX_B];
    int x, y, e;
    dctc_oracle_basis(j->b, B);
    for (y = j->y0; y < j->y1; y++)
        for (x = 0; x < j->w; x++) {
            double v = pixel_energy(j->luma, j->w, j->h, x, y, j->b, B, (double) j->edges, (double) j->textures,
                                    j->first_off, &e);
            size_t i = (size_t) y * j->w + x;
            if (j->out) j->out[i] = (float) v;          /* gfloat return, render.c:134,155-156 */
            if (j->out_d) j->out_d[i] = v;
            if (j->cls) j->cls[i] = (uint8_t) e;
        }
    return NULL;
}

static int run_rows(const double *luma, int w, int h, int b, float edges, float textures, int first_off,
                    float *out, double *out_d, uint8_t *cls, int y_begin, int y_end, int nthreads)
{
    int t, n = nthreads < 1 ? 1 : nthreads, rows = y_end - y_begin;
    pthread_t *tid;
    Job *jobs;
    if (b != 2 && b != 4 && b != 8 && b != 16) return -2;
    if (rows <= 0 || w <= 0) return 0;
    if (n > rows) n = rows;
    tid = (pthread_t *) malloc(sizeof(pthread_t) * n);
    jobs = (Job *) malloc(sizeof(Job) * n);
    for (t = 0; t < n; t++) {
        Job *j = &jobs[t];
        j->luma = luma; j->w = w; j->h = h; j->b = b; j->first_off = first_off;
        j->edges = edges; j->textures = textures; j->out = out; j->out_d = out_d; j->cls = cls;
        j->y0 = y_begin + (int) ((long long) rows * t / n);
        j->y1 = y_begin + (int) ((long long) rows * (t + 1) / n);
        if (n == 1) worker(j); else pthread_create(&tid[t], NULL, worker, j);
    }
    if (n > 1) for (t = 0; t < n; t++) pthread_join(tid[t], NULL);
    free(tid); free(jobs);
    return 0;
}

/* Carver-path energy (what liblqr's build_emap obtains from dct_pixel_energy) for rows [y_begin,y_end). */
int dctc_oracle_energy_rows(const double *luma, int w, int h, int blocksize, float edges, float textures,
                            float *out, uint8_t *cls_or_null, int y_begin, int y_end, int nthreads)
{
    return run_rows(luma, w, h, blocksize, edges, textures, -blocksize / 2 + 1, out, NULL, cls_or_null,
                    y_begin, y_end, nthreads);
}

int dctc_oracle_luma(const uint8_t *img, int w, int h, int channels, size_t pitch, double *luma)
{
    if (channels < 1 || channels > 4) return -2;
    dctc_oracle_luma_plane(img, w, h, channels, pitch, luma);
    return 0;
}

int dctc_oracle_energy_image(const uint8_t *img, int w, int h, int channels, size_t pitch, int blocksize,
                             float edges, float textures, float *out, uint8_t *cls_or_null, int nthreads)
{
    double *luma;
    int rc;
    if (channels < 1 || channels > 4) return -2;
    luma = (double *) malloc(sizeof(double) * (size_t) w * h);
    if (!luma) return -1;
    dctc_oracle_luma_plane(img, w, h, channels, pitch, luma);
    rc = dctc_oracle_energy_rows(luma, w, h, blocksize, edges, textures, out, cls_or_null, 0, h, nthreads);
    free(luma);
    return rc;
}

/* Preview-path operator (render.c:31-60): window [-(C-1), b-C], C=(b-1)/2 integer division (dct.h:8-9),
 * samples are the 0..255 BT.601 luminance bytes of render.h:5, output is double and NOT transposed:
 * data[ii][jj-left] = rows[ii][clamp(jj)]  =>  first index walks y.  Since the edge LUT and the basis are
 * symmetric, transposing only swaps T[k1][k2] <-> T[k2][k1]; the last-arg-max scan order then differs, which
 * is visible only on exact ties between the two classes. */
int dctc_oracle_preview_energy(const uint8_t *lum8, int w, int h, int blocksize, float edges, float textures,
                               double *out)
{
    double *plane, *lumT;
    int x, y, rc, C = (blocksize - 1) / 2;
    if (w <= 0 || h <= 0) return 0;
    /* evaluate on the transposed plane so that "first index" walks y as in render.c:47-51 */
    plane = (double *) malloc(sizeof(double) * (size_t) w * h);
    lumT = (double *) malloc(sizeof(double) * (size_t) w * h);
    if (!plane || !lumT) { free(plane); free(lumT); return -1; }
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++) lumT[(size_t) x * h + y] = (double) lum8[(size_t) y * w + x];
    rc = run_rows(lumT, h, w, blocksize, edges, textures, -(C - 1), NULL, plane, NULL, 0, w, 1);
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++) out[(size_t) y * w + x] = plane[(size_t) x * h + y];
    free(plane); free(lumT);
    return rc;
}

/* BT.601 studio-swing byte luminance of the preview path (render.h:5, render.c:62-79). */
void dctc_oracle_preview_luminance(const uint8_t *img, int w, int h, int channels, size_t pitch, uint8_t *lum8)
{
    int x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++) {
            const uint8_t *p = img + (size_t) y * pitch + (size_t) x * channels;
            lum8[(size_t) y * w + x] = channels == 1 ? p[0]
                : (uint8_t) (16.0 + p[0] * 0.2568 + p[1] * 0.5041 + p[2] * 0.0979);
        }
}
