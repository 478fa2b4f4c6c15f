/* TEST INFRASTRUCTURE (oracle/_ref).  Drives the reference's OWN, unmodified hot path
 *   dct_pixel_energy            /root/reference/src/render.c:134-157
 *   dctNxN / weighted_max_...   /root/reference/src/dct.c:77-110
 *   ddct8x8s/ddct16x16s/ddct2d  /root/reference/src/fft2d/{shrtdct,fftsg2d,fftsg}.c
 * over a whole image, the way liblqr's build_emap does (for y<h, for x<w: en = nrg(x,y,w,h,rw,extra)).
 * It supplies the three symbols the hot path needs from outside: lqr_rwindow_read,
 * lqr_rwindow_get_radius (liblqr) and error() (the reference's interface.c:570-573, GTK-only).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#include <stdint.h>
#include <stdlib.h>
#include <math.h>
#include <pthread.h>
#include "shim/dctc_shim_types.h"
#include "oracle_luma.h"

/* reference types/functions we call (declared, not copied: layouts restated from render.h:9-18) */
typedef struct {
    gfloat edges;
    gfloat textures;
    gint blocksize;
    int *ip;
    double *w;
    double **data;
} RefEnergyParameters;
gfloat dct_pixel_energy(gint x, gint y, gint w, gint h, LqrReadingWindow *rw, gpointer extra_data);
int *alloc_1d_int(int n1);
void free_1d_int(int *i);
double *alloc_1d_double(int n1);
void free_1d_double(double *d);
double **alloc_2d_double(int n1, int n2);
void free_2d_double(double **dd);

struct ShimLqrReadingWindow_ {
    const double *luma;
    int w, h, x, y, radius;
};

gdouble lqr_rwindow_read(LqrReadingWindow *rw, gint dx, gint dy, gint channel)
{
    int xx = rw->x + dx, yy = rw->y + dy;
    (void) channel;
    if (dx < -rw->radius || dx > rw->radius || dy < -rw->radius || dy > rw->radius) return 0.0;
    if (xx < 0 || xx >= rw->w || yy < 0 || yy >= rw->h) return 0.0;
    return rw->luma[(size_t) yy * rw->w + xx];
}

gint lqr_rwindow_get_radius(LqrReadingWindow *rw) { return rw->radius; }

static int g_error_count = 0;
void error(const gchar *message) { (void) message; g_error_count++; }
int dctc_ref_error_count(void) { return g_error_count; }

typedef struct {
    const double *luma;
    int w, h, b, y0, y1;
    float edges, textures;
    float *out;
} RefJob;

static void *ref_worker(void *arg)
{
    RefJob *job = (RefJob *) arg;
    RefEnergyParameters ep;
    LqrReadingWindow rw;
    int x, y;
    /* scratch exactly as render.c:296-305 fills it */
    ep.edges = job->edges;
    ep.textures = job->textures;
    ep.blocksize = job->b;
    ep.ip = alloc_1d_int(2 + (int) sqrt(job->b / 2 + 0.5));
    ep.w = alloc_1d_double(job->b * 3 / 2);
    ep.data = alloc_2d_double(job->b, job->b);
    ep.ip[0] = 0;
    rw.luma = job->luma; rw.w = job->w; rw.h = job->h; rw.radius = job->b / 2;
    for (y = job->y0; y < job->y1; y++) {
        for (x = 0; x < job->w; x++) {
            rw.x = x; rw.y = y;
            job->out[(size_t) y * job->w + x] = dct_pixel_energy(x, y, job->w, job->h, &rw, &ep);
        }
    }
    free_1d_int(ep.ip);
    free_1d_double(ep.w);
    free_2d_double(ep.data);
    return NULL;
}

/* Energy for rows [y_begin, y_end) of a w*h luma plane; out is indexed as the full w*h map.
 * nthreads > 1 splits the rows, one private EnergyParameters scratch per thread (the reference's
 * callback is not re-entrant over shared scratch, render.c:140,154). */
int dctc_ref_energy_rows(const double *luma, int w, int h, int blocksize, float edges, float textures,
                         float *out, int y_begin, int y_end, int nthreads)
{
    int t, n = nthreads < 1 ? 1 : nthreads, rows = y_end - y_begin;
    pthread_t *tid;
    RefJob *jobs;
    if (rows <= 0) return 0;
    if (n > rows) n = rows;
    tid = (pthread_t *) malloc(sizeof(pthread_t) * n);
    jobs = (RefJob *) malloc(sizeof(RefJob) * n);
    for (t = 0; t < n; t++) {
        jobs[t].luma = luma; jobs[t].w = w; jobs[t].h = h; jobs[t].b = blocksize;
        jobs[t].edges = edges; jobs[t].textures = textures; jobs[t].out = out;
        jobs[t].y0 = y_begin + (int) ((long long) rows * t / n);
        jobs[t].y1 = y_begin + (int) ((long long) rows * (t + 1) / n);
        if (n == 1) ref_worker(&jobs[t]);
        else pthread_create(&tid[t], NULL, ref_worker, &jobs[t]);
    }
    if (n > 1) for (t = 0; t < n; t++) pthread_join(tid[t], NULL);
    free(tid); free(jobs);
    return 0;
}

/* Same, from the interleaved 8-bit buffer lqr_carver_new() receives (render.c:310-312). */
int dctc_ref_energy_image(const uint8_t *img, int w, int h, int channels, size_t pitch, int blocksize,
                          float edges, float textures, float *out, int nthreads)
{
    double *luma = (double *) malloc(sizeof(double) * (size_t) w * h);
    if (!luma) return -1;
    dctc_oracle_luma_plane(img, w, h, channels, pitch, luma);
    dctc_ref_energy_rows(luma, w, h, blocksize, edges, textures, out, 0, h, nthreads);
    free(luma);
    return 0;
}

/* ---- preview path: the reference's own dct_energy_preview_rows / convert_row_to_luminance / normalize_image
 * (render.c:31-109, made linkable by oracle/Makefile), driven the way dct_energy_preview does (render.c:459-482):
 * a sliding set of `blocksize` luminance rows, row i of the first set = image row clamp(i - (CENTER_ROW - 1)). */
typedef struct {            /* PlugInVals, layout restated from main.h:12-22 */
    gfloat edges;
    gfloat textures;
    gint blocksize;
    gint seams_number;
    gint new_layer, resize_canvas, output_energy, output_seams, vertically;
} RefPlugInVals;
void dct_energy_preview_rows(RefPlugInVals *vals, guchar **current_rows, gdouble *energy_image, gint row_number, gint width);
void convert_row_to_luminance(guchar *inrow, guchar *outrow, gint channels, gint width);
void normalize_image(gdouble *energy_image, guchar *output_image, gint height, gint width, gint channels);

int dctc_ref_preview(const uint8_t *img, int w, int h, int channels, size_t pitch, int blocksize, float edges,
                     float textures, double *energy, uint8_t *out_image)
{
    RefPlugInVals vals;
    guchar **rows;
    int i, row_number, center = (blocksize - 1) / 2;   /* CENTER_ROW, dct.h:8 */
    vals.edges = edges; vals.textures = textures; vals.blocksize = blocksize; vals.seams_number = 0;
    vals.new_layer = vals.resize_canvas = vals.output_energy = vals.output_seams = vals.vertically = 0;
    rows = (guchar **) malloc(sizeof(guchar *) * blocksize);
    if (!rows) return -1;
    for (i = 0; i < blocksize; i++) {
        int yy = i - (center - 1);
        yy = yy < 0 ? 0 : (yy > h - 1 ? h - 1 : yy);
        rows[i] = (guchar *) malloc(w);
        convert_row_to_luminance((guchar *) img + (size_t) yy * pitch, rows[i], channels, w);
    }
    for (row_number = 0; row_number < h; row_number++) {
        int yy = row_number + blocksize - (center - 1);
        dct_energy_preview_rows(&vals, rows, energy, row_number, w);
        free(rows[0]);
        for (i = 1; i < blocksize; i++) rows[i - 1] = rows[i];
        if (yy > h - 1) yy = h - 1;
        rows[blocksize - 1] = (guchar *) malloc(w);
        convert_row_to_luminance((guchar *) img + (size_t) yy * pitch, rows[blocksize - 1], channels, w);
    }
    if (out_image) normalize_image(energy, out_image, h, w, channels);
    for (i = 0; i < blocksize; i++) free(rows[i]);
    free(rows);
    return 0;
}
