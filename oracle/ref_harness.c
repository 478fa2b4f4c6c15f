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
