/* dctc.h — C ABI of the B200-native DCT-Carver energy hot path.
 *
 * This is the drop-in boundary for ONE path of avivrosenberg/dct-carver: the per-pixel block-DCT energy
 * map that the plug-in registers with liblqr as a custom energy function.  Every entry point names the
 * reference interface it replaces (paths relative to the reference tree).  Plain C: pointers and sizes
 * only, no C++/torch types.  All compute runs in hand-written CUDA kernels for sm_100a
 * (dct_carver_b200/csrc); there is NO CPU fallback: without a CUDA device every compute call returns
 * DCTC_ERR_NO_DEVICE / DCTC_ERR_CUDA.
 *
 * Threading: one caller per context (the reference's callback is single-threaded and non re-entrant,
 * src/render.c:140,154).  Use one context per GPU / stream.
 */
#ifndef DCTC_H
#define DCTC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCTC_VERSION 1

/* status codes (the reference has no error channel: bad blocksize -> error()+untransformed data,
 * src/dct.c:89-92; OOM -> exit(1), src/fft2d/alloc.c:5-10.  Here every batch call returns int.) */
enum {
    DCTC_OK = 0,
    DCTC_ERR_INVALID = -1,    /* NULL pointer, non-positive size, channels not in 1..4, pitch too small */
    DCTC_ERR_BLOCKSIZE = -2,  /* blocksize not in {2,4,8,16}  (src/dct.c:77-93, UI combo src/interface.c:281) */
    DCTC_ERR_NOMEM = -3,      /* cudaMalloc / cudaMallocHost failed */
    DCTC_ERR_CUDA = -4,       /* a CUDA call failed; see dctc_last_cuda_error() */
    DCTC_ERR_NO_DEVICE = -5,  /* no usable CUDA device */
    DCTC_ERR_STATE = -6,      /* carver call without a loaded image / seam outside the current width */
    DCTC_ERR_UNSUPPORTED = -7 /* requested kernel variant does not exist for this block size */
};

/* The reference's operator state, src/render.h:9-18, filled at src/render.c:296-305.
 * Layout is kept bit-for-bit so a caller's existing struct can be passed as is.  ip / w / data are the
 * Ooura scratch areas of the CPU path; the GPU path ignores them (scratch lives in registers/shared memory)
 * and never frees caller memory (ownership: src/render.c:413-416). */
typedef struct DctcEnergyParameters_ {
    float edges;
    float textures;
    int blocksize;
    int *ip;
    double *w;
    double **data;
} DctcEnergyParameters;

typedef struct dctc_context dctc_context;

/* Superset handed to liblqr as extra_data (src/render.c:314-315): the prefix IS EnergyParameters, so the
 * reference's own dct_pixel_energy could still read it; the tail carries the GPU context. */
typedef struct DctcCarverEnergyParams_ {
    DctcEnergyParameters base;
    dctc_context *gpu;
} DctcCarverEnergyParams;

/* kernel selection for block size 8 (other block sizes always use the FP32 CUDA-core kernel) */
enum {
    DCTC_KERNEL_AUTO = 0,
    DCTC_KERNEL_FP32_TILE = 1,   /* generic shared-memory tile kernel, all block sizes */
    DCTC_KERNEL_FP32_MARCH = 2,  /* b=8 register-sliding column march */
    DCTC_KERNEL_TC_SPLIT = 3,    /* tcgen05 Toeplitz GEMM, fp16 hi/lo split operands, fp32 accumulate (b = 8; b = 16 by default) */
    DCTC_KERNEL_FP32_STREAM = 4  /* b=2, 4 streaming register-march kernel (what AUTO picks for them on aligned 1- or 3-channel rows) */
};

/* ---- lifetime -------------------------------------------------------------------------------------- */

/* Creates a context on CUDA device `device` with its own stream.  Replaces the scratch allocation of
 * src/render.c:296-305 (alloc_1d_int / alloc_1d_double / alloc_2d_double). */
int dctc_create(dctc_context **ctx, int device);
/* Frees device buffers, pinned staging and the stream.  Counterpart of the frees at src/render.c:413-416. */
void dctc_destroy(dctc_context *ctx);
/* edges / textures / blocksize as PlugInVals carries them (src/main.h:12-22, src/main.c:151-153).  With a carver
 * session loaded, a change of any of the three rebuilds the session's resident energy map with the new operator
 * (liblqr: lqr_carver_set_energy_function invalidates the map, src/render.c:314-315). */
int dctc_set_params(dctc_context *ctx, const DctcEnergyParameters *params);
int dctc_set_kernel(dctc_context *ctx, int kernel);
int dctc_last_cuda_error(const dctc_context *ctx);
const char *dctc_strerror(int status);
int dctc_version(void);
/* number of CUDA devices visible, or a negative status */
int dctc_device_count(void);
/* CUDA stream of the context as an opaque handle (cudaStream_t) */
void *dctc_stream(dctc_context *ctx);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
unsigned long long dctc_launch_count(const dctc_context *ctx);

/* ---- K1: full energy map ----------------------------------------------------------------------------
 * Replaces liblqr's build_emap loop "for y<h for x<w: en = dct_pixel_energy(x,y,w,h,rw,extra)"
 * (callback src/render.c:134-157; dctNxN src/dct.c:77-94; weighted_max_dct_correlation src/dct.c:96-110)
 * over the interleaved 8-bit buffer that lqr_carver_new() receives (src/render.c:310-312), with the
 * LQR_ER_LUMA reader fused in.  out[y*w+x] is the gfloat the callback would return.
 * Device-pointer variants: every row must be readable for its full pitch_bytes (the kernels stage rows in 16-byte
 * chunks up to the pitch), i.e. the buffer spans h*pitch_bytes; the host-buffer variants copy exactly w*channels
 * bytes per row into their own 16-byte-pitched staging. */
int dctc_energy_full(dctc_context *ctx, const uint8_t *img, int w, int h, int channels, size_t pitch_bytes,
                     float *out);

/* Same with device-resident input/output (no copies); out_pitch in floats.  Asynchronous on the context
 * stream unless `sync` is non-zero. */
int dctc_energy_full_dev(dctc_context *ctx, const uint8_t *d_img, int w, int h, int channels,
                         size_t pitch_bytes, float *d_out, size_t out_pitch, int sync);

/* Batch of n_frames equally sized frames (BASELINE config 4), one launch. */
int dctc_energy_batch_dev(dctc_context *ctx, const uint8_t *d_imgs, int n_frames, size_t frame_stride_bytes,
                          int w, int h, int channels, size_t pitch_bytes, float *d_out,
                          size_t out_frame_stride, size_t out_pitch, int sync);

/* Row band of a taller image (BASELINE config 5).  d_band holds `band_rows` rows; d_top points to the
 * top_rows image rows directly above the band (in image order, so its last row touches the band) and d_bot to
 * the bot_rows rows directly below.  Either may be NULL with 0 rows at the image edge, where
 * the reference's edge replication (src/render.c:122-132) applies.  d_top / d_bot may point into a peer
 * GPU's memory (NVLink P2P): the kernel then reads the halo straight from the neighbour, no exchange step.
 * Needs blocksize/2-1 rows above and blocksize/2 rows below for an exact result. */
int dctc_energy_band_dev(dctc_context *ctx, const uint8_t *d_band, int w, int band_rows, int channels,
                         size_t pitch_bytes, const uint8_t *d_top, int top_rows, size_t top_pitch_bytes,
                         const uint8_t *d_bot, int bot_rows, size_t bot_pitch_bytes, float *d_out,
                         size_t out_pitch, int sync);

/* Same for a caller that knows where the band sits in the whole image: band_y0 = image row of the band's first row.
 * The block-size-16 tensor-core kernel anchors its 16-row accumulation steps to the image's row grid, so with band_y0
 * given the band's map is bit-identical to the rows a single full-image call produces, however the image is cut
 * (dctc_energy_band_dev is band_y0 = 0; every other kernel ignores it).  The multi-GPU layer below passes it. */
int dctc_energy_band_dev_at(dctc_context *ctx, const uint8_t *d_band, int w, int band_rows, int band_y0, int channels,
                            size_t pitch_bytes, const uint8_t *d_top, int top_rows, size_t top_pitch_bytes,
                            const uint8_t *d_bot, int bot_rows, size_t bot_pitch_bytes, float *d_out,
                            size_t out_pitch, int sync);

/* Host-buffer batch through pinned staging with copy/compute overlap (what bench.py's e2e measures). */
int dctc_energy_batch(dctc_context *ctx, const uint8_t *imgs, int n_frames, size_t frame_stride_bytes, int w,
                      int h, int channels, size_t pitch_bytes, float *out, size_t out_frame_stride);

/* ---- K2: carver session, incremental per-seam energy ---------------------------------------------------
 * Replaces liblqr's update_emap, which re-invokes the callback for the pixels within +-radius of the removed
 * seam (radius = blocksize/2 as registered at src/render.c:314-315). */

/* Uploads the image, keeps it device-resident and builds the full map (K1). */
int dctc_carver_load(dctc_context *ctx, const uint8_t *img, int w, int h, int channels, size_t pitch_bytes);
/* Current carver width / height. */
int dctc_carver_width(const dctc_context *ctx);
int dctc_carver_height(const dctc_context *ctx);
/* Copies the current w*h energy map to the host (row pitch = current width). */
int dctc_carver_energy(dctc_context *ctx, float *out);
/* Removes one vertical seam (seam_x[y] = column removed in row y, h entries), compacts the device image and
 * energy planes and recomputes the energy only in the band the seam touched:
 *   row y: x in [min_{|y'-y|<=r} seam_x[y'] - r, max_{|y'-y|<=r} seam_x[y'] + r - 1] clipped to [0, w-2].
 * On return *xmin / *xmax (h entries each, may be NULL) hold that band and band_out (may be NULL) receives, row
 * by row, the xmax[y]-xmin[y]+1 recomputed values packed back to back (at most 4*(blocksize/2) per row: band_out needs
 * room for h*4*(blocksize/2) floats).  The seam must be connected, |seam_x[y]-seam_x[y-1]| <= 1, as liblqr's are with
 * delta_x = 1 (src/render.c:313); anything else returns DCTC_ERR_INVALID, a column outside [0, w) DCTC_ERR_STATE. */
int dctc_carve_and_update(dctc_context *ctx, const int *seam_x, float *band_out, int *xmin, int *xmax);
/* Copies the current (carved) interleaved image back to the host, pitch = w*channels. */
int dctc_carver_image(dctc_context *ctx, uint8_t *out);
/* Whole retarget loop on the device (energy -> seam DP -> back-track -> carve -> band update), `n_seams`
 * vertical seams; seams_out (may be NULL) receives n_seams*h removed columns in the coordinates of the image
 * at the time of removal.  Mirrors lqr_carver_resize(carver, w - n_seams, h) (src/render.c:377) with
 * delta_x = 1, rigidity = 0 (src/render.c:313). */
int dctc_carver_resize_width(dctc_context *ctx, int n_seams, int *seams_out);
/* Enlarging: lqr_carver_resize(carver, w + n_seams, h) as the plug-in calls it for seams_number > 0
 * (src/render.c:357-363,377).  The n_seams seams a shrink would remove are computed on a freshly loaded session (same
 * energies, same order; seams_out as above), then every pixel of those seams is doubled on the device: the new pixel
 * sits on its left with the integer mean (a + b) / 2 of the pixel and its left neighbour in the original row (a copy in
 * column 0) [liblqr lqr_carver_inflate, from memory: parity unpinned].  The session continues on the enlarged image
 * (dctc_carver_image / _energy / _width reflect it); dctc_carver_vmap returns the start frame's visibility map. */
int dctc_carver_enlarge_width(dctc_context *ctx, int n_seams, int *seams_out);
/* Optional: let the device loop update its cumulative map incrementally after each seam, like liblqr's update_mmap
 * (only the cells the removed seam can have changed are recomputed; a row whose changed range gets too wide falls
 * back to the full rebuild).  Same seams bit for bit; off by default because the full rebuild on an 8-CTA cluster is
 * currently the faster of the two. */
int dctc_carver_set_incremental(dctc_context *ctx, int on);
/* Number of full cumulative-map rebuilds the incremental mode fell back to since dctc_carver_load (-1 on error). */
int dctc_carver_rebuild_count(dctc_context *ctx);

/* ---- visibility map and seam display ------------------------------------------------------------------------
 * lqr_carver_set_dump_vmaps (src/render.c:374): ask the session to record, for every ORIGINAL pixel, the order in
 * which dctc_carver_resize_width removed it (1-based, 0 = still visible).  Call before the first seam is removed. */
int dctc_carver_set_dump_vmaps(dctc_context *ctx, int on);
/* The map (w0*h ints, row pitch = original width; lqr_vmap_get_data / _get_depth, src/render.c:216-219). */
int dctc_carver_vmap(dctc_context *ctx, int *vmap_out, int *depth_out);
/* display_carver_seams (src/render.c:204-240) on a host image of the ORIGINAL size: every removed pixel with
 * x < w-1, y < h-1 becomes (0, (guchar)(255.0 * vis / depth), 0); painted on the device. */
int dctc_carver_paint_seams(dctc_context *ctx, uint8_t *img, int channels, size_t pitch_bytes);

/* ---- K3: energy-image export ------------------------------------------------------------------------------
 * Replaces lqr_carver_get_energy_image(carver, buf, orientation, LQR_COLDEPTH_8I, LQR_GREY_IMAGE) as called at
 * src/render.c:191 (the plug-in's "output energy" option, src/render.c:175-202): e -> 1/(1+1/e) (= e/(1+e)),
 * min-max normalise in float, 8-bit grey by truncation (guchar)(val*255) [liblqr, from memory: parity unpinned]. */

/* (min, max) of e/(1+e) over a device-resident w*h float plane; lo_hi receives two floats.  For a map sharded into
 * row bands each rank calls this on its band and the pairs are all-reduced (min, max) before dctc_energy_image_dev. */
int dctc_energy_minmax_dev(dctc_context *ctx, const float *d_en, size_t en_pitch, int w, int h, float *lo_hi);
/* Writes the 8-bit grey energy image of a device-resident plane.  lo_hi = the pair to normalise with (e.g. the
 * all-reduced one), or NULL to compute it from this plane on the device. */
int dctc_energy_image_dev(dctc_context *ctx, const float *d_en, size_t en_pitch, int w, int h, const float *lo_hi,
                          uint8_t *d_out, size_t out_pitch, int sync);
/* Energy image of the carver session's current map, copied to the host (w*h bytes, pitch w). */
int dctc_carver_energy_image(dctc_context *ctx, uint8_t *out);

/* ---- preview path --------------------------------------------------------------------------------------
 * Replaces dct_energy_preview (src/render.c:421-501), the GIMP preview / "energy image" filter: the same operator
 * with the preview window (offsets -(C-1) .. b-C, C = (b-1)/2, src/render.c:43-44 and src/dct.h:8-9), BT.601 byte
 * luminance (RGB2LUMINANCE, src/render.h:5; channels 1, 3 or 4 as convert_row_to_luminance accepts), weights applied
 * on the 0..255 scale, first transform index along y (src/render.c:47-51); then normalize_image
 * (src/render.c:81-109).  energy_out (w*h floats, may be NULL) receives the un-normalised map, image_out
 * (w*h*channels bytes, may be NULL) the normalised image replicated to every channel.  Uses the context's
 * blocksize / edges / textures (any block size). */
int dctc_preview_energy(dctc_context *ctx, const uint8_t *img, int w, int h, int channels, size_t pitch_bytes,
                        float *energy_out, uint8_t *image_out);

/* ---- per-pixel symbol, kept for ABI parity ---------------------------------------------------------------
 * Same signature as the reference's LqrEnergyFunc dct_pixel_energy (src/render.c:134).  `extra_data` must
 * point to a DctcCarverEnergyParams.  A per-pixel call cannot be GPU-backed, so it is served from the host
 * mirror of the device energy map of the context's carver session (valid when (w,h) match the session's
 * current size); otherwise it returns NaN and records DCTC_ERR_STATE.  `rw` is not dereferenced. */
struct DctcLqrReadingWindow_;
float dctc_pixel_energy(int x, int y, int w, int h, struct DctcLqrReadingWindow_ *rw, void *extra_data);

/* ---- synthetic inputs (bench / tests) ----------------------------------------------------------------- */

/* Counter-based generator, identical on CPU (tests) and GPU: byte = hash32(seed, frame, y, x, c) >> 24 for
 * pattern 0 (noise); patterns 1..3 are the gradient / checkerboard / step-edge stress images of SURVEY §8d. */
int dctc_synth_fill_dev(dctc_context *ctx, uint8_t *d_img, int n_frames, size_t frame_stride_bytes, int w, int h,
                        int channels, size_t pitch_bytes, uint32_t seed, int pattern, int first_frame, int y_offset);
uint8_t dctc_synth_byte(uint32_t seed, uint32_t frame, uint32_t y, uint32_t x, uint32_t c, int pattern);

/* ---- peer access for row-band sharding, one process per GPU (BASELINE config 5) ---------------------------
 * A rank exports its band buffer with dctc_ipc_export (64-byte CUDA IPC handle), sends the handle to its
 * neighbours over any channel, and they map it with dctc_ipc_open.  The mapped pointer is then passed as
 * d_top / d_bot of dctc_energy_band_dev, so the halo rows are loaded over NVLink by the energy kernel itself. */
#define DCTC_IPC_HANDLE_BYTES 64
int dctc_ipc_export(dctc_context *ctx, void *d_ptr, unsigned char handle[DCTC_IPC_HANDLE_BYTES]);
int dctc_ipc_open(dctc_context *ctx, const unsigned char handle[DCTC_IPC_HANDLE_BYTES], void **d_peer_ptr);
int dctc_ipc_close(dctc_context *ctx, void *d_peer_ptr);

/* ---- multi-GPU: frames and row bands over the GPUs of one box ---------------------------
 * The reference runs in one process on one CPU thread (src/render.c:310-315 is the C caller of the energy path);
 * the path shards with a fixed halo and no collective: frame f -> device f mod G (BASELINE config 4), or one image
 * split into row bands, device g owning rows [h*g/G, h*(g+1)/G) (config 5).  A band needs blocksize/2-1 rows from
 * the band above and blocksize/2 rows from the band below (window offsets -b/2+1..b/2, src/render.c:146-147); the
 * image's own edges replicate (src/render.c:122-132).  Halo rows are not exchanged: the energy kernel of a band
 * loads them straight from the neighbour's HBM over NVLink (peer pointers). */

/* Band geometry of `rank` out of `world`: first row, row count, halo rows needed above / below.
 * DCTC_ERR_INVALID if some band would be thinner than blocksize/2 rows (it could not serve its neighbours' halos). */
int dctc_band_plan(int h, int world, int rank, int blocksize, int *y0, int *rows, int *top_need, int *bot_need);

/* -- one process drives all devices (what a C caller like render() would link against) -- */
typedef struct dctc_multi dctc_multi;
typedef struct dctc_multi_bands dctc_multi_bands;
/* One context (own streams) per entry of devices[] (NULL: every visible device); peer access is enabled between all
 * pairs.  The same device may be listed more than once (the bands then share one GPU; used by single-GPU tests). */
int dctc_multi_create(dctc_multi **m, const int *devices, int n_devices);
void dctc_multi_destroy(dctc_multi *m);
int dctc_multi_device_count(const dctc_multi *m);
dctc_context *dctc_multi_context(dctc_multi *m, int i);
int dctc_multi_set_params(dctc_multi *m, const DctcEnergyParameters *params);
int dctc_multi_set_kernel(dctc_multi *m, int kernel);
unsigned long long dctc_multi_launch_count(const dctc_multi *m);
/* dctc_energy_batch over all devices: frame f is processed by device f mod G, every device with its own
 * H2D / kernel / D2H pipeline, one host thread per device.  Same arguments and result as dctc_energy_batch. */
int dctc_multi_energy_batch(dctc_multi *m, const uint8_t *imgs, int n_frames, size_t frame_stride_bytes, int w, int h,
                            int channels, size_t pitch_bytes, float *out, size_t out_frame_stride);
/* One host image -> row bands over the devices -> the full energy map `out` (w*h floats, may be NULL) and / or the
 * 8-bit energy image `image_out` (w*h bytes, may be NULL; K3 with the (min, max) pair reduced over the bands).
 * Bit-identical to dctc_energy_full / dctc_carver_energy_image on one device. */
int dctc_multi_energy_bands(dctc_multi *m, const uint8_t *img, int w, int h, int channels, size_t pitch_bytes,
                            float *out, uint8_t *image_out);
/* The same in steps, bands resident on the devices (bench.py times dctc_multi_bands_energy). */
int dctc_multi_bands_create(dctc_multi *m, int w, int h, int channels, dctc_multi_bands **bands);
void dctc_multi_bands_destroy(dctc_multi_bands *bands);
int dctc_multi_bands_geometry(const dctc_multi_bands *bands, int g, int *y0, int *rows);
int dctc_multi_bands_upload(dctc_multi_bands *bands, const uint8_t *img, size_t pitch_bytes);
int dctc_multi_bands_synth(dctc_multi_bands *bands, uint32_t seed, int pattern);
int dctc_multi_bands_energy(dctc_multi_bands *bands, int sync);
int dctc_multi_bands_download(dctc_multi_bands *bands, float *out);
int dctc_multi_bands_energy_image(dctc_multi_bands *bands, uint8_t *out);

/* -- one process per GPU (bench.py under torchrun): each rank owns one band -- */
typedef struct dctc_band_runner dctc_band_runner;
/* All-gather of `bytes` bytes per rank through a POSIX shared-memory segment named after `name` (unique per
 * exchange); also a barrier.  DCTC_ERR_STATE on timeout (timeout_ms <= 0: wait forever). */
int dctc_rendezvous_allgather(const char *name, int rank, int world, const void *mine, size_t bytes, void *all,
                              int timeout_ms);
/* Allocates this rank's band of a w x h image on the context's device.  `rendezvous` names the job (e.g. the
 * launcher's port + run id); the collective calls below exchange through "<rendezvous>_<sequence number>". */
int dctc_band_runner_create(dctc_context *ctx, const char *rendezvous, int rank, int world, int w, int h, int channels,
                            dctc_band_runner **runner);
int dctc_band_runner_geometry(const dctc_band_runner *runner, int *y0, int *rows, size_t *pitch_bytes);
void *dctc_band_runner_image_dev(dctc_band_runner *runner);
float *dctc_band_runner_energy_dev(dctc_band_runner *runner);
int dctc_band_runner_synth(dctc_band_runner *runner, uint32_t seed, int pattern);
int dctc_band_runner_upload(dctc_band_runner *runner, const uint8_t *band_rows, size_t pitch_bytes);
/* COLLECTIVE: exports the band buffer as a CUDA IPC handle, all-gathers the handles and maps the neighbours' bands.
 * The band must hold its content: the exchange is the barrier between "filled" and "neighbours may read". */
int dctc_band_runner_connect(dctc_band_runner *runner);
/* COLLECTIVE: stream sync + barrier (e.g. between re-filling the bands and the next energy launch). */
int dctc_band_runner_barrier(dctc_band_runner *runner);
/* K1 on this rank's band, halo rows loaded from the neighbours' HBM by the kernel itself. */
int dctc_band_runner_energy(dctc_band_runner *runner, int sync);
int dctc_band_runner_download(dctc_band_runner *runner, float *out_rows);
/* COLLECTIVE: K3 over the sharded map ((min, max) all-gathered and reduced on the host), this rank's rows. */
int dctc_band_runner_energy_image(dctc_band_runner *runner, uint8_t *out_rows);
/* COLLECTIVE: barrier, then frees the band (no rank frees memory a neighbour's kernel may still be reading). */
void dctc_band_runner_destroy(dctc_band_runner *runner);

/* Pinned host <-> device copy bandwidth of this context's device in GB/s: H2D alone, D2H alone, and per direction
 * with both running (the ceiling of the host-buffer entry points; bench.py reports it next to `e2e`). */
int dctc_pcie_probe(dctc_context *ctx, size_t bytes, int iters, double *h2d_gbs, double *d2h_gbs,
                    double *bidir_gbs_per_dir);

/* ---- raw device memory helpers so that C / ctypes callers need no other CUDA binding -------------------- */
int dctc_dev_alloc(dctc_context *ctx, void **d_ptr, size_t bytes);
int dctc_dev_free(dctc_context *ctx, void *d_ptr);
int dctc_host_alloc_pinned(void **h_ptr, size_t bytes);
/* write-combined pinned memory for INPUT frames the host only writes (uploads do not snoop the CPU caches; CPU reads of
 * such memory are slow); freed with dctc_host_free_pinned */
int dctc_host_alloc_pinned_wc(void **h_ptr, size_t bytes);
int dctc_host_free_pinned(void *h_ptr);
int dctc_memcpy_h2d(dctc_context *ctx, void *d_dst, const void *h_src, size_t bytes);
int dctc_memcpy_d2h(dctc_context *ctx, void *h_dst, const void *d_src, size_t bytes);
int dctc_memset_dev(dctc_context *ctx, void *d_ptr, int value, size_t bytes);
int dctc_sync(dctc_context *ctx);
/* Event timing on the context stream: returns elapsed milliseconds between begin and end. */
int dctc_timer_begin(dctc_context *ctx);
int dctc_timer_end(dctc_context *ctx, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* DCTC_H */
