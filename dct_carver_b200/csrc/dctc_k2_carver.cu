// K2: carver session with incremental per-seam energy.
//
// Replaces liblqr's per-seam "carve + update_emap" pair as driven by lqr_carver_resize (src/render.c:377):
// after a vertical seam is removed, liblqr re-invokes the energy callback (src/render.c:134-157) for the pixels
// within +-radius of the seam (radius = blocksize/2, registered at src/render.c:314-315).  Here the image and the
// energy plane stay resident in HBM; one kernel compacts both over the seam, then the K1 tile kernel runs in band
// mode over just the touched pixels, with arithmetic identical to a full recompute (bit-identical results).
#include <cstring>
#include "dctc_common.cuh"
#include "dctc_launch.h"

#define CK(ctx, call)                                             \
    do {                                                          \
        cudaError_t e_ = (call);                                  \
        if (e_ != cudaSuccess) return dctc_fail_cuda((ctx), e_);  \
    } while (0)

int dctc_run_k1(dctc_context* ctx, DctcK1Args& a, int n_frames, cudaStream_t stream);

// One CTA per row: shift image bytes and energy floats right of the seam one pixel to the left.
// Chunks are processed left to right; the barrier between a chunk's loads and its stores also orders them
// after the previous chunk's loads, so the one-element overlap between neighbouring chunks is safe.
template <int NT, int K>
__global__ void __launch_bounds__(NT) dctc_carve_rows_kernel(uint8_t* __restrict__ img, size_t pitch, int channels,
                                                             float* __restrict__ en, size_t en_pitch,
                                                             const int* __restrict__ seam, int w_old)
{
    const int y = blockIdx.x;
    const int s = seam[y];
    uint8_t* row = img + (size_t) y * pitch;
    float* erow = en + (size_t) y * en_pitch;
    // energy floats: dst x in [s, w_old-1)
    for (int base = s; base < w_old - 1; base += NT * K) {
        float v[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int x = base + k * NT + threadIdx.x;
            v[k] = x < w_old - 1 ? erow[x + 1] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int x = base + k * NT + threadIdx.x;
            if (x < w_old - 1) erow[x] = v[k];
        }
    }
    // image bytes: dst byte in [s*ch, (w_old-1)*ch)
    const int b0 = s * channels, b1 = (w_old - 1) * channels;
    for (int base = b0; base < b1; base += NT * K) {
        uint8_t v[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int i = base + k * NT + threadIdx.x;
            v[k] = i < b1 ? row[i + channels] : (uint8_t) 0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int i = base + k * NT + threadIdx.x;
            if (i < b1) row[i] = v[k];
        }
    }
}

void dctc_carver_release(dctc_context* ctx)
{
    if (ctx->c_img) cudaFree(ctx->c_img);
    if (ctx->c_en) cudaFree(ctx->c_en);
    if (ctx->c_m) cudaFree(ctx->c_m);
    if (ctx->c_seam) cudaFree(ctx->c_seam);
    if (ctx->c_band) cudaFree(ctx->c_band);
    if (ctx->c_band_vals) cudaFree(ctx->c_band_vals);
    if (ctx->h_mirror) cudaFreeHost(ctx->h_mirror);
    if (ctx->h_band) cudaFreeHost(ctx->h_band);
    ctx->c_img = nullptr; ctx->c_en = nullptr; ctx->c_m = nullptr; ctx->c_seam = nullptr; ctx->c_band = nullptr;
    ctx->c_band_vals = nullptr; ctx->h_mirror = nullptr; ctx->h_band = nullptr;
    ctx->c_w0 = ctx->c_w = ctx->c_h = ctx->c_ch = 0;
    ctx->c_pitch = 0;
    ctx->mirror_valid = false;
}

static void carver_args(const dctc_context* ctx, DctcK1Args& a)
{
    memset(&a, 0, sizeof(a));
    a.img = ctx->c_img; a.pitch = ctx->c_pitch; a.w = ctx->c_w; a.h = ctx->c_h; a.channels = ctx->c_ch;
    a.out = ctx->c_en; a.out_pitch = (size_t) ctx->c_w0;
}

static int band_stride(const dctc_context* ctx) { return 4 * (ctx->blocksize / 2); }

extern "C" {

int dctc_carver_load(dctc_context* ctx, const uint8_t* img, int w, int h, int channels, size_t pitch)
{
    if (!ctx || !img || w <= 0 || h <= 0 || channels < 1 || channels > 4 || pitch < (size_t) w * channels)
        return DCTC_ERR_INVALID;
    const int b = ctx->blocksize;
    if (!(b == 2 || b == 4 || b == 8 || b == 16)) return DCTC_ERR_BLOCKSIZE;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    dctc_carver_release(ctx);
    ctx->c_w0 = ctx->c_w = w; ctx->c_h = h; ctx->c_ch = channels;
    ctx->c_pitch = ((size_t) w * channels + 15) & ~(size_t) 15;
    const size_t npx = (size_t) w * h;
    CK(ctx, cudaMalloc((void**) &ctx->c_img, ctx->c_pitch * h));
    CK(ctx, cudaMalloc((void**) &ctx->c_en, sizeof(float) * npx));
    CK(ctx, cudaMalloc((void**) &ctx->c_seam, sizeof(int) * h));
    CK(ctx, cudaMalloc((void**) &ctx->c_band_vals, sizeof(float) * (size_t) h * 32));
    CK(ctx, cudaMallocHost((void**) &ctx->h_mirror, sizeof(float) * npx));
    CK(ctx, cudaMallocHost((void**) &ctx->h_band, sizeof(float) * (size_t) h * 32 + sizeof(int) * h));
    CK(ctx, cudaMemcpy2DAsync(ctx->c_img, ctx->c_pitch, img, pitch, (size_t) w * channels, h, cudaMemcpyHostToDevice,
                              ctx->stream));
    DctcK1Args a;
    carver_args(ctx, a);
    // A carver session keeps one arithmetic for the whole seam loop: the per-seam band updates run in the FP32 tile
    // kernel, so the initial full map uses the bit-identical FP32 march kernel rather than the tensor-core kernel
    // (liblqr: update_emap must reproduce what build_emap would give on the carved image).
    const int saved_kernel = ctx->kernel;
    if (ctx->kernel == DCTC_KERNEL_AUTO || ctx->kernel == DCTC_KERNEL_TC_SPLIT) ctx->kernel = DCTC_KERNEL_FP32_MARCH;
    int rc = dctc_run_k1(ctx, a, 1, ctx->stream);
    ctx->kernel = saved_kernel;
    if (rc) return rc;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return DCTC_OK;
}

int dctc_carver_width(const dctc_context* ctx) { return ctx ? ctx->c_w : 0; }
int dctc_carver_height(const dctc_context* ctx) { return ctx ? ctx->c_h : 0; }

int dctc_carver_energy(dctc_context* ctx, float* out)
{
    if (!ctx || !out) return DCTC_ERR_INVALID;
    if (!ctx->c_img) return DCTC_ERR_STATE;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMemcpy2DAsync(out, sizeof(float) * ctx->c_w, ctx->c_en, sizeof(float) * ctx->c_w0,
                              sizeof(float) * ctx->c_w, ctx->c_h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return DCTC_OK;
}

int dctc_carver_image(dctc_context* ctx, uint8_t* out)
{
    if (!ctx || !out) return DCTC_ERR_INVALID;
    if (!ctx->c_img) return DCTC_ERR_STATE;
    CK(ctx, cudaSetDevice(ctx->device));
    const size_t rb = (size_t) ctx->c_w * ctx->c_ch;
    CK(ctx, cudaMemcpy2DAsync(out, rb, ctx->c_img, ctx->c_pitch, rb, ctx->c_h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return DCTC_OK;
}

int dctc_carve_and_update(dctc_context* ctx, const int* seam_x, float* band_out, int* xmin, int* xmax)
{
    if (!ctx || !seam_x) return DCTC_ERR_INVALID;
    if (!ctx->c_img || ctx->c_w <= 1) return DCTC_ERR_STATE;
    const int h = ctx->c_h, w_old = ctx->c_w, r = ctx->blocksize / 2, bs = band_stride(ctx);
    int* h_seam = (int*) ((char*) ctx->h_band + sizeof(float) * (size_t) h * 32);
    for (int y = 0; y < h; y++) {
        if (seam_x[y] < 0 || seam_x[y] >= w_old) return DCTC_ERR_STATE;
        h_seam[y] = seam_x[y];
    }
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMemcpyAsync(ctx->c_seam, h_seam, sizeof(int) * h, cudaMemcpyHostToDevice, ctx->stream));
    dctc_carve_rows_kernel<256, 4><<<h, 256, 0, ctx->stream>>>(ctx->c_img, ctx->c_pitch, ctx->c_ch, ctx->c_en,
                                                               (size_t) ctx->c_w0, ctx->c_seam, w_old);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    ctx->c_w = w_old - 1;
    ctx->mirror_valid = false;
    DctcK1Args a;
    carver_args(ctx, a);
    a.seam = ctx->c_seam; a.band_r = r; a.band_vals = ctx->c_band_vals; a.band_stride = bs;
    int rc = dctc_run_k1(ctx, a, 1, ctx->stream);
    if (rc) return rc;
    if (band_out) {
        CK(ctx, cudaMemcpyAsync(ctx->h_band, ctx->c_band_vals, sizeof(float) * (size_t) h * bs, cudaMemcpyDeviceToHost,
                                ctx->stream));
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    size_t k = 0;
    for (int y = 0; y < h; y++) {
        int lo, hi;
        dctc_band_limits(h_seam, y, h, ctx->c_w, r, &lo, &hi);
        if (xmin) xmin[y] = lo;
        if (xmax) xmax[y] = hi;
        if (band_out && hi >= lo) {
            memcpy(band_out + k, ctx->h_band + (size_t) y * bs, sizeof(float) * (size_t) (hi - lo + 1));
            k += (size_t) (hi - lo + 1);
        }
    }
    return DCTC_OK;
}

int dctc_carver_resize_width(dctc_context* ctx, int n_seams, int* seams_out)
{
    (void) n_seams; (void) seams_out;
    if (!ctx) return DCTC_ERR_INVALID;
    return DCTC_ERR_UNSUPPORTED;  // device-side seam DP: SURVEY section 8(f) rank 1, not built yet
}

float dctc_pixel_energy(int x, int y, int w, int h, struct DctcLqrReadingWindow_* rw, void* extra_data)
{
    (void) rw;
    DctcCarverEnergyParams* p = (DctcCarverEnergyParams*) extra_data;
    dctc_context* ctx = p ? p->gpu : nullptr;
    if (!ctx || !ctx->c_img || w != ctx->c_w || h != ctx->c_h || x < 0 || y < 0 || x >= w || y >= h) {
        if (ctx) ctx->last_cuda = 0;
        return __builtin_nanf("");
    }
    if (!ctx->mirror_valid) {
        if (dctc_carver_energy(ctx, ctx->h_mirror) != DCTC_OK) return __builtin_nanf("");
        ctx->mirror_valid = true;
    }
    return ctx->h_mirror[(size_t) y * w + x];
}

}  // extern "C"
