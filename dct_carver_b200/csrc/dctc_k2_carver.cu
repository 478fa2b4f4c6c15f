// K2: carver session with incremental per-seam energy.
//
// Replaces liblqr's per-seam "carve + update_emap" pair as driven by lqr_carver_resize (src/render.c:377):
// after a vertical seam is removed, liblqr re-invokes the energy callback (src/render.c:134-157) for the pixels
// within +-radius of the seam (radius = blocksize/2, registered at src/render.c:314-315).  Here the image and the
// energy plane stay resident in HBM; one kernel compacts both over the seam, then the K1 tile kernel runs in band
// mode over just the touched pixels, with arithmetic identical to a full recompute (bit-identical results).
#include <cstdio>
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


// ---- seam DP + back-track on the device ---------------------------------------------
// liblqr's build_mmap / build_vpath for delta_x = 1, rigidity = 0 (src/render.c:313,377; restated in
// dct_carver_b200/host/dctc_lqr.c and oracle/oracle_carver.c):
//   m[0][x] = en[0][x];  m[y][x] = en[y][x] + min(m[y-1][x-1], m[y-1][x], m[y-1][x+1])   (FP32, range clipped)
//   parent of (y, x) = FIRST strict minimum scanning x-1, x, x+1;  seam end = LEFTMOST minimum of the last row.
// The map is rebuilt from scratch for every seam (same values as liblqr's incremental update_mmap).  Row y depends on
// the whole row y-1, so the rows are a chain of h barrier-separated steps on ONE SM.  A first version that kept the
// rows in shared memory was bound by shared-memory bandwidth (measured ~600 clk per 1920-px row); here every thread
// keeps its four adjacent cells of the previous row in REGISTERS, gets the two neighbouring cells with warp
// shuffles (warp-edge lanes through a 2-float-per-warp shared exchange), prefetches the energies two rows ahead
// with 128-bit global loads and writes the cumulative rows to a global float plane.  The parent choice is not
// recorded: warp 0 re-derives it during the back-track from that plane, in batches of 32 rows -- the path moves at
// most one column per row, so the 32 rows' 80-float windows around the current column are fetched with independent
// coalesced loads (one L2 round trip per batch).  Cells right of the image hold +inf, which is the range clipping.
#ifndef DP_EXP
#define DP_EXP 0   // timing experiments only
#endif
constexpr int DP_NT = 512;        // threads of the single DP CTA; every thread owns groups of 4 adjacent columns
constexpr int DP_MAXP = 4;        // groups per thread: widths up to 4 * 512 * 4 = 8192
constexpr int DP_WIN = 80;        // back-track window (floats)

template <int DP_P>               // column groups per thread of this instantiation (1, 2 or 4)
__global__ void __launch_bounds__(DP_NT) dctc_seam_dp_kernel(const float* __restrict__ en, size_t en_pitch, int w, int h,
                                                             float* __restrict__ mplane, size_t m_pitch,
                                                             int* __restrict__ seam, int* __restrict__ seam_log)
{
    constexpr int NW = DP_NT / 32;
    __shared__ float edge_l[2][DP_P][NW], edge_r[2][DP_P][NW];   // first / last cell of every warp's span, double buffered
    __shared__ float red_v[NW];
    __shared__ int red_i[NW];
    __shared__ __align__(16) float win[32][DP_WIN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = 4 * tid;
    const int W4 = (w + 3) & ~3;
    const float INF = __int_as_float(0x7f800000);
    auto load_row = [&](int y, int p) -> float4 {
        const int xp = x0 + 4 * p * DP_NT;
        return (y < h && xp < W4) ? __ldg(reinterpret_cast<const float4*>(en + (size_t) y * en_pitch + xp)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto patch_tail = [&](float4& v, int xp) {
        if (xp >= w) v.x = INF;
        if (xp + 1 >= w) v.y = INF;
        if (xp + 2 >= w) v.z = INF;
        if (xp + 3 >= w) v.w = INF;
    };
    constexpr int PF = 8;                 // energy rows in flight per thread (registers): covers the L2 latency of a lone SM
    float4 cur[DP_P], e[PF][DP_P];
#pragma unroll
    for (int p = 0; p < DP_P; p++) {
        const int xp = x0 + 4 * p * DP_NT;
        cur[p] = load_row(0, p);
        if (xp < W4) *reinterpret_cast<float4*>(mplane + xp) = cur[p];
        patch_tail(cur[p], xp);
#pragma unroll
        for (int k = 0; k < PF; k++) e[k][p] = load_row(1 + k, p);     // e[k] holds row y with (y - 1) % PF == k
        if (lane == 0) edge_l[0][p][warp] = cur[p].x;
        if (lane == 31) edge_r[0][p][warp] = cur[p].w;
    }
    __syncthreads();
#ifdef DCTC_SYNC_DEBUG
    const long long dbg_t0 = clock64();
#endif
    float* mrow = mplane + m_pitch + x0;
    for (int yb = 1; yb < h; yb += PF) {
#pragma unroll
        for (int k = 0; k < PF; k++) {
            const int y = yb + k;
            if (y < h) {                                                    // uniform across the CTA
                const int rb = (y - 1) & 1, wb = y & 1;
#pragma unroll
                for (int p = 0; p < DP_P; p++) {
                    const int xp = x0 + 4 * p * DP_NT;
                    const float4 ee = e[k][p];
#if !(DP_EXP & 1)
                    e[k][p] = load_row(y + PF, p);                          // PF rows ahead: off the chain
#endif
                    float l = __shfl_up_sync(0xffffffffu, cur[p].w, 1);
                    float r = __shfl_down_sync(0xffffffffu, cur[p].x, 1);
                    // warp-edge lanes take their neighbour from the shared exchange; every lane issues the (broadcast)
                    // loads so that the row chain stays free of divergent branches
                    const bool has_l = warp > 0 || p > 0, has_r = warp < NW - 1 || p < DP_P - 1;
                    const float el = edge_r[rb][warp > 0 ? p : (p > 0 ? p - 1 : 0)][warp > 0 ? warp - 1 : NW - 1];
                    const float er = edge_l[rb][warp < NW - 1 ? p : (p < DP_P - 1 ? p + 1 : p)][warp < NW - 1 ? warp + 1 : 0];
                    l = lane == 0 ? (has_l ? el : INF) : l;
                    r = lane == 31 ? (has_r ? er : INF) : r;
                    float4 o;
                    o.x = ee.x + fminf(fminf(l, cur[p].x), cur[p].y);
                    o.y = ee.y + fminf(fminf(cur[p].x, cur[p].y), cur[p].z);
                    o.z = ee.z + fminf(fminf(cur[p].y, cur[p].z), cur[p].w);
                    o.w = ee.w + fminf(fminf(cur[p].z, cur[p].w), r);
#if DP_EXP & 2
                    if (o.x == 12345.678f)
#endif
                    if (xp < W4) *reinterpret_cast<float4*>(mrow + 4 * p * DP_NT) = o;
                    if (xp + 4 > w) patch_tail(o, xp);                      // cells right of the image stay +inf
                    cur[p] = o;
                    if (lane == 0) edge_l[wb][p][warp] = o.x;
                    if (lane == 31) edge_r[wb][p][warp] = o.w;
                }
                mrow += m_pitch;
#if DP_EXP & 4
                __syncwarp();
#else
                __syncthreads();
#endif
            }
        }
    }
#ifdef DCTC_SYNC_DEBUG
    const long long dbg_t1 = clock64();
#endif
    // leftmost minimum of the last row
    float bv = INF;
    int bi = 0x7fffffff;
#pragma unroll
    for (int p = 0; p < DP_P; p++) {
        const int xp = x0 + 4 * p * DP_NT;
        const float c4[4] = {cur[p].x, cur[p].y, cur[p].z, cur[p].w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int x = xp + k;
            if (x < w && (c4[k] < bv || bi == 0x7fffffff)) { bv = c4[k]; bi = x; }   // ascending x per thread: strict < keeps the leftmost
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();   // also orders this CTA's global writes of the cumulative plane before warp 0 reads them back
    if (tid < 32) {
        bv = tid < NW ? red_v[tid] : INF;
        bi = tid < NW ? red_i[tid] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi != 0x7fffffff && (bi == 0x7fffffff || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
        }
        // back-track (warp 0; every lane follows the same path, lane 0 records it)
        int x = bi;
        if (lane == 0) { seam[h - 1] = x; if (seam_log) seam_log[h - 1] = x; }
        const int max_base = (int) m_pitch - DP_WIN;
        for (int ytop = h - 1; ytop >= 1; ytop -= 32) {
            // the parents of rows ytop, ytop-1, ... are chosen in rows ytop-1, ytop-2, ...: window row i = image row ytop-1-i
            int base = (x - 36) & ~3;
            base = base < 0 ? 0 : (base > max_base ? max_base : base);
            float4 t[32];
#pragma unroll
            for (int i = 0; i < 32; i++) {
                const int yy = ytop - 1 - i;
                t[i] = (yy >= 0 && lane < DP_WIN / 4) ? __ldcg(reinterpret_cast<const float4*>(mplane + (size_t) yy * m_pitch + base) + lane)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (lane < DP_WIN / 4) {
#pragma unroll
                for (int i = 0; i < 32; i++) reinterpret_cast<float4*>(win[i])[lane] = t[i];
            }
            __syncwarp();
            const int steps = ytop < 32 ? ytop : 32;
            for (int i = 0; i < steps; i++) {
                const float* wr = win[i] - base;
                const float a = x > 0 ? wr[x - 1] : INF;
                const float b = wr[x];
                const float c = x < w - 1 ? wr[x + 1] : INF;
                int arg = x - 1;
                float best = a;
                if (b < best) { best = b; arg = x; }
                if (c < best) arg = x + 1;
                x = arg;
                if (lane == 0) { seam[ytop - 1 - i] = x; if (seam_log) seam_log[ytop - 1 - i] = x; }
            }
            __syncwarp();
        }
#ifdef DCTC_SYNC_DEBUG
        if (tid == 0) {
            unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            printf("dp kernel w %d h %d: chain %lld clk, reduce+backtrack %lld clk, globaltimer %llu ns\n", w, h, dbg_t1 - dbg_t0, clock64() - dbg_t1, gt);
        }
#endif
    }
}

void dctc_carver_release(dctc_context* ctx)
{
    if (ctx->c_img) cudaFree(ctx->c_img);
    if (ctx->c_en) cudaFree(ctx->c_en);
    if (ctx->c_m) cudaFree(ctx->c_m);
    if (ctx->c_dir) cudaFree(ctx->c_dir);
    if (ctx->c_seam_log) cudaFree(ctx->c_seam_log);
    if (ctx->c_seam) cudaFree(ctx->c_seam);
    if (ctx->c_band) cudaFree(ctx->c_band);
    if (ctx->c_band_vals) cudaFree(ctx->c_band_vals);
    if (ctx->h_mirror) cudaFreeHost(ctx->h_mirror);
    if (ctx->h_band) cudaFreeHost(ctx->h_band);
    ctx->c_img = nullptr; ctx->c_en = nullptr; ctx->c_m = nullptr; ctx->c_dir = nullptr; ctx->c_seam_log = nullptr; ctx->c_seam_log_cap = 0; ctx->c_seam = nullptr; ctx->c_band = nullptr;
    ctx->c_band_vals = nullptr; ctx->h_mirror = nullptr; ctx->h_band = nullptr;
    ctx->c_w0 = ctx->c_w = ctx->c_h = ctx->c_ch = 0;
    ctx->c_pitch = 0;
    ctx->c_en_pitch = 0;
    ctx->mirror_valid = false;
}

static void carver_args(const dctc_context* ctx, DctcK1Args& a)
{
    memset(&a, 0, sizeof(a));
    a.img = ctx->c_img; a.pitch = ctx->c_pitch; a.w = ctx->c_w; a.h = ctx->c_h; a.channels = ctx->c_ch;
    a.out = ctx->c_en; a.out_pitch = ctx->c_en_pitch;
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
    ctx->c_en_pitch = ((size_t) w + 3) & ~(size_t) 3;
    CK(ctx, cudaMalloc((void**) &ctx->c_img, ctx->c_pitch * h));
    CK(ctx, cudaMalloc((void**) &ctx->c_en, sizeof(float) * ctx->c_en_pitch * h));
    CK(ctx, cudaMemsetAsync(ctx->c_en, 0, sizeof(float) * ctx->c_en_pitch * h, ctx->stream));
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
    CK(ctx, cudaMemcpy2DAsync(out, sizeof(float) * ctx->c_w, ctx->c_en, sizeof(float) * ctx->c_en_pitch,
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
                                                               ctx->c_en_pitch, ctx->c_seam, w_old);
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
    if (!ctx || n_seams < 0) return DCTC_ERR_INVALID;
    if (!ctx->c_img) return DCTC_ERR_STATE;
    if (n_seams == 0) return DCTC_OK;
    if (n_seams >= ctx->c_w) return DCTC_ERR_STATE;
    const int h = ctx->c_h, r = ctx->blocksize / 2, bs = band_stride(ctx);
    if (ctx->c_w > 4 * DP_NT * DP_MAXP) return DCTC_ERR_UNSUPPORTED;   // wider than one CTA's columns: use the host seam loop
    CK(ctx, cudaSetDevice(ctx->device));
    // cumulative-map plane, rows padded so that the 80-float back-track windows stay inside
    const size_t m_pitch = ctx->c_en_pitch < (size_t) DP_WIN ? (size_t) DP_WIN : ctx->c_en_pitch;
    if (!ctx->c_m) {
        CK(ctx, cudaMalloc((void**) &ctx->c_m, sizeof(float) * m_pitch * (size_t) h));
        CK(ctx, cudaMemsetAsync(ctx->c_m, 0, sizeof(float) * m_pitch * (size_t) h, ctx->stream));
    }
    if (ctx->c_seam_log_cap < (size_t) n_seams * h) {
        if (ctx->c_seam_log) cudaFree(ctx->c_seam_log);
        ctx->c_seam_log = nullptr; ctx->c_seam_log_cap = 0;
        CK(ctx, cudaMalloc((void**) &ctx->c_seam_log, sizeof(int) * (size_t) n_seams * h));
        ctx->c_seam_log_cap = (size_t) n_seams * h;
    }
    const size_t w4 = ((size_t) ctx->c_w + 3) & ~(size_t) 3;
    const int groups = (int) ((w4 / 4 + DP_NT - 1) / DP_NT);
    auto dp = groups <= 1 ? dctc_seam_dp_kernel<1> : groups <= 2 ? dctc_seam_dp_kernel<2> : dctc_seam_dp_kernel<4>;
    for (int s = 0; s < n_seams; s++) {
        const int w_old = ctx->c_w;
        // build_mmap + build_vpath
        dp<<<1, DP_NT, 0, ctx->stream>>>(ctx->c_en, ctx->c_en_pitch, w_old, h, ctx->c_m, m_pitch, ctx->c_seam,
                                         ctx->c_seam_log + (size_t) s * h);
#ifdef DCTC_SYNC_DEBUG
        { cudaError_t e_ = cudaStreamSynchronize(ctx->stream); if (e_ != cudaSuccess) { printf("seam %d: dp kernel failed: %s (w %d)\n", s, cudaGetErrorString(e_), w_old); return dctc_fail_cuda(ctx, e_); } }
#endif
        // carve: compact image and energy rows over the seam
        dctc_carve_rows_kernel<256, 4><<<h, 256, 0, ctx->stream>>>(ctx->c_img, ctx->c_pitch, ctx->c_ch, ctx->c_en,
                                                                   ctx->c_en_pitch, ctx->c_seam, w_old);
        CK(ctx, cudaGetLastError());
        ctx->launches += 2;
        ctx->c_w = w_old - 1;
        // update_emap: K1 in band mode around the removed seam
        DctcK1Args a;
        carver_args(ctx, a);
        a.seam = ctx->c_seam; a.band_r = r; a.band_vals = ctx->c_band_vals; a.band_stride = bs;
#ifdef DCTC_SYNC_DEBUG
        { cudaError_t e_ = cudaStreamSynchronize(ctx->stream); if (e_ != cudaSuccess) { printf("seam %d: carve kernel failed: %s (w %d)\n", s, cudaGetErrorString(e_), w_old); return dctc_fail_cuda(ctx, e_); } }
#endif
        int rc = dctc_run_k1(ctx, a, 1, ctx->stream);
        if (rc) return rc;
#ifdef DCTC_SYNC_DEBUG
        { cudaError_t e_ = cudaStreamSynchronize(ctx->stream); if (e_ != cudaSuccess) { printf("seam %d: band kernel failed: %s (w %d)\n", s, cudaGetErrorString(e_), w_old); return dctc_fail_cuda(ctx, e_); } }
#endif
    }
    ctx->mirror_valid = false;
    if (seams_out)
        CK(ctx, cudaMemcpyAsync(seams_out, ctx->c_seam_log, sizeof(int) * (size_t) n_seams * h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return DCTC_OK;
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
