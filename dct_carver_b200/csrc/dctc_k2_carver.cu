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
// The map is rebuilt from scratch for every seam (same values as liblqr's incremental update_mmap).  One CTA: row y
// depends on the whole row y-1, so the rows are a chain of h barrier-separated steps; the previous row lives in
// shared memory (guarded by +inf at both ends, which reproduces the clipping), the parent offsets (-1/0/+1) go to a
// byte plane in global memory for the back-track, which warp 0 runs afterwards in batches of 32 rows: the path
// moves at most one column per row, so the 32 rows' 72-byte windows around the current column are fetched with
// independent loads (one L2 round trip per batch instead of one per row).
constexpr int DP_NT = 512;        // threads of the single DP CTA; every thread owns groups of 4 adjacent columns
constexpr int DP_MAXP = 4;        // groups per thread: widths up to 4 * 512 * 4 = 8192

// first strict minimum among (a, b, c) scanned in that order: value and parent offset -1 / 0 / +1
__device__ __forceinline__ float dp_cell(float a, float b, float c, float e, int& d)
{
    const float t = fminf(a, b);
    d = (b < a) ? 0 : -1;
    d = (c < t) ? 1 : d;
    return e + fminf(t, c);
}

// The whole kernel runs on ONE SM, so it is bound by instruction issue: four adjacent columns per thread (one 128-bit
// shared load of the previous row + two scalar halo loads, one 128-bit store, one packed 4-byte store of the parent
// offsets), energies staged through a cp.async ring a few rows ahead.  Cell x of a row lives at index x + 4; the four
// floats in front and behind hold +inf, which reproduces the range clipping at the image borders.
template <int DP_P>               // column groups per thread of this instantiation (1, 2 or 4)
__global__ void __launch_bounds__(DP_NT) dctc_seam_dp_kernel(const float* __restrict__ en, size_t en_pitch, int w, int h,
                                                             int8_t* __restrict__ dir, size_t dir_pitch,
                                                             int* __restrict__ seam, int* __restrict__ seam_log, int ring)
{
    extern __shared__ __align__(16) float dp_sm[];
    const int W4 = (w + 3) & ~3;
    const int RB = W4 + 8;
    float* prev = dp_sm;
    float* next = dp_sm + RB;
    float* enr = dp_sm + 2 * RB;                     // `ring` staged energy rows of W4 floats (ring is a power of two)
    __shared__ float red_v[DP_NT / 32];
    __shared__ int red_i[DP_NT / 32];
    __shared__ int win[32][18];
    const int tid = threadIdx.x;
    const float INF = __int_as_float(0x7f800000);
    if (tid < 4) { prev[tid] = INF; next[tid] = INF; prev[W4 + 4 + tid] = INF; next[W4 + 4 + tid] = INF; }
    // Every thread copies and later reads only its own columns: cp.async.wait_group alone orders the staged rows.
    auto stage = [&](int y) {
        if (y < h) {
            float* dst = enr + (size_t) (y & (ring - 1)) * W4;
            const float* src = en + (size_t) y * en_pitch;
#pragma unroll
            for (int p = 0; p < DP_P; p++) {
                const int x0 = 4 * (tid + p * DP_NT);
                if (x0 < W4) {
                    const uint32_t d32 = (uint32_t) __cvta_generic_to_shared(dst + x0);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d32), "l"(src + x0) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int p = 0; p < DP_P; p++) {
        const int x0 = 4 * (tid + p * DP_NT);
        if (x0 < W4) {
            float4 v = *reinterpret_cast<const float4*>(en + x0);
            if (x0 + 1 >= w) v.y = INF;
            if (x0 + 2 >= w) v.z = INF;
            if (x0 + 3 >= w) v.w = INF;
            *reinterpret_cast<float4*>(prev + x0 + 4) = v;
        }
    }
    for (int y = 1; y < ring; y++) stage(y);
    __syncthreads();
#ifdef DCTC_SYNC_DEBUG
    const long long dbg_t0 = clock64();
#endif
    // Main chain, kept to a few dozen instructions per row (the single SM is issue-bound): all addresses are running
    // pointers, the staging ring wraps by pointer comparison, only the thread that owns the row tail patches it.
    {
        const int x0 = 4 * tid;                                   // DP_P == 1 fast path uses group 0 only; others loop below
        const bool tail = x0 < W4 && x0 + 4 > w;
        const uint32_t ring_bytes = (uint32_t) ring * (uint32_t) W4 * 4u;
        const uint32_t enr_s = (uint32_t) __cvta_generic_to_shared(enr);
        uint32_t st_dst = enr_s + (uint32_t) ((ring & (ring - 1)) == 0 ? ((ring) & (ring - 1)) : 0) * 0u;   // row (ring) & (ring-1) == 0
        st_dst = enr_s + (uint32_t) x0 * 4u;                      // next row to stage is y = ring -> slot 0
        uint32_t ld_off = (uint32_t) W4 * 4u;                     // row 1 -> slot 1 (byte offset inside the ring)
        if (ring == 1) ld_off = 0;
        const float* st_src = en + (size_t) ring * en_pitch + x0; // source of the next row to stage
        int st_rows = h - ring;                                    // rows still to be staged
        int8_t* drow = dir + dir_pitch + x0;
        for (int y = 1; y < h; y++) {
#pragma unroll
            for (int p = 0; p < DP_P; p++) {
                const int xo = 4 * p * DP_NT;
                if (st_rows > 0 && x0 + xo < W4)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(st_dst + (uint32_t) xo * 4u), "l"(st_src + xo) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            st_rows--;
            st_src += en_pitch;
            st_dst += (uint32_t) W4 * 4u;
            if (st_dst >= enr_s + ring_bytes + (uint32_t) x0 * 4u) st_dst -= ring_bytes;
            switch (ring) {   // all but the newest ring-1 groups have landed: row y is in shared memory
            case 8: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
            case 4: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
            default: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
            }
            const float* erow = reinterpret_cast<const float*>(reinterpret_cast<const char*>(enr) + ld_off);
#pragma unroll
            for (int p = 0; p < DP_P; p++) {
                const int xp = x0 + 4 * p * DP_NT;
                if (xp < W4) {
                    const float l = prev[xp + 3], r = prev[xp + 8];
                    const float4 m = *reinterpret_cast<const float4*>(prev + xp + 4);
                    const float4 e = *reinterpret_cast<const float4*>(erow + xp);
                    int d0, d1, d2, d3;
                    float4 o;
                    o.x = dp_cell(l, m.x, m.y, e.x, d0);
                    o.y = dp_cell(m.x, m.y, m.z, e.y, d1);
                    o.z = dp_cell(m.y, m.z, m.w, e.z, d2);
                    o.w = dp_cell(m.z, m.w, r, e.w, d3);
                    if (xp + 4 > w) {   // tail group: cells right of the image stay +inf
                        if (xp + 1 >= w) o.y = INF;
                        if (xp + 2 >= w) o.z = INF;
                        if (xp + 3 >= w) o.w = INF;
                    }
                    *reinterpret_cast<float4*>(next + xp + 4) = o;
                    *reinterpret_cast<uint32_t*>(drow + 4 * p * DP_NT) =
                        (uint32_t) (d0 & 255) | ((uint32_t) (d1 & 255) << 8) | ((uint32_t) (d2 & 255) << 16) | ((uint32_t) (d3 & 255) << 24);
                }
            }
            (void) tail;
            drow += dir_pitch;
            ld_off += (uint32_t) W4 * 4u;
            if (ld_off >= ring_bytes) ld_off = 0;
            __syncthreads();
            float* t = prev; prev = next; next = t;
        }
    }
#ifdef DCTC_SYNC_DEBUG
    const long long dbg_t1 = clock64();
#endif
    // leftmost minimum of the last row
    const float* last = prev;
    float bv = INF;
    int bi = 0x7fffffff;
#pragma unroll
    for (int p = 0; p < DP_P; p++) {
        const int x0 = 4 * (tid + p * DP_NT);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int x = x0 + k;
            if (x < w) {
                const float v = last[x + 4];
                if (v < bv || bi == 0x7fffffff) { bv = v; bi = x; }   // ascending x per thread: strict < keeps the leftmost
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { red_v[tid >> 5] = bv; red_i[tid >> 5] = bi; }
    __threadfence_block();
    __syncthreads();
    if (tid < 32) {
        bv = tid < DP_NT / 32 ? red_v[tid] : INF;
        bi = tid < DP_NT / 32 ? red_i[tid] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi != 0x7fffffff && (bi == 0x7fffffff || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
        }
        // back-track (warp 0; every lane follows the same path, lane 0 records it)
        const int lane = tid;
        int x = bi;
        if (lane == 0) { seam[h - 1] = x; if (seam_log) seam_log[h - 1] = x; }
        const int max_base = (int) dir_pitch - 72;
        for (int ytop = h - 1; ytop >= 1; ytop -= 32) {
            int base = (x - 32) & ~3;
            base = base < 0 ? 0 : (base > max_base ? max_base : base);
            const int yy = ytop - lane;
            if (yy >= 1) {
                const int* src = reinterpret_cast<const int*>(dir + (size_t) yy * dir_pitch + base);
#pragma unroll
                for (int k = 0; k < 18; k++) win[lane][k] = __ldcg(src + k);
            }
            __syncwarp();
            for (int i = 0; i < 32 && ytop - i >= 1; i++) {
                x += (int) reinterpret_cast<const int8_t*>(win[i])[x - base];
                if (lane == 0) { seam[ytop - i - 1] = x; if (seam_log) seam_log[ytop - i - 1] = x; }
            }
            __syncwarp();
        }
#ifdef DCTC_SYNC_DEBUG
        if (tid == 0) printf("dp kernel w %d h %d: chain %lld clk, reduce+backtrack %lld clk\n", w, h, dbg_t1 - dbg_t0, clock64() - dbg_t1);
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
    // parent-offset plane: one byte per pixel, rows padded so that the 72-byte back-track windows stay inside
    const size_t dir_pitch = (((size_t) ctx->c_w0 + 3) & ~(size_t) 3) < 72 ? 72 : (((size_t) ctx->c_w0 + 3) & ~(size_t) 3);
    if (!ctx->c_dir) {
        CK(ctx, cudaMalloc((void**) &ctx->c_dir, dir_pitch * (size_t) h));
        CK(ctx, cudaMemsetAsync(ctx->c_dir, 0, dir_pitch * (size_t) h, ctx->stream));
    }
    if (ctx->c_seam_log_cap < (size_t) n_seams * h) {
        if (ctx->c_seam_log) cudaFree(ctx->c_seam_log);
        ctx->c_seam_log = nullptr; ctx->c_seam_log_cap = 0;
        CK(ctx, cudaMalloc((void**) &ctx->c_seam_log, sizeof(int) * (size_t) n_seams * h));
        ctx->c_seam_log_cap = (size_t) n_seams * h;
    }
    // staged energy rows: 8, 4 or 2 deep, whatever fits beside the two cumulative rows
    const size_t w4 = ((size_t) ctx->c_w + 3) & ~(size_t) 3;
    const size_t row_bytes = sizeof(float) * w4;
    const int ring = (8 * row_bytes <= 160 * 1024) ? 8 : (4 * row_bytes <= 160 * 1024) ? 4 : 2;
    const size_t dp_smem = sizeof(float) * 2 * (w4 + 8) + (size_t) ring * row_bytes;
    const int groups = (int) ((w4 / 4 + DP_NT - 1) / DP_NT);
    auto dp = groups <= 1 ? dctc_seam_dp_kernel<1> : groups <= 2 ? dctc_seam_dp_kernel<2> : dctc_seam_dp_kernel<4>;
    if (dp_smem > 48 * 1024) CK(ctx, cudaFuncSetAttribute(dp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) dp_smem));
    for (int s = 0; s < n_seams; s++) {
        const int w_old = ctx->c_w;
        // build_mmap + build_vpath
        dp<<<1, DP_NT, dp_smem, ctx->stream>>>(ctx->c_en, ctx->c_en_pitch, w_old, h, ctx->c_dir, dir_pitch,
                                                                ctx->c_seam, ctx->c_seam_log + (size_t) s * h, ring);
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
