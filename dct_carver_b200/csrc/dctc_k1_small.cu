// K1 (block sizes 2 and 4, full maps): streaming register-march kernel.  These block sizes need only ~20 / ~55
// instructions per pixel, so the operator is bound by HBM traffic (7 B/px for RGB), not by the FP32 pipe
// (SURVEY 8d): the kernel is organised around keeping the image rows in flight, not around the arithmetic.
//
// Same operator as dctc_k1_tile.cu (reference chain src/render.c:134-157 -> dctNxN src/dct.c:77-94 -> ddct2d,
// src/fft2d/fftsg2d.c:566-627, unnormalised -> weighted_max_dct_correlation src/dct.c:96-110) with the same transform
// code (dctc_dct_fwd<B> along x, its packed twin along y, last-arg-max fold).  RGB luma is the EXACT integer
// 2126 R + 7152 G + 722 B (two dp2a per pixel instead of three byte->float conversions and an FMA chain; the factor
// 1/10000 folds into the final weight), so RGB maps agree with the tile kernel within the stated tolerance, grey maps
// bit for bit.  Carver sessions never mix the two (they pin the FP32 tile / march kernels).
//
//   * a CTA owns a strip of 128 columns and marches down SEG rows; thread t owns column x0+t
//   * raw interleaved rows are copied global -> shared with 16-byte cp.async, 8 rows per chunk, two chunks ahead
//   * conversion: 4 pixels per task from three 32-bit shared loads (two PRMT, one shift, eight dp2a, four I2F), the
//     tasks cover the halo columns too (34 quads per row); luma rows double buffered in shared memory
//   * per new image row ONE DCT-B along x per thread; the B coefficients of the last B rows live in a register
//     ring packed as k1 pairs (float2), the y-pass is B/2 packed DCT-B straight from registers, then the fold and
//     one coalesced float store per pixel
#include "dctc_common.cuh"
#include "dctc_launch.h"

namespace {

constexpr int MW = 128;        // columns per CTA (= threads)
constexpr int LWP = MW + 8;    // staged luma row: index i <-> column x0 - 4 + i

template <int CH, int B>
struct RawGeom {
    static constexpr int R1 = B / 2;                                    // samples after the pixel
    static constexpr int CHUNKS = (16 + (MW + 4) * CH + 15) / 16;       // 16-byte chunks per staged raw row (quads up to x0+131)
    static constexpr int ROW = CHUNKS * 16;
    static constexpr int PER = (8 * CHUNKS + MW - 1) / MW;              // chunks per thread and row block
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

// packed twins of dctc_dct_fwd<2> / <4> (tools/gen_dct.py): the same operations in the same order on two lanes
template <int N>
__device__ __forceinline__ void dct_fwd2(const float2* __restrict__ v, float2* __restrict__ X);

template <>
__device__ __forceinline__ void dct_fwd2<2>(const float2* __restrict__ v, float2* __restrict__ X)
{
    const float2 t1 = dctc_f2add(v[0], v[1]), t2 = dctc_f2sub(v[0], v[1]);
    X[1] = dctc_f2mul(7.071067691e-01f, t2);
    X[0] = t1;
}

template <>
__device__ __forceinline__ void dct_fwd2<4>(const float2* __restrict__ v, float2* __restrict__ X)
{
    const float2 t1 = dctc_f2add(v[0], v[3]), t2 = dctc_f2sub(v[0], v[3]);
    const float2 t3 = dctc_f2add(v[1], v[2]), t4 = dctc_f2sub(v[1], v[2]);
    X[1] = dctc_f2fma(3.826834261e-01f, t4, dctc_f2mul(9.238795042e-01f, t2));
    X[3] = dctc_f2fma(-9.238795042e-01f, t4, dctc_f2mul(3.826834261e-01f, t2));
    const float2 t5 = dctc_f2add(t1, t3), t6 = dctc_f2sub(t1, t3);
    X[2] = dctc_f2mul(7.071067691e-01f, t6);
    X[0] = t5;
}

// fold of one k1 pair: X[k2] = (T[2p][k2], T[2p+1][k2]); same rule as DctcTracker (dctc_common.cuh)
template <int B, bool UNIFORM>
struct Fold;

template <int B>
struct Fold<B, true> {
    float m;
    __device__ __forceinline__ void init() { m = 0.0f; }
    template <int PAIR>
    __device__ __forceinline__ void add(const float2* X)
    {
#pragma unroll
        for (int k2 = 0; k2 < B; k2++) {
            if (PAIR == 0 && k2 == 0) m = fmaxf(m, fabsf(X[0].y));   // (0,0) is skipped (src/dct.c:101)
            else m = fmaxf(m, fmaxf(fabsf(X[k2].x), fabsf(X[k2].y)));
        }
    }
    __device__ __forceinline__ float result(float we, float wt) const { (void) we; return m * wt; }
};

template <int B>
struct Fold<B, false> {
    float a, mm, bv, z;
    __device__ __forceinline__ void init() { a = 0.0f; mm = -1.0f; bv = 0.0f; z = 0.0f; }
    template <int PAIR>
    __device__ __forceinline__ void add(const float2* X)
    {
        if (PAIR == 0) {   // .x is k1 = 0, .y is k1 = 1
            a = fabsf(X[1].x);
            bv = fabsf(X[0].y);
#pragma unroll
            for (int k2 = 2; k2 < B; k2++) mm = fmaxf(mm, fabsf(X[k2].x));
#pragma unroll
            for (int k2 = 1; k2 < B; k2++) z = fmaxf(z, fabsf(X[k2].y));
        } else {
#pragma unroll
            for (int k2 = 0; k2 < B; k2++) z = fmaxf(z, fmaxf(fabsf(X[k2].x), fabsf(X[k2].y)));
        }
    }
    __device__ __forceinline__ float result(float we, float wt) const
    {
        const float am = fmaxf(a, mm);
        const float top = fmaxf(fmaxf(am, bv), z);
        const bool tex = (z >= fmaxf(am, bv)) || (!(bv >= am) && (mm >= a));
        return top * (tex ? wt : we);
    }
};

// ---- staging ---------------------------------------------------------------------------------------------------
// raw bytes [x0*CH - 16, x0*CH + (MW + R1)*CH) of 8 rows; chunks outside [0, pitch) are skipped (clamped pixel indices
// never read them)
template <int CH, int B, int NROWS = 8>
__device__ __forceinline__ void stage_raw_async(const DctcK1Args& a, const uint8_t* __restrict__ img, uint8_t* __restrict__ R,
                                                int vy0, int x0, int tid)
{
    using G = RawGeom<CH, B>;
    constexpr int PER = (NROWS * G::CHUNKS + MW - 1) / MW;
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int c = tid + i * MW;
        const int ly = c / G::CHUNKS, k = c - ly * G::CHUNKS;
        const long long gb = (long long) x0 * CH - 16 + 16 * k;
        if (c < NROWS * G::CHUNKS && gb >= 0 && gb + 16 <= (long long) a.pitch) {
            const uint8_t* src = dctc_row_ptr(a, img, vy0 + ly) + gb;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(R + ly * G::ROW + 16 * k)), "l"(src) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// The chunk -> (row, byte offset) mapping of a thread is the same for every 8-row chunk of a segment: it is computed
// once, and while consecutive chunks lie inside the band itself (no halo rows, no edge replication) the source pointers
// just advance by 8 pitches instead of being looked up row by row.
template <int CH, int B>
struct StageMap {
    using G = RawGeom<CH, B>;
    int ly[G::PER];              // row of the chunk, -1: no copy
    int soff[G::PER];            // byte offset inside a raw buffer
    const uint8_t* src[G::PER];
    int last_vy0;
    __device__ __forceinline__ void init(const DctcK1Args& a, int x0, int tid)
    {
#pragma unroll
        for (int i = 0; i < G::PER; i++) {
            const int c = tid + i * MW;
            const int r = c / G::CHUNKS, k = c - r * G::CHUNKS;
            const long long gb = (long long) x0 * CH - 16 + 16 * k;
            soff[i] = r * G::ROW + 16 * k;
            ly[i] = (c < 8 * G::CHUNKS && gb >= 0 && gb + 16 <= (long long) a.pitch) ? r : -1;
            src[i] = nullptr;
        }
        last_vy0 = (int) 0x80000000;
    }
    __device__ __forceinline__ void stage(const DctcK1Args& a, const uint8_t* __restrict__ img, uint8_t* __restrict__ R, int vy0, int x0)
    {
        const bool step8 = vy0 == last_vy0 + 8 && last_vy0 >= 0 && vy0 + 7 < a.h;
#pragma unroll
        for (int i = 0; i < G::PER; i++) {
            if (ly[i] >= 0) {
                if (step8) src[i] += 8 * a.pitch;
                else src[i] = dctc_row_ptr(a, img, vy0 + ly[i]) + ((long long) x0 * CH - 16 + (soff[i] - ly[i] * G::ROW));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(R + soff[i])), "l"(src[i]) : "memory");
            }
        }
        last_vy0 = vy0;
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
};

// RGB luma as the exact integer 2126 R + 7152 G + 722 B (< 2^22: exact as a float); grey stays the byte itself
constexpr float LUMA_INT_SCALE = 1.0f / 10000.0f;
constexpr uint32_t LUMA_RG = (7152u << 16) | 2126u;    // dp2a.lo: R * 2126 + G * 7152 from bytes 0, 1
constexpr uint32_t LUMA_B = 722u;                       // dp2a.hi: B * 722 (+ 0 * byte 3) from bytes 2, 3

template <int CH>
__device__ __forceinline__ float luma_raw(const uint8_t* __restrict__ p)
{
    if (CH == 3) return (float) (2126u * p[0] + 7152u * p[1] + 722u * p[2]);
    return (float) p[0];
}

// byte k (0..3) of w -> float, exactly: the byte becomes the low mantissa bits of 2^23 + byte
__device__ __forceinline__ float byte_to_float(uint32_t w, int k)
{
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | (uint32_t) k)) - 8388608.0f;
}

// luma of four consecutive pixels from their CH*4 raw bytes (4-byte aligned); same values as luma_raw
template <int CH>
__device__ __forceinline__ float4 quad_luma(const uint8_t* __restrict__ p)
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
    float4 l;
    if (CH == 3) {
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
        const uint32_t p1 = __byte_perm(w0, w1, 0x6543), p2 = __byte_perm(w1, w2, 0x5432), p3 = w2 >> 8;
        l.x = (float) __dp2a_lo(LUMA_RG, w0, __dp2a_hi(LUMA_B, w0, 0u));
        l.y = (float) __dp2a_lo(LUMA_RG, p1, __dp2a_hi(LUMA_B, p1, 0u));
        l.z = (float) __dp2a_lo(LUMA_RG, p2, __dp2a_hi(LUMA_B, p2, 0u));
        l.w = (float) __dp2a_lo(LUMA_RG, p3, __dp2a_hi(LUMA_B, p3, 0u));
    } else {
        const uint32_t w0 = w[0];
        l.x = byte_to_float(w0, 0);
        l.y = byte_to_float(w0, 1);
        l.z = byte_to_float(w0, 2);
        l.w = byte_to_float(w0, 3);
    }
    return l;
}

// NROWS raw rows -> luma rows; staged index i <-> image column clamp(x0 - 4 + i) (src/render.c:122-132).
// A task is one 4-pixel group of one row: 34 quads per row cover the columns x0-4 .. x0+131, i.e. the strip and its
// halo columns; a thread owns the same tasks for every chunk of a segment, so their offsets and the border test are
// computed once (ConvMap).  Quads that touch the image border take the per-pixel clamped path.
constexpr int NQUAD = LWP / 4;
template <int CH, int B>
struct ConvMap {
    using G = RawGeom<CH, B>;
    static constexpr int PER = (8 * NQUAD + MW - 1) / MW;   // 3
    int roff[PER];     // byte offset of the task's raw quad inside a raw buffer, -1: no task; bit 30 set: the quad
                       // touches the image border (per-pixel clamped path)
    int loff[PER];     // float index inside a luma buffer
    static constexpr int BORDER = 1 << 30;
    __device__ __forceinline__ void init(const DctcK1Args& a, int x0, int tid)
    {
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const int task = tid + i * MW;
            const int r = task / NQUAD, q = task - r * NQUAD - 1;   // row 0..7, quad -1..32
            const int g = x0 + 4 * q;
            roff[i] = task < 8 * NQUAD ? (r * G::ROW + 16 + 4 * CH * q) | ((g >= 0 && g + 3 < a.w) ? 0 : BORDER) : -1;
            loff[i] = r * LWP + 4 * q + 4;
        }
    }
    template <int NROWS>
    __device__ __forceinline__ void convert(const DctcK1Args& a, const uint8_t* __restrict__ R, float* __restrict__ L, int x0) const
    {
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const int ro = roff[i] & ~BORDER;
            if (roff[i] < 0 || (NROWS < 8 && ro >= NROWS * G::ROW)) continue;
            const uint8_t* r = R + ro;
            float4 l;
            if (!(roff[i] & BORDER)) {
                l = quad_luma<CH>(r);
            } else {
                const int g = x0 + (loff[i] % LWP) - 4;              // image column of the quad's first pixel
                float t[4];
#pragma unroll
                for (int k = 0; k < 4; k++) t[k] = luma_raw<CH>(r + (max(0, min(g + k, a.w - 1)) - g) * CH);
                l = make_float4(t[0], t[1], t[2], t[3]);
            }
            *reinterpret_cast<float4*>(L + loff[i]) = l;
        }
    }
};

// ---- march -----------------------------------------------------------------------------------------------------
template <int B, int SLOT>
__device__ __forceinline__ void xpass(float2 (&H2)[B][B / 2], const float* __restrict__ Lrow, int tid)
{
    constexpr int R0 = B / 2 - 1;
    float v[B], X[B];
#pragma unroll
    for (int j = 0; j < B; j++) v[j] = Lrow[tid + 4 - R0 + j];
    dctc_dct_fwd<B>(v, X);
#pragma unroll
    for (int p = 0; p < B / 2; p++) H2[SLOT][p] = make_float2(X[2 * p], X[2 * p + 1]);
}

// y-pass over the window whose oldest row sits in ring slot J
template <int B, int J, bool UNIFORM>
__device__ __forceinline__ float ypass(const float2 (&H2)[B][B / 2], float we, float wt)
{
    Fold<B, UNIFORM> f;
    f.init();
    {
        float2 v[B], X[B];
#pragma unroll
        for (int i = 0; i < B; i++) v[i] = H2[(J + i) % B][0];
        dct_fwd2<B>(v, X);
        f.template add<0>(X);
    }
#pragma unroll
    for (int p = 1; p < B / 2; p++) {
        float2 v[B], X[B];
#pragma unroll
        for (int i = 0; i < B; i++) v[i] = H2[(J + i) % B][p];
        dct_fwd2<B>(v, X);
        f.template add<1>(X);
    }
    return f.result(we, wt);
}

// row RW of the chunk: the new image row lands in ring slot (RW + B - 1) % B, the window of output row gy+RW starts in
// slot RW % B; `o` points at this thread's pixel of that output row
template <int B, int RW, bool UNIFORM>
__device__ __forceinline__ void step(float2 (&H2)[B][B / 2], const float* __restrict__ Lbuf, int tid, float we, float wt,
                                     float* __restrict__ o, bool ok)
{
    xpass<B, (RW + B - 1) % B>(H2, Lbuf + RW * LWP, tid);
    const float e = ypass<B, RW % B, UNIFORM>(H2, we, wt);
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %2, 0;\n@p st.global.f32 [%0], %1;\n}" ::"l"(o), "f"(e), "r"((int) ok) : "memory");
}

template <int B, bool UNIFORM, int CH>
__global__ void __launch_bounds__(MW, B == 2 ? 8 : 6) dctc_k1_small_kernel(const DctcK1Args a, int seg_rows)
{
    using G = RawGeom<CH, B>;
    constexpr int R0 = B / 2 - 1, R1 = B / 2;
    __shared__ __align__(16) float L[2][8 * LWP];
    __shared__ __align__(16) uint8_t Raw[3][8 * G::ROW];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * MW;
    const int y0 = blockIdx.y * seg_rows;                       // first output row of this segment (multiple of 8)
    const int y1 = min(y0 + seg_rows, a.h);
    const uint8_t* __restrict__ img = a.img + (size_t) blockIdx.z * a.frame_stride;
    float* __restrict__ out = a.out + (size_t) blockIdx.z * a.out_frame_stride;
    const int gx = x0 + tid;
    const bool colok = gx < a.w;
    float2 H2[B][B / 2];
    const float lscale = CH == 3 ? LUMA_INT_SCALE : 1.0f;       // RGB luma is 10000 x the 0..255 luma
    const float we = a.w_edges * lscale, wt = a.w_textures * lscale;
    ConvMap<CH, B> cm;
    cm.init(a, x0, tid);

    // chunk 0 = virtual rows y0-R0 .. (only the first B-1 are used: prologue); chunk c >= 1 = rows y0+R1+8(c-1) .. +7,
    // which feed output rows y0+8(c-1) .. +7.  Raw buffers rotate over three slots (two chunks in flight).
    const int nchunks = 1 + (y1 - y0 + 7) / 8;
    stage_raw_async<CH, B, B - 1>(a, img, Raw[0], y0 - R0, x0, tid);    // the prologue needs B-1 rows only
    StageMap<CH, B> sm;
    sm.init(a, x0, tid);
    sm.stage(a, img, Raw[1], y0 + R1, x0);
    if (nchunks > 2) sm.stage(a, img, Raw[2], y0 + R1 + 8, x0);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    __syncthreads();
    cm.template convert<B - 1>(a, Raw[0], L[0], x0);
    __syncthreads();
    if (B == 4) {
        xpass<B, 0>(H2, L[0] + 0 * LWP, tid);
        xpass<B, 1 % B>(H2, L[0] + 1 * LWP, tid);
        xpass<B, 2 % B>(H2, L[0] + 2 * LWP, tid);
    } else {
        xpass<B, 0>(H2, L[0] + 0 * LWP, tid);
    }
    int slot = 1;                                               // raw buffer of the chunk being converted next
    for (int c = 1; c < nchunks; c++) {
        const int gy = y0 + 8 * (c - 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");   // this thread's copies of chunk c have landed
        __syncthreads();                                        // ... everybody's; L[c&1] and the raw buffer of chunk c-1 are free
        // refill the raw buffer chunk c-1 sat in with chunk c+2
        const int fslot = slot == 0 ? 2 : slot - 1;
        if (c + 2 < nchunks) sm.stage(a, img, Raw[fslot], y0 + R1 + 8 * (c + 1), x0);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        float* Lb = L[c & 1];
        cm.template convert<8>(a, Raw[slot], Lb, x0);
        __syncthreads();
        // one pointer per chunk, advanced by the pitch: no 64-bit multiply and no divergent guard per pixel
        float* o = out + (size_t) gy * a.out_pitch + gx;
        const size_t op = a.out_pitch;
        if (gy + 8 <= a.h) {
            step<B, 0, UNIFORM>(H2, Lb, tid, we, wt, o, colok); o += op;
            step<B, 1, UNIFORM>(H2, Lb, tid, we, wt, o, colok); o += op;
            step<B, 2, UNIFORM>(H2, Lb, tid, we, wt, o, colok); o += op;
            step<B, 3, UNIFORM>(H2, Lb, tid, we, wt, o, colok); o += op;
            step<B, 4, UNIFORM>(H2, Lb, tid, we, wt, o, colok); o += op;
            step<B, 5, UNIFORM>(H2, Lb, tid, we, wt, o, colok); o += op;
            step<B, 6, UNIFORM>(H2, Lb, tid, we, wt, o, colok); o += op;
            step<B, 7, UNIFORM>(H2, Lb, tid, we, wt, o, colok);
        } else {
            step<B, 0, UNIFORM>(H2, Lb, tid, we, wt, o, colok && gy + 0 < a.h); o += op;
            step<B, 1, UNIFORM>(H2, Lb, tid, we, wt, o, colok && gy + 1 < a.h); o += op;
            step<B, 2, UNIFORM>(H2, Lb, tid, we, wt, o, colok && gy + 2 < a.h); o += op;
            step<B, 3, UNIFORM>(H2, Lb, tid, we, wt, o, colok && gy + 3 < a.h); o += op;
            step<B, 4, UNIFORM>(H2, Lb, tid, we, wt, o, colok && gy + 4 < a.h); o += op;
            step<B, 5, UNIFORM>(H2, Lb, tid, we, wt, o, colok && gy + 5 < a.h); o += op;
            step<B, 6, UNIFORM>(H2, Lb, tid, we, wt, o, colok && gy + 6 < a.h); o += op;
            step<B, 7, UNIFORM>(H2, Lb, tid, we, wt, o, colok && gy + 7 < a.h);
        }
        slot = slot == 2 ? 0 : slot + 1;
    }
}

template <int B>
cudaError_t launch_small(const DctcK1Args& a, int n_frames, bool uniform, int sm_count, cudaStream_t stream)
{
    const int strips = (a.w + MW - 1) / MW;
    // segment height: long segments amortise the prologue, short ones fill the machine for small inputs
    // (measured on 16 frames of 4K, b=2: 256-row segments 11.1 us per frame, 128 rows 10.7, 64 rows 11.0, 32 rows 11.7)
    int seg = 256;
    while (seg > 32 && (long long) strips * ((a.h + seg - 1) / seg) * n_frames < 32LL * sm_count) seg >>= 1;
    const int segs = (a.h + seg - 1) / seg;
    if (segs > 65535 || n_frames > 65535) return cudaErrorInvalidConfiguration;
    dim3 grid(strips, segs, n_frames), block(MW);
#define DCTC_SMALL_LAUNCH(U, C) dctc_k1_small_kernel<B, U, C><<<grid, block, 0, stream>>>(a, seg)
    if (a.channels == 3) { if (uniform) DCTC_SMALL_LAUNCH(true, 3); else DCTC_SMALL_LAUNCH(false, 3); }
    else { if (uniform) DCTC_SMALL_LAUNCH(true, 1); else DCTC_SMALL_LAUNCH(false, 1); }
#undef DCTC_SMALL_LAUNCH
    return cudaGetLastError();
}

}  // namespace

// Returns cudaErrorNotSupported when the configuration is outside this kernel's fast path (the caller then uses the
// tile kernel): needs block size 2 or 4, 1 or 3 channels, 16-byte aligned row pointers / pitches, no band / preview mode.
cudaError_t dctc_launch_k1_small(const DctcK1Args& a, int blocksize, int n_frames, bool uniform, int sm_count, cudaStream_t stream)
{
    if (a.w <= 0 || a.h <= 0 || n_frames <= 0) return cudaSuccess;
    if (a.seam || a.preview || (blocksize != 2 && blocksize != 4)) return cudaErrorNotSupported;
    auto aligned16 = [](const void* p, size_t pitch) { return (((uintptr_t) p | pitch) & 15) == 0; };
    const bool fast = (a.channels == 3 || a.channels == 1) && aligned16(a.img, a.pitch) && (a.frame_stride & 15) == 0 &&
                      (!a.top || aligned16(a.top, a.top_pitch)) && (!a.bot || aligned16(a.bot, a.bot_pitch));
    if (!fast) return cudaErrorNotSupported;
    return blocksize == 2 ? launch_small<2>(a, n_frames, uniform, sm_count, stream) : launch_small<4>(a, n_frames, uniform, sm_count, stream);
}
