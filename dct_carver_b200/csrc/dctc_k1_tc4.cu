// K1-TC4 (block size 4, full maps and row bands): the y-pass of the 4x4 block DCT on the 5th-generation tensor cores
// (tcgen05, sm_100a).  Same operator as the b = 4 instances of dctc_k1_small.cu / dctc_k1_tile.cu (reference chain
// src/render.c:134-157 -> dctNxN src/dct.c:77-94 -> ddct2d, src/fft2d/fftsg2d.c:566-627, UNNORMALISED -> src/dct.c:96-110).
//
// On the CUDA cores block size 4 is bound by the FP32 pipe (56 FP32 lane-operations per pixel, 44 of them the four
// sliding DCT-4 of the y-pass: dctc_k1_small.cu reaches 45-48 % of the HBM roofline with the pipe ~70 % busy).  The
// y-pass is the same sliding-window contraction as for block size 8 -- Tz[y'][(i,k2)] = B4[k2][y'-i], output row i of a
// step of 8 rows reads window rows i..i+3 of the 16 rows of two consecutive groups -- so it can run as
//     D_k1[128 x 32] = A_k1[128 x 16] * Tz[16 x 32]     (three kind::f16 MMAs: hi*Bh + (-lo)*Bh + hi*Bl, FP32 in TMEM).
// A first version with K1-TC8's warp-specialised pipeline (converter / producer / MMA / consumer warps, persistent) lost to
// the per-step bookkeeping of its thirteen warps (23.3 us per 4K frame).  This one has NO roles:
// a CTA of 128 threads (thread = pixel column = TMEM lane) owns 128 columns x SEG rows and goes through every group of
// 8 rows in lock step --
//     stage raw rows two groups ahead (TMA / cp.async gather) -> convert to exact integer luma (all threads) ->
//     packed FP32x2 DCT-4 along x per row pair, fp16 hi / -lo split, tcgen05.st into the operand ring (64 TMEM columns) ->
//     one thread issues the 6 MMAs of k1 = 0, 1 into the accumulator tile (64 columns), everybody loads and folds it,
//     then k1 = 2, 3 the same way -> |.|-max / class rule -> one coalesced store per row --
// and relies on FOUR co-resident CTAs per SM (4 x 128 TMEM columns) to fill each other's waits.
#include <cuda.h>       // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, no libcuda link)
#include <cuda_fp16.h>
#include <cstdio>
#include <cstring>
#include "dctc_common.cuh"
#include "dctc_launch.h"
#include "dctc_tc_tables.cuh"

namespace {

constexpr int MW = 128;            // columns per CTA = MMA M = threads
constexpr int NTHREADS = 128;
constexpr int NCONV = 128;         // every thread stages and converts
constexpr int LWP = MW + 8;        // staged luma row: index i <-> column x0-4+i (index 0 is a pad, 1..135 are read)
constexpr int NQUAD = 34;          // 4-pixel groups per staged row: columns x0-4 .. x0+131
constexpr uint32_t TMEM_COLS = 128;
constexpr uint32_t TM_A = 0;       // operand ring: (k1*2 + part)*8 + slot*4 + pair, k1 = 0..3 (64 columns)
constexpr uint32_t TM_D = 64;      // one accumulator tile of 64 columns: (k1 & 1)*32 + i*4 + k2
constexpr int PAD_SMEM = 26 * 1024;  // dynamic shared memory requested only to cap residency at 4 CTAs/SM (4 x 128 TMEM columns)

template <int CH>
struct RawGeom {
    static constexpr int CHUNKS = (16 + (MW + 4) * CH + 15) / 16;   // 16-byte chunks per staged raw row
    static constexpr int ROW = CHUNKS * 16;
};

struct alignas(128) TcSmem {
    __half B[4][32 * 16];            // Tz as UMMA K-major no-swizzle operands: [0] Bh, [1] Bl, [2]/[3] the K-swapped copies
    float2 L[2][4][LWP];             // luma of two groups: [buffer][row pair][column], .x = even row
    uint8_t Raw[3][8 * RawGeom<3>::ROW];
    uint64_t bar_d;                  // the MMAs of an accumulator tile have completed
    uint64_t bar_raw[3];             // raw buffer filled: NCONV arrivals (+ the bytes of a tensor copy)
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_inval(uint32_t bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
// Tight parity wait (labels are local to the braces).  try_wait suspends the warp in hardware for a bounded time per
// attempt; after 2^22 failed attempts (seconds) a protocol error traps, so the launch fails instead of hanging.
#ifdef DCTC_TC_DEBUG
__device__ __noinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0)
{
    for (uint32_t n = 0;; n++) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (n > (1u << 20)) { printf("mbar timeout tag %d parity %u block %d thread %d\n", tag, parity, blockIdx.x, threadIdx.x); __trap(); }
    }
}
#else
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0)
{
    (void) tag;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .u32 n;\n"
        "mov.u32 n, 0;\n"
        "DCTC_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DCTC_DONE;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DCTC_DONE;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DCTC_DONE;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DCTC_DONE;\n"
        "add.u32 n, n, 1;\n"
        "setp.lt.u32 p, n, 0x100000;\n"
        "@p bra DCTC_WAIT;\n"
        "trap;\n"
        "DCTC_DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
#endif
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, no swizzle, K-major: LBO = byte stride between core matrices along K,
// SBO = byte stride between 8-row groups along N (validated by tools/tc_probe.cu)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t) ((addr & 0x3FFFF) >> 4);
    d |= (uint64_t) ((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t) ((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t) 1 << 46;
    return d;
}
// kind::f16 instruction descriptor: D = F32, A = B = F16, both K-major, dense
__device__ __forceinline__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24); }
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

// 64 consecutive TMEM columns -> registers as two 32-column loads and one wait (a single .x64 needs 82 registers at its
// point of issue, which ptxas checks against the launch-time register target, not the setmaxnreg value of the region)
__device__ __forceinline__ void tmem_st_x4(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void tmem_ld_x64(uint32_t taddr, uint32_t (&v)[64])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
                 : "r"(taddr + 32u));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- staging + conversion (converter warps) ----------------------------------------------------------------------
// The raw interleaved bytes [x0*CH-16, x0*CH-16+ROW) of the 8 rows of a group are staged global -> shared two groups ahead
// of their conversion (x0*CH is 16-byte aligned: x0 is a multiple of 128).  Every converter thread arrives once per group
// on the raw buffer's mbarrier:
//   * group inside the image (no halo rows, no edge replication): thread 0 issues one 3-D tensor copy (box = ROW/4 x 8 x 1
//     32-bit elements at (x0*CH/4 - 4, vy0, frame); bytes left of the row start / beyond the pitch are zero-filled and
//     never read) and arrives with the expected byte count, the others just arrive;
//   * otherwise each thread gathers its 16-byte chunks with cp.async from the clamped / halo row pointers (chunks
//     outside [0, pitch) are skipped) and arrives through cp.async.mbarrier.arrive.noinc.
template <int CH>
struct StageMap {
    static constexpr int CHUNKS = RawGeom<CH>::CHUNKS;
    static constexpr int PER = (8 * CHUNKS + NCONV - 1) / NCONV;   // chunks per thread (CH=3: 3, CH=1: 1)
    __device__ __forceinline__ static void gather(const DctcK1Args& a, const uint8_t* __restrict__ img, uint8_t* __restrict__ R, int vy0, int x0, int ct)
    {
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const int c = ct + i * NCONV;
            const int r = c / CHUNKS, k = c - r * CHUNKS;
            const long long gb = (long long) x0 * CH - 16 + 16 * k;
            if (c < 8 * CHUNKS && gb >= 0 && gb + 16 <= (long long) a.pitch) {
                const uint8_t* src = dctc_row_ptr(a, img, vy0 + r) + gb;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(R + r * RawGeom<CH>::ROW + 16 * k)), "l"(src) : "memory");
            }
        }
    }
    __device__ __forceinline__ static void stage(const DctcK1Args& a, const CUtensorMap* tmap, int use_tmap, const uint8_t* __restrict__ img,
                                                 int frame, uint8_t* __restrict__ R, uint32_t bar, int vy0, int x0, int ct)
    {
        if (use_tmap && vy0 >= 0 && vy0 + 7 < a.h) {
            if (ct == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t) (8 * RawGeom<CH>::ROW)) : "memory");
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(smem_u32(R)), "l"(tmap), "r"(x0 * CH / 4 - 4), "r"(vy0), "r"(frame), "r"(bar) : "memory");
            } else {
                mbar_arrive(bar);
            }
        } else {
            gather(a, img, R, vy0, x0, ct);
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
        }
    }
};


// Luma in this kernel is the EXACT integer 2126 R + 7152 G + 722 B (= 10000 * 255 * liblqr's LQR_ER_LUMA value, below
// 2^22, so its float is exact too); grey is 10000 * v.  Two dp2a per pixel (16-bit coefficients times the pixel's bytes)
// replace three byte->float conversions and an FMA chain.  The factor 2^-13 of the scaled x-pass
// (fp16 range of the hi/lo operands) and the 1/10000 are folded into the final weight.
constexpr float LUMA_WEIGHT_SCALE = 8192.0f / 10000.0f;

template <int CH>
__device__ __forceinline__ float luma_raw(const uint8_t* __restrict__ p)
{
    if (CH == 3) return (float) (2126u * p[0] + 7152u * p[1] + 722u * p[2]);
    return (float) (10000u * p[0]);
}

// luma of four consecutive pixels from their CH*4 raw bytes (4-byte aligned); same values as luma_raw
template <int CH>
__device__ __forceinline__ void quad_luma(const uint8_t* __restrict__ p, float (&l)[4])
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
    if (CH == 3) {
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
        // dp2a.lo: R * 2126 + G * 7152 from bytes 0, 1; dp2a.hi: B * 722 (+ 0 * byte 3) from bytes 2, 3
        constexpr uint32_t CRG = (7152u << 16) | 2126u, CB = 722u;
        const uint32_t p1 = __byte_perm(w0, w1, 0x6543), p2 = __byte_perm(w1, w2, 0x5432), p3 = w2 >> 8;
        l[0] = (float) __dp2a_lo(CRG, w0, __dp2a_hi(CB, w0, 0u));
        l[1] = (float) __dp2a_lo(CRG, p1, __dp2a_hi(CB, p1, 0u));
        l[2] = (float) __dp2a_lo(CRG, p2, __dp2a_hi(CB, p2, 0u));
        l[3] = (float) __dp2a_lo(CRG, p3, __dp2a_hi(CB, p3, 0u));
    } else {
        const uint32_t w0 = w[0];
#pragma unroll
        for (int i = 0; i < 4; i++) l[i] = (float) (10000u * ((w0 >> (8 * i)) & 255u));
    }
}

// the packed DCT-4 of dctc_k1_small.cu (dct_fwd2<4>, tools/gen_dct.py: unnormalised, like ddct2d) with every constant
// scaled by 2^-13 (exact), so that the x-pass coefficients of integer luma values up to 2.55e6 stay inside the fp16 range
// of the hi/lo operand split
__device__ __forceinline__ void dct4_fwd2_scaled(const float2* __restrict__ v, float2* __restrict__ X)
{
    constexpr float S = 1.0f / 8192.0f;
    const float2 t1 = dctc_f2add(v[0], v[3]), t2 = dctc_f2sub(v[0], v[3]);
    const float2 t3 = dctc_f2add(v[1], v[2]), t4 = dctc_f2sub(v[1], v[2]);
    X[1] = dctc_f2fma(S * 3.826834261e-01f, t4, dctc_f2mul(S * 9.238795042e-01f, t2));
    X[3] = dctc_f2fma(S * -9.238795042e-01f, t4, dctc_f2mul(S * 3.826834261e-01f, t2));
    const float2 t5 = dctc_f2add(t1, t3), t6 = dctc_f2sub(t1, t3);
    X[2] = dctc_f2mul(S * 7.071067691e-01f, t6);
    X[0] = dctc_f2mul(S, t5);
}

// raw rows -> luma row pairs; staged index i <-> image column clamp(x0 - 4 + i) (src/render.c:122-132).
// A task is one 4-pixel group of one row pair (4 x 34 tasks per group of rows); a converter thread owns the same one or
// two tasks for every group of an item, so their offsets and the border test are computed once per item (ConvMap).
// Groups that touch the image border take the per-pixel clamped path.
template <int CH>
struct ConvMap {
    static constexpr int ROW = RawGeom<CH>::ROW;
    static constexpr int PER = (4 * NQUAD + NCONV - 1) / NCONV;   // 2
    int roff[PER];     // byte offset of the task's first raw row inside a raw buffer, -1: no task
    int loff[PER];     // float2 index inside a luma buffer
    int gx[PER];       // image column of the first pixel; INT_MIN when the four pixels are all inside the image
    __device__ __forceinline__ void init(const DctcK1Args& a, int x0, int ct)
    {
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const int task = ct + i * NCONV;
            const int p = task / NQUAD, q = task - p * NQUAD - 1;   // row pair 0..3, quad -1..32
            const int g = x0 + 4 * q;
            roff[i] = task < 4 * NQUAD ? (2 * p) * ROW + 16 + 4 * CH * q : -1;
            loff[i] = p * LWP + 4 * q + 4;
            gx[i] = (g >= 0 && g + 3 < a.w) ? (int) 0x80000000 : g;
        }
    }
    __device__ __forceinline__ void convert(const DctcK1Args& a, const uint8_t* __restrict__ R, float2* __restrict__ L) const
    {
#pragma unroll
        for (int i = 0; i < PER; i++) {
            if (roff[i] < 0) continue;
            const uint8_t* r0 = R + roff[i];
            const uint8_t* r1 = r0 + ROW;
            float l0[4], l1[4];
            if (gx[i] == (int) 0x80000000) {
                quad_luma<CH>(r0, l0);
                quad_luma<CH>(r1, l1);
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int off = (max(0, min(gx[i] + k, a.w - 1)) - gx[i]) * CH;
                    l0[k] = luma_raw<CH>(r0 + off);
                    l1[k] = luma_raw<CH>(r1 + off);
                }
            }
            float4* dst = reinterpret_cast<float4*>(L + loff[i]);
            dst[0] = make_float4(l0[0], l1[0], l0[1], l1[1]);
            dst[1] = make_float4(l0[2], l1[2], l0[3], l1[3]);
        }
    }
};

// H -> fp16 hi and fp16 MINUS lo for two vertically adjacent rows (low half = even row = even K index).
// hi = rn16(H); the residual comes from one mixed-precision subtract per value (sub.f32.f16 = FHADD: hi - H, exact),
// so no half->float conversion is needed; the sign is undone by the negate-A bit of the lo*Bh MMA's descriptor.
__device__ __forceinline__ void split_pair(float2 x, uint32_t& hi, uint32_t& nlo)
{
    const __half2 h = __floats2half2_rn(x.x, x.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const uint16_t h0 = (uint16_t) (hi & 0xffffu), h1 = (uint16_t) (hi >> 16);
    float r0, r1;
    asm("sub.rn.f32.f16 %0, %1, %2;" : "=f"(r0) : "h"(h0), "f"(x.x));
    asm("sub.rn.f32.f16 %0, %1, %2;" : "=f"(r1) : "h"(h1), "f"(x.y));
    const __half2 l = __floats2half2_rn(r0, r1);
    nlo = *reinterpret_cast<const uint32_t*>(&l);
}

// ---- fold ----------------------------------------------------------------------------------------------------
// accumulator column of (k1, output row i, k2) inside the tile that holds k1 (tile T = k1 >> 1)
__host__ __device__ constexpr int tc4_col(int k1, int i, int k2) { return (k1 & 1) * 32 + i * 4 + k2; }
__device__ __forceinline__ float absu(uint32_t v) { return fabsf(__uint_as_float(v)); }

template <bool UNIFORM>
struct TcFold;

template <>
struct TcFold<true> {   // edges == textures: only the maximum matters
    float m[8];
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int i = 0; i < 8; i++) m[i] = 0.0f;
    }
    template <int T>
    __device__ __forceinline__ void add(const uint32_t (&v)[64])
    {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float t = m[i];
            if (T == 0) t = fmaxf(t, absu(v[tc4_col(0, i, 1)]));      // (0,0) is skipped (src/dct.c:101)
            else t = fmaxf(t, fmaxf(absu(v[tc4_col(2, i, 0)]), absu(v[tc4_col(2, i, 1)])));
            t = fmaxf(t, fmaxf(absu(v[tc4_col(2 * T, i, 2)]), absu(v[tc4_col(2 * T, i, 3)])));
#pragma unroll
            for (int k2 = 0; k2 < 4; k2 += 2)
                t = fmaxf(t, fmaxf(absu(v[tc4_col(2 * T + 1, i, k2)]), absu(v[tc4_col(2 * T + 1, i, k2 + 1)])));
            m[i] = t;
        }
    }
    __device__ __forceinline__ float result(int i, float we, float wt) const { (void) we; return m[i] * wt; }
};

template <>
struct TcFold<false> {  // last-arg-max class rule of DctcTracker<false>
    // With A = |T[0][1]|, M = max|T[0][2..]|, Bv = |T[1][0]|, Z = max of the rest, the winner is a texture atom iff
    //   Z >= max(A, M, Bv)  or  (Bv < max(A, M) and M >= A).
    // Tile 0 holds k1 = 0 and 1, so A, M and Bv are final after it: pre = max(A, M, Bv) and the bit
    // tex_pre = (Bv < max(A, M) and M >= A) are kept next to the running Z:  texture iff  Z >= pre  or  tex_pre.
    float z[8], pre[8];
    unsigned flags;          // bit i: tex_pre of row i
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int i = 0; i < 8; i++) z[i] = 0.0f;
        flags = 0u;
    }
    template <int T>
    __device__ __forceinline__ void add(const uint32_t (&v)[64])
    {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (T == 0) {
                const float a = absu(v[tc4_col(0, i, 1)]);
                const float mm = fmaxf(absu(v[tc4_col(0, i, 2)]), absu(v[tc4_col(0, i, 3)]));
                const float am = fmaxf(a, mm);
                const float bv = absu(v[tc4_col(1, i, 0)]);
                if (mm >= a && !(bv >= am)) flags |= 1u << i;
                pre[i] = fmaxf(am, bv);
#pragma unroll
                for (int k2 = 1; k2 < 4; k2++) z[i] = fmaxf(z[i], absu(v[tc4_col(1, i, k2)]));
            } else {
#pragma unroll
                for (int k2 = 0; k2 < 4; k2++) z[i] = fmaxf(z[i], fmaxf(absu(v[tc4_col(2, i, k2)]), absu(v[tc4_col(3, i, k2)])));
            }
        }
    }
    __device__ __forceinline__ float result(int i, float we, float wt) const
    {
        const float top = fmaxf(pre[i], z[i]);
        const bool tex = (z[i] >= pre[i]) || ((flags >> i) & 1u);
        return top * (tex ? wt : we);
    }
};

// ---- kernel --------------------------------------------------------------------------------------------------
template <bool UNIFORM, int CH>
__global__ void __launch_bounds__(NTHREADS, 4) dctc_k1_tc4_kernel(const DctcK1Args a, int seg_rows, const __grid_constant__ CUtensorMap tmap,
                                                                   int use_tmap)
{
    __shared__ TcSmem s;
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler

    // programmatic dependent launch: nothing an earlier grid writes is touched before griddepcontrol.wait
    asm volatile("griddepcontrol.launch_dependents;");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // Toeplitz operands: Tz[n = i*4 + k2][k] = B4[k2][r - i] with window row r = k (normal) or k ^ 8 (K-swapped)
    {
        uint16_t* Bq = reinterpret_cast<uint16_t*>(&s.B[0][0]);
        for (int idx = tid; idx < 4 * 512; idx += NTHREADS) {
            const int v = idx >> 9, n = (idx >> 4) & 31, k = idx & 15;
            const int i = n >> 2, k2 = n & 3;
            const int c = ((v & 2) ? (k ^ 8) : k) - i;
            const uint16_t val = (c >= 0 && c < 4) ? DCTC_TC_BASIS4[v & 1][k2 * 4 + c] : (uint16_t) 0;
            Bq[v * 512 + (n >> 3) * 128 + (k >> 3) * 64 + (n & 7) * 8 + (k & 7)] = val;
        }
    }
    if (tid == 0) {
        mbar_init(smem_u32(&s.bar_d), 1);
        for (int i = 0; i < 3; i++) mbar_init(smem_u32(&s.bar_raw[i]), NCONV);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    const uint32_t tmem_lane = tmem + ((uint32_t) (warp * 32) << 16);

    const int x0 = blockIdx.x * MW;
    const int y0 = blockIdx.y * seg_rows;                       // first output row of this segment (multiple of 8)
    const int y1 = min(y0 + seg_rows, a.h);
    const int nsteps = (y1 - y0 + 7) >> 3;
    const int frame = blockIdx.z;
    const uint8_t* __restrict__ img = a.img + (size_t) frame * a.frame_stride;
    float* __restrict__ out = a.out + (size_t) frame * a.out_frame_stride;
    const int gx = x0 + tid;
    const size_t op = a.out_pitch;
    float* __restrict__ orow = out + (size_t) y0 * op + gx;
    const float we = a.w_edges * LUMA_WEIGHT_SCALE, wt = a.w_textures * LUMA_WEIGHT_SCALE;
    const uint32_t idesc = make_idesc(128, 32);
    const uint64_t bd0 = make_smem_desc(smem_u32(&s.B[0][0]), 128, 256);

    // group g = virtual rows y0-1+8g .. y0+6+8g; step j (output rows y0+8j .. +7) multiplies groups j and j+1
    ConvMap<CH> cm;
    cm.init(a, x0, tid);
    StageMap<CH>::stage(a, &tmap, use_tmap, img, frame, s.Raw[0], smem_u32(&s.bar_raw[0]), y0 - 1, x0, tid);
    StageMap<CH>::stage(a, &tmap, use_tmap, img, frame, s.Raw[1], smem_u32(&s.bar_raw[1]), y0 + 7, x0, tid);
    int slot = 0;                                               // raw buffer of group g (g % 3)
    uint32_t par = 0u;                                          // bit i: parity of the next completion of raw buffer i
    uint32_t dpar = 0u;                                         // parity of the next completion of bar_d
    for (int g = 0; g <= nsteps; g++) {
        mbar_wait(smem_u32(&s.bar_raw[slot]), (par >> slot) & 1u);   // the copies of group g have landed
        par ^= 1u << slot;
        cm.convert(a, s.Raw[slot], &s.L[g & 1][0][0]);
        __syncthreads();                                        // luma buffer g&1 is complete; every thread has left group g-1
        const int nslot = slot == 0 ? 2 : slot - 1;             // (g + 2) % 3 = (g - 1) % 3: converted before that barrier
        if (g + 2 <= nsteps)
            StageMap<CH>::stage(a, &tmap, use_tmap, img, frame, s.Raw[nslot], smem_u32(&s.bar_raw[nslot]), y0 - 1 + 8 * (g + 2), x0, tid);
        slot = slot == 2 ? 0 : slot + 1;
        // x-pass + hi / -lo split of the group; ring slot g&1 holds group g-2, whose last readers (the MMAs of step g-2)
        // completed before this thread loaded their results in the previous iteration
        {
            const float2 (*Lg)[LWP] = s.L[g & 1];
            const uint32_t ta = tmem_lane + TM_A + (uint32_t) (g & 1) * 4u;
            uint32_t hi[4][4], lo[4][4];
#pragma unroll
            for (int p = 0; p < 4; p++) {
                float2 v[4], X[4];
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = Lg[p][tid + j + 3];     // columns x0 + tid - 1 .. + 2
                dct4_fwd2_scaled(v, X);
#pragma unroll
                for (int k1 = 0; k1 < 4; k1++) split_pair(X[k1], hi[k1][p], lo[k1][p]);
            }
#pragma unroll
            for (int k1 = 0; k1 < 4; k1++) {
                tmem_st_x4(ta + (uint32_t) (k1 * 16), hi[k1][0], hi[k1][1], hi[k1][2], hi[k1][3]);
                tmem_st_x4(ta + (uint32_t) (k1 * 16 + 8), lo[k1][0], lo[k1][1], lo[k1][2], lo[k1][3]);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        __syncthreads();                                        // the operands of group g are in TMEM; the accumulator tile is free
        tc_fence_after();
        if (g == 0) continue;
        const int st = g - 1;
        // each operand copy is 1024 bytes = 64 descriptor address units; odd steps use the K-swapped copies (the older
        // group then sits in ring slot 1)
        const uint64_t bh = bd0 + (uint64_t) ((st & 1) ? 128 : 0);
        const uint64_t bl = bh + 64;
        TcFold<UNIFORM> f;
        f.init();
        auto issue_tile = [&](int t) {                         // the six MMAs of k1 = 2t, 2t+1 into the accumulator tile
            if (warp == 0) {
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 2; kk++) {
                        const uint32_t d = tmem + TM_D + 32u * kk;
                        const uint32_t ah = tmem + TM_A + (uint32_t) (2 * t + kk) * 16u, al = ah + 8u;
                        mma_ts(d, ah, bh, idesc, 0u);
                        mma_ts(d, al, bh, idesc | (1u << 13), 1u);   // A negated: the ring holds -lo
                        mma_ts(d, ah, bl, idesc, 1u);
                    }
                    mma_commit(smem_u32(&s.bar_d));
                }
                __syncwarp();
            }
        };
        uint32_t v[64];
        issue_tile(0);
        mbar_wait(smem_u32(&s.bar_d), dpar);
        dpar ^= 1u;
        tc_fence_after();
        tmem_ld_x64(tmem_lane + TM_D, v);
        tc_fence_before();
        __syncthreads();                                        // everybody has loaded tile 0: the MMAs of k1 = 2, 3 may overwrite it
        tc_fence_after();
        issue_tile(1);
        f.template add<0>(v);                                   // ... and run while tile 0 is folded
        mbar_wait(smem_u32(&s.bar_d), dpar);
        dpar ^= 1u;
        tc_fence_after();
        tmem_ld_x64(tmem_lane + TM_D, v);
        tc_fence_before();
        f.template add<1>(v);
        const int gy = y0 + 8 * st;
        if (gx < a.w) {
            float* __restrict__ o = orow;
            if (gy + 8 <= y1) {
#pragma unroll
                for (int i = 0; i < 8; i++) { *o = f.result(i, we, wt); o += op; }
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (gy + i < y1) *o = f.result(i, we, wt);
                    o += op;
                }
            }
        }
        orow += 8 * op;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

}  // namespace

// Returns cudaErrorNotSupported when the configuration is outside this kernel's fast path (the caller then uses the
// streaming kernel): needs 1 or 3 channels and 16-byte aligned row pointers / pitches.
cudaError_t dctc_launch_k1_tc4(const DctcK1Args& a, int n_frames, bool uniform, int sm_count, cudaStream_t stream)
{
    if (a.w <= 0 || a.h <= 0 || n_frames <= 0) return cudaSuccess;
    if (a.seam || a.preview) return cudaErrorNotSupported;
    auto aligned16 = [](const void* p, size_t pitch) { return (((uintptr_t) p | pitch) & 15) == 0; };
    const bool fast = (a.channels == 3 || a.channels == 1) && aligned16(a.img, a.pitch) && (a.frame_stride & 15) == 0 &&
                      (!a.top || aligned16(a.top, a.top_pitch)) && (!a.bot || aligned16(a.bot, a.bot_pitch));
    if (!fast) return cudaErrorNotSupported;
    const int strips = (a.w + MW - 1) / MW;
    // segment height: long segments amortise the per-CTA ramp (TMEM allocation, Toeplitz operands, one prologue group),
    // short ones fill the machine for small inputs
    int seg = getenv("DCTC_TC4_SEG") ? atoi(getenv("DCTC_TC4_SEG")) : 256;
    while (seg > 16 && (long long) strips * ((a.h + seg - 1) / seg) * n_frames < 8LL * sm_count) seg >>= 1;
    const int segs = (a.h + seg - 1) / seg;
    if (segs > 65535 || n_frames > 65535) return cudaErrorInvalidConfiguration;
    // tensor map of the frames as 32-bit elements: (pitch / 4, h, frames); box = one staged raw tile (ROW / 4 x 8 x 1)
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int use_tmap = 0;
    {
        typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static const encode_fn encode = []() -> encode_fn {
            void* fp = nullptr;
            cudaDriverEntryPointQueryResult qres;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) == cudaSuccess && fp &&
                qres == cudaDriverEntryPointSuccess)
                return (encode_fn) fp;
            (void) cudaGetLastError();
            return nullptr;
        }();
        const size_t fstride = n_frames > 1 ? a.frame_stride : a.pitch * (size_t) a.h;
        const int row = a.channels == 3 ? RawGeom<3>::ROW : RawGeom<1>::ROW;
        if (encode && a.h >= 8 && (fstride & 15) == 0 && fstride >= a.pitch && a.pitch < (1ull << 40) && fstride < (1ull << 40)) {
            const cuuint64_t gdim[3] = {(cuuint64_t) (a.pitch / 4), (cuuint64_t) a.h, (cuuint64_t) n_frames};
            const cuuint64_t gstr[2] = {(cuuint64_t) a.pitch, (cuuint64_t) fstride};
            const cuuint32_t box[3] = {(cuuint32_t) (row / 4), 8u, 1u};
            const cuuint32_t estr[3] = {1, 1, 1};
            if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(a.img), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                use_tmap = 1;
        }
    }
    dim3 grid(strips, segs, n_frames), block(NTHREADS);
#define DCTC_TC4_LAUNCH(U, C)                                                                                          \
    do {                                                                                                               \
        cudaError_t ea = cudaFuncSetAttribute(dctc_k1_tc4_kernel<U, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAD_SMEM); \
        if (ea != cudaSuccess) return ea;                                                                              \
        ea = dctc_launch_pdl(dctc_k1_tc4_kernel<U, C>, grid, block, PAD_SMEM, stream, true, a, seg, tmap, use_tmap);   \
        if (ea != cudaSuccess) return ea;                                                                              \
    } while (0)
    if (a.channels == 3) { if (uniform) DCTC_TC4_LAUNCH(true, 3); else DCTC_TC4_LAUNCH(false, 3); }
    else { if (uniform) DCTC_TC4_LAUNCH(true, 1); else DCTC_TC4_LAUNCH(false, 1); }
#undef DCTC_TC4_LAUNCH
    return cudaGetLastError();
}
