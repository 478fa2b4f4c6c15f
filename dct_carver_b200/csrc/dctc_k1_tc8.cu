// K1-TC (block size 8, full maps): the y-pass of the block DCT on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Same operator as dctc_k1_march8.cu / dctc_k1_tile.cu (reference chain src/render.c:134-157 -> src/dct.c:77-110 ->
// ddct8x8s, src/fft2d/shrtdct.c:61-117).  A CTA owns a strip of 128 columns and marches down its segment 8 rows
// ("a group") at a time; pixel column x0+m is row m of every MMA (TMEM lane m).
//
//   producer warps 0-3 (thread = column): cp.async the raw rows of the next group, convert them to luma, run the
//       x-pass (one packed FP32x2 DCT-8 per row pair), split each coefficient H[k1] into fp16 hi + fp16 lo
//       (hi = rn16(H), lo = rn16(H - hi): 22 significant bits) and store the group into the TMEM A operand ring
//       with tcgen05.st: columns [k1][hi|lo][slot g&1][row pair], two consecutive rows per 32-bit column
//   MMA warp 8 (one elected thread): for every k1, D[128 x 64] = A_k1[128 x 16] * Tz[16 x 64] with
//       Tz[y'][(i,k2)] = B8[k2][y'-i] (Toeplitz expansion of the DCT basis: output row i of the group reads window
//       rows i..i+7 of the 16 staged rows), as three kind::f16 MMAs with FP32 accumulation in TMEM:
//       hi*Bh + lo*Bh + hi*Bl (the dropped lo*Bl term is 2^-22 relative).  Odd steps use the K-swapped copy of Tz
//       because the older group then sits in ring slot 1.
//   consumer warps 4-7 (thread = column): tcgen05.ld the 64 accumulators of (8 rows x 8 k2), fold |.|-max over k2
//       and k1 with FMNMX3 (or the last-arg-max class tracker when edges != textures), scale, coalesced store.
//
// Per pixel the CUDA cores execute ~100 instructions instead of the ~264 of the FP32 march kernel; the tensor pipe
// does 24 M128 N64 K16 MMAs per 1024 pixels (36.3 clk each measured, profiles/r01_tcgen05_probe2.txt).
#include <cuda_fp16.h>
#include <cstdio>
#include "dctc_common.cuh"
#include "dctc_launch.h"
#include "dctc_tc_tables.cuh"

namespace {

constexpr int MW = 128;            // columns per CTA = MMA M
constexpr int LWP = MW + 8;        // staged luma row: columns x0-3 .. x0+MW+3 (+1 pad)
constexpr int NTHREADS = 288;      // 4 producer warps, 4 consumer warps, 1 MMA warp
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t TM_A = 0;       // A ring: (k1*2 + part)*8 + slot*4 + pair
constexpr uint32_t TM_D = 128;     // two accumulator tiles of 64 columns
constexpr int PAD_SMEM = 64 * 1024;  // dynamic shared memory requested only to cap residency at 2 CTAs/SM (2 x 256 TMEM columns)

template <int CH>
struct RawGeom {
    static constexpr int CHUNKS = (16 + (MW + 4) * CH + 15) / 16;   // 16-byte chunks per staged raw row
    static constexpr int ROW = CHUNKS * 16;
};

struct alignas(128) TcSmem {
    __half B[4][64 * 16];            // Tz as UMMA K-major no-swizzle operands: [0] Bh, [1] Bl, [2]/[3] the K-swapped copies
    float2 L[4][LWP];                // luma of the current group: [row pair][column], .x = even row
    uint8_t Raw[2][8 * RawGeom<3>::ROW];
    uint64_t bar_a_full, bar_a_free, bar_d_full[2], bar_d_free[2];
    uint32_t tmem_base;
    int work;
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
#if DCTC_TC_WAITMODE == 1
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
#endif
        "@p bra DCTC_DONE;\n"
        "add.u32 n, n, 1;\n"
        "setp.lt.u32 p, n, 0x400000;\n"
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
__device__ __forceinline__ void bar_producers() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// Accumulator tile t is handed back to the MMA warp through hardware named barrier 2+t (4 consumer warps arrive, the
// MMA warp syncs: 160 threads): the wake-up is immediate, unlike polling an mbarrier from the issuing thread.
__device__ __forceinline__ void bar_tile_arrive(int t) { asm volatile("bar.arrive %0, 160;" ::"r"(2 + t) : "memory"); }
__device__ __forceinline__ void bar_tile_sync(int t) { asm volatile("bar.sync %0, 160;" ::"r"(2 + t) : "memory"); }

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

__device__ __forceinline__ void tmem_st_x2(uint32_t taddr, uint32_t r0, uint32_t r1)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(r0), "r"(r1) : "memory");
}
__device__ __forceinline__ void tmem_ld_x64(uint32_t taddr, uint32_t (&v)[64])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- staging (producer warps; same scheme as dctc_k1_march8.cu) ------------------------------------------------
// The raw interleaved bytes [x0*CH-16, x0*CH+(MW+4)*CH) of the 8 rows of a group are copied global -> shared with
// 16-byte cp.async one group ahead of their conversion (x0*CH is 16-byte aligned: x0 is a multiple of 128).
// Chunks outside [0, pitch) are skipped: clamped pixel indices never read them.
template <int CH>
__device__ __forceinline__ void stage_raw_async(const DctcK1Args& a, const uint8_t* __restrict__ img, uint8_t* __restrict__ R,
                                                int vy0, int x0, int tid)
{
    const int warp = tid >> 5, lane = tid & 31;
    if (lane < RawGeom<CH>::CHUNKS) {
        const long long gb = (long long) x0 * CH - 16 + 16 * lane;
        if (gb >= 0 && gb + 16 <= (long long) a.pitch) {
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int ly = 2 * warp + r;
                const uint8_t* src = dctc_row_ptr(a, img, vy0 + ly) + gb;
                const uint32_t dst = smem_u32(R + ly * RawGeom<CH>::ROW + 16 * lane);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int CH>
__device__ __forceinline__ float luma_raw(const uint8_t* __restrict__ p)
{
    if (CH == 3) return fmaf(0.2126f, (float) p[0], fmaf(0.7152f, (float) p[1], 0.0722f * (float) p[2]));
    return (float) p[0];
}

// raw rows -> luma row pairs; staged column lx <-> image column clamp(x0 + lx - 3) (src/render.c:122-132)
template <int CH>
__device__ __forceinline__ void convert_raw(const DctcK1Args& a, const uint8_t* __restrict__ R, float2 (*L)[LWP], int x0, int tid)
{
    constexpr int ROW = RawGeom<CH>::ROW;
    const int b0 = (max(0, min(x0 + tid - 3, a.w - 1)) - x0) * CH + 16;
#pragma unroll
    for (int p = 0; p < 4; p++)
        L[p][tid] = make_float2(luma_raw<CH>(R + (2 * p) * ROW + b0), luma_raw<CH>(R + (2 * p + 1) * ROW + b0));
    if (tid < 28) {
        const int p = tid / 7, lx = MW + tid - p * 7;
        const int b1 = (max(0, min(x0 + lx - 3, a.w - 1)) - x0) * CH + 16;
        L[p][lx] = make_float2(luma_raw<CH>(R + (2 * p) * ROW + b1), luma_raw<CH>(R + (2 * p + 1) * ROW + b1));
    }
}

// H -> fp16 hi + fp16 lo for two vertically adjacent rows (low half = even row = even K index)
__device__ __forceinline__ void split_pair(float2 x, uint32_t& hi, uint32_t& lo)
{
    const __half2 h = __floats2half2_rn(x.x, x.y);
    const float2 r = dctc_f2sub(x, __half22float2(h));
    const __half2 l = __floats2half2_rn(r.x, r.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// x-pass + split + tcgen05.st of group g (rows staged in s.L) into ring slot g&1; arrives on bar_a_full
__device__ __forceinline__ void produce_group(TcSmem& s, int g, int tid, uint32_t tmem_lane)
{
    const uint32_t ta = tmem_lane + TM_A + (uint32_t) (g & 1) * 4u;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint32_t hi[8][2], lo[8][2];
#pragma unroll
        for (int pp = 0; pp < 2; pp++) {
            const int p = half * 2 + pp;
            float2 v[8], X[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = s.L[p][tid + j];
            dctc_dct_fwd2<8>(v, X);
#pragma unroll
            for (int k1 = 0; k1 < 8; k1++) split_pair(X[k1], hi[k1][pp], lo[k1][pp]);
        }
        if (half == 0 && g >= 2) {
            // slot g&1 still holds group g-2, read by the MMAs of step g-2: wait for their completion
            mbar_wait(smem_u32(&s.bar_a_free), (uint32_t) (g & 1));
            tc_fence_after();
        }
#pragma unroll
        for (int k1 = 0; k1 < 8; k1++) {
            tmem_st_x2(ta + (uint32_t) (k1 * 16 + half * 2), hi[k1][0], hi[k1][1]);
            tmem_st_x2(ta + (uint32_t) (k1 * 16 + 8 + half * 2), lo[k1][0], lo[k1][1]);
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");   // warp-wide: every lane's stores have completed
    tc_fence_before();
    __syncwarp();
    // Group 0 is not announced on its own: the arrival of group 1 covers both (same warp, program order), so the MMA
    // warp needs exactly one parity wait per step and can never be two phases behind the producers.
    if (g >= 1 && (tid & 31) == 0) mbar_arrive(smem_u32(&s.bar_a_full));   // one arrival per producer warp
}

// ---- consumer fold -------------------------------------------------------------------------------------------
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
    template <int K1>
    __device__ __forceinline__ void add(const uint32_t (&v)[64])
    {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float t = m[i];
            if (K1 != 0) t = fmaxf(t, fabsf(__uint_as_float(v[i * 8])));   // (0,0) is skipped (src/dct.c:101)
            t = fmaxf(t, fabsf(__uint_as_float(v[i * 8 + 1])));
#pragma unroll
            for (int k2 = 2; k2 < 8; k2 += 2)
                t = fmaxf(t, fmaxf(fabsf(__uint_as_float(v[i * 8 + k2])), fabsf(__uint_as_float(v[i * 8 + k2 + 1]))));
            m[i] = t;
        }
    }
    __device__ __forceinline__ float result(int i, float we, float wt) const { (void) we; return m[i] * wt; }
};

template <>
struct TcFold<false> {  // last-arg-max class rule of DctcTracker<false>
    float a[8], mm[8], bv[8], z[8];
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int i = 0; i < 8; i++) { a[i] = 0.0f; mm[i] = -1.0f; bv[i] = 0.0f; z[i] = 0.0f; }
    }
    template <int K1>
    __device__ __forceinline__ void add(const uint32_t (&v)[64])
    {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (K1 == 0) {
                a[i] = fabsf(__uint_as_float(v[i * 8 + 1]));
#pragma unroll
                for (int k2 = 2; k2 < 8; k2++) mm[i] = fmaxf(mm[i], fabsf(__uint_as_float(v[i * 8 + k2])));
            } else {
                if (K1 == 1) bv[i] = fabsf(__uint_as_float(v[i * 8]));
                else z[i] = fmaxf(z[i], fabsf(__uint_as_float(v[i * 8])));
#pragma unroll
                for (int k2 = 1; k2 < 8; k2++) z[i] = fmaxf(z[i], fabsf(__uint_as_float(v[i * 8 + k2])));
            }
        }
    }
    __device__ __forceinline__ float result(int i, float we, float wt) const
    {
        const float am = fmaxf(a[i], mm[i]);
        const float top = fmaxf(fmaxf(am, bv[i]), z[i]);
        const bool tex = (z[i] >= fmaxf(am, bv[i])) || (!(bv[i] >= am) && (mm[i] >= a[i]));
        return top * (tex ? wt : we);
    }
};

template <int K1, bool UNIFORM>
__device__ __forceinline__ void consume_k1(TcSmem& s, TcFold<UNIFORM>& f, uint32_t tmem_lane, bool lane0)
{
    constexpr int b = K1 & 1, q = K1 >> 1;
    mbar_wait(smem_u32(&s.bar_d_full[b]), (uint32_t) (q & 1));
    tc_fence_after();
    uint32_t v[64];
    tmem_ld_x64(tmem_lane + TM_D + 64u * b, v);
    tc_fence_before();
    (void) lane0;
    bar_tile_arrive(b);
    f.template add<K1>(v);
}

// Persistent kernel: two CTAs per SM (256 TMEM columns each), work items = (frame, segment, strip) handed out by an
// atomic counter.
template <bool UNIFORM, int CH>
__global__ void __launch_bounds__(NTHREADS, 2) dctc_k1_tc8_kernel(const DctcK1Args a, int seg_rows, int strips, int segs, int n_items,
                                                                    int* __restrict__ counter)
{
    __shared__ TcSmem s;
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler

    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // Toeplitz operands: Tz[n = i*8 + k2][k] = B8[k2][r - i] with window row r = k (normal) or k ^ 8 (K-swapped)
    {
        uint16_t* Bq = reinterpret_cast<uint16_t*>(&s.B[0][0]);
        for (int idx = tid; idx < 4 * 1024; idx += NTHREADS) {
            const int v = idx >> 10, n = (idx >> 4) & 63, k = idx & 15;
            const int i = n >> 3, k2 = n & 7;
            const int c = ((v & 2) ? (k ^ 8) : k) - i;
            const uint16_t val = (c >= 0 && c < 8) ? DCTC_TC_BASIS8[v & 1][k2 * 8 + c] : (uint16_t) 0;
            Bq[v * 1024 + (n >> 3) * 128 + (k >> 3) * 64 + (n & 7) * 8 + (k & 7)] = val;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t lane_off = (uint32_t) ((warp & 3) * 32) << 16;
    bool first = true;

  for (;;) {
    if (tid == 0) {
        s.work = atomicAdd(counter, 1);
        if (!first) {
            mbar_inval(smem_u32(&s.bar_a_full));
            mbar_inval(smem_u32(&s.bar_a_free));
            mbar_inval(smem_u32(&s.bar_d_full[0]));
            mbar_inval(smem_u32(&s.bar_d_full[1]));
            mbar_inval(smem_u32(&s.bar_d_free[0]));
            mbar_inval(smem_u32(&s.bar_d_free[1]));
        }
        mbar_init(smem_u32(&s.bar_a_full), 4);
        mbar_init(smem_u32(&s.bar_a_free), 1);
        mbar_init(smem_u32(&s.bar_d_full[0]), 1);
        mbar_init(smem_u32(&s.bar_d_full[1]), 1);
        mbar_init(smem_u32(&s.bar_d_free[0]), 4);
        mbar_init(smem_u32(&s.bar_d_free[1]), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    first = false;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int item = s.work;
    if (item >= n_items) break;
    const uint32_t tmem = s.tmem_base;
    const uint32_t tmem_lane = tmem + lane_off;
    const int strip = item % strips;
    const int rest = item / strips;
    const int seg = rest % segs, frame = rest / segs;
    const int x0 = strip * MW;
    const int y0 = seg * seg_rows;
    const int y1 = min(y0 + seg_rows, a.h);
    const int nsteps = (y1 - y0 + 7) >> 3;
    const uint8_t* __restrict__ img = a.img + (size_t) frame * a.frame_stride;
    float* __restrict__ out = a.out + (size_t) frame * a.out_frame_stride;

    if (warp < 4) {
        // ===== producers: group g = virtual rows y0-3+8g .. y0+4+8g; step j consumes groups j and j+1 =====
        stage_raw_async<CH>(a, img, s.Raw[0], y0 - 3, x0, tid);
        stage_raw_async<CH>(a, img, s.Raw[1], y0 + 5, x0, tid);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        bar_producers();
        convert_raw<CH>(a, s.Raw[0], s.L, x0, tid);
        bar_producers();
        produce_group(s, 0, tid, tmem_lane);
        for (int g = 1; g <= nsteps; g++) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            bar_producers();                                  // Raw[g&1] landed everywhere, s.L free
            if (g + 1 <= nsteps) stage_raw_async<CH>(a, img, s.Raw[(g + 1) & 1], y0 - 3 + 8 * (g + 1), x0, tid);
            convert_raw<CH>(a, s.Raw[g & 1], s.L, x0, tid);
            bar_producers();
            produce_group(s, g, tid, tmem_lane);
        }
    } else if (warp < 8) {
        // ===== consumers =====
        const int px = tid - 128;
        const int gx = x0 + px;
        const bool lane0 = (tid & 31) == 0;
        for (int st = 0; st < nsteps; st++) {
            TcFold<UNIFORM> f;
            f.init();
            consume_k1<0, UNIFORM>(s, f, tmem_lane, lane0);
            consume_k1<1, UNIFORM>(s, f, tmem_lane, lane0);
            consume_k1<2, UNIFORM>(s, f, tmem_lane, lane0);
            consume_k1<3, UNIFORM>(s, f, tmem_lane, lane0);
            consume_k1<4, UNIFORM>(s, f, tmem_lane, lane0);
            consume_k1<5, UNIFORM>(s, f, tmem_lane, lane0);
            consume_k1<6, UNIFORM>(s, f, tmem_lane, lane0);
            consume_k1<7, UNIFORM>(s, f, tmem_lane, lane0);
            const int gy = y0 + 8 * st;
            if (gx < a.w) {
                float* __restrict__ o = out + (size_t) gy * a.out_pitch + gx;
                if (gy + 8 <= y1) {
#pragma unroll
                    for (int i = 0; i < 8; i++) o[(size_t) i * a.out_pitch] = f.result(i, a.w_edges, a.w_textures);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        if (gy + i < y1) o[(size_t) i * a.out_pitch] = f.result(i, a.w_edges, a.w_textures);
                }
            }
        }
    } else {
        // ===== MMA issuer =====
        const uint32_t idesc = make_idesc(128, 64);
        const uint64_t bd0 = make_smem_desc(smem_u32(&s.B[0][0]), 128, 256);
        for (int st = 0; st < nsteps; st++) {
            mbar_wait(smem_u32(&s.bar_a_full), (uint32_t) (st & 1));         // groups st and st+1 are in TMEM
            tc_fence_after();
            // each operand copy is 2048 bytes = 128 descriptor address units
            const uint64_t bh = bd0 + (uint64_t) ((st & 1) ? 256 : 0);
            const uint64_t bl = bh + 128;
#pragma unroll
            for (int k1 = 0; k1 < 8; k1++) {
                const int b = k1 & 1;
                if (st > 0 || k1 >= 2) bar_tile_sync(b);     // the consumers have loaded the previous contents of tile b
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d = tmem + TM_D + 64u * b;
                    const uint32_t ah = tmem + TM_A + (uint32_t) k1 * 16u, al = ah + 8u;
                    mma_ts(d, ah, bh, idesc, 0u);
                    mma_ts(d, al, bh, idesc, 1u);
                    mma_ts(d, ah, bl, idesc, 1u);
                    mma_commit(smem_u32(&s.bar_d_full[b]));
                    if (k1 == 7 && st + 2 <= nsteps) mma_commit(smem_u32(&s.bar_a_free));   // waited on by the producers of group st+2
                }
                __syncwarp();
            }
        }
        // the last use of each tile was arrived on but never waited for: drain both named barriers so that their
        // generations start clean for the next work item
        bar_tile_sync(0);
        bar_tile_sync(1);
    }
    tc_fence_before();
    __syncthreads();    // every role is done with the barriers, TMEM and staging buffers of this item
  }

    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s.tmem_base), "r"(TMEM_COLS));
}

}  // namespace

// Returns cudaErrorNotSupported when the configuration is outside this kernel's fast path (the caller then uses
// the FP32 march kernel): needs 1 or 3 channels and 16-byte aligned row pointers / pitches.
// `counter` is a device int owned by the context (work-item counter of the persistent kernel).
cudaError_t dctc_launch_k1_tc8(const DctcK1Args& a, int n_frames, bool uniform, int* counter, int sm_count, cudaStream_t stream)
{
    if (a.w <= 0 || a.h <= 0 || n_frames <= 0) return cudaSuccess;
    if (a.seam) return cudaErrorInvalidValue;  // band mode lives in the tile kernel
    auto aligned16 = [](const void* p, size_t pitch) { return (((uintptr_t) p | pitch) & 15) == 0; };
    const bool fast = (a.channels == 3 || a.channels == 1) && aligned16(a.img, a.pitch) && (a.frame_stride & 15) == 0 &&
                      (!a.top || aligned16(a.top, a.top_pitch)) && (!a.bot || aligned16(a.bot, a.bot_pitch));
    if (!fast || !counter) return cudaErrorNotSupported;
    const int strips = (a.w + MW - 1) / MW;
    // segment height: long segments amortise the 8-row prologue; keep >= ~6 items per SM so the tail stays short
    int seg = 512;
    while (seg > 32 && (long long) strips * ((a.h + seg - 1) / seg) * n_frames < 12LL * sm_count) seg >>= 1;
    const int segs = (a.h + seg - 1) / seg;
    const long long items = (long long) strips * segs * n_frames;
    if (items > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    const int grid = (int) (items < 2LL * sm_count ? items : 2LL * sm_count);
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;
#define DCTC_TC_LAUNCH(U, C)                                                                                           \
    do {                                                                                                               \
        cudaError_t ea = cudaFuncSetAttribute(dctc_k1_tc8_kernel<U, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAD_SMEM); \
        if (ea != cudaSuccess) return ea;                                                                              \
        dctc_k1_tc8_kernel<U, C><<<grid, NTHREADS, PAD_SMEM, stream>>>(a, seg, strips, segs, (int) items, counter);    \
    } while (0)
    if (a.channels == 3) { if (uniform) DCTC_TC_LAUNCH(true, 3); else DCTC_TC_LAUNCH(false, 3); }
    else { if (uniform) DCTC_TC_LAUNCH(true, 1); else DCTC_TC_LAUNCH(false, 1); }
#undef DCTC_TC_LAUNCH
    return cudaGetLastError();
}
