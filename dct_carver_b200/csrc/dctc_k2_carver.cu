// K2: carver session with incremental per-seam energy.
//
// Replaces liblqr's per-seam "carve + update_emap" pair as driven by lqr_carver_resize (src/render.c:377):
// after a vertical seam is removed, liblqr re-invokes the energy callback (src/render.c:134-157) for the pixels
// within +-radius of the seam (radius = blocksize/2, registered at src/render.c:314-315).  Here the image and the
// energy plane stay resident in HBM; one kernel compacts both over the seam, then the K1 tile kernel runs in band
// mode over just the touched pixels, with arithmetic identical to a full recompute (bit-identical results).
#include <cuda.h>      // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, no libcuda link)
#include <cstdio>
#include <cstdlib>
#include <type_traits>
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
                                                             const int* __restrict__ seam, int w_old,
                                                             float* __restrict__ mplane, size_t m_pitch)
{
    DCTC_PDL_PROLOGUE();
    const int y = blockIdx.x;
    const int s = seam[y];
    uint8_t* row = img + (size_t) y * pitch;
    float* erow = en + (size_t) y * en_pitch;
    // cumulative map of the seam DP (liblqr keeps it per pixel, so it moves with the pixels): dst x in [s, w_old-1)
    if (mplane) {
        float* mrow = mplane + (size_t) y * m_pitch;
        for (int base = s; base < w_old - 1; base += NT * K) {
            float v[K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int x = base + k * NT + threadIdx.x;
                v[k] = x < w_old - 1 ? mrow[x + 1] : 0.0f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int x = base + k * NT + threadIdx.x;
                if (x < w_old - 1) mrow[x] = v[k];
            }
        }
    }
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
// the whole row y-1: a chain of h steps.  Versions with one CTA-wide barrier per row cost ~600-950 clk per row
// (shared-memory bandwidth, then the latency of a 16-warp barrier round); this one removes the per-row barrier with
// warp-private trapezoids: a warp owns a strip of 128*P columns, keeps the cells of the previous row in registers
// (4*P adjacent cells per lane) and exchanges neighbours with warp shuffles only.  What the strip's edge lanes
// cannot see contaminates one more column per row, so after DP_R rows the outer DP_R columns on each side are
// garbage and the warp publishes only its central 128*P - 2*DP_R columns; every DP_R rows the warps exchange the
// last row through shared memory (one barrier per DP_R rows) and restart with fresh halos.  Energies are staged
// three blocks ahead by bulk async copies (TMA variant) or prefetched into a register ring; the cumulative rows go to a global float plane from which warp 0 re-derives
// the parent choices during the back-track (dp_backtrack: 32-row batches through double-buffered shared-memory windows).
// Cells outside the image hold +inf, which reproduces the range clipping.
#ifndef DP_EXP
#define DP_EXP 0   // timing experiments only
#endif
constexpr int DP_R = 32;          // rows between two exchanges = halo columns on each side of a strip
constexpr int DP_MAXW = 8;        // warps per CTA of the DP cluster (256 threads: room for a deep register prefetch ring)

constexpr int DP_CL = 8;          // CTAs of the cluster the strips are spread over (a lone SM can only pull ~35 GB/s from L2:
                                  // measured 425 clk per 1920-px row for the energy loads alone, against ~40 clk of math)

__device__ __forceinline__ uint32_t dp_cta_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void dp_cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `p` (a shared-memory location of this CTA's layout) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dp_map(const void* p, uint32_t rank)
{
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"((uint32_t) __cvta_generic_to_shared(p)), "r"(rank));
    return ra;
}
__device__ __forceinline__ float4 dp_ld_cluster_v4(uint32_t ra)
{
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra) : "memory");
    return v;
}
// 16 bytes into a peer CTA's shared memory; their arrival is counted on the peer's mbarrier (complete_tx), so the
// receiver needs no fence and no cluster barrier: it waits on its own mbarrier for the bytes it expects
__device__ __forceinline__ void dp_st_async_v4(uint32_t remote_addr, float4 v, uint32_t remote_mbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(remote_addr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void dp_st_cluster_f32(uint32_t ra, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory"); }
__device__ __forceinline__ void dp_st_cluster_s32(uint32_t ra, int v) { asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(ra), "r"(v) : "memory"); }

// ---- back-track (build_vpath) ------------------------------------------------------------------------------------
// One warp walks up from the last row: parent of (y, x) = FIRST strict minimum of m[y-1][x-1], m[y-1][x], m[y-1][x+1].
// Every lane follows the same path, lane 0 records it.  The cumulative rows are read in batches of 32 rows through
// shared-memory windows of DP_WIN columns, double buffered: the path moves at most one column per row, so the window
// of the NEXT batch (columns xs-66 .. xs+66 around the column xs at which the current batch starts) is known before
// the current batch is walked, and its cp.async copies land while the walk (a chain of dependent shared-memory reads)
// is running.
constexpr int DP_WIN = 144;       // back-track window (floats); the cumulative plane's pitch is at least this
constexpr int DP_WINP = DP_WIN + 8;   // window row in shared memory: 3 pad floats, the left sentinel, DP_WIN columns, the right
                                      // sentinel (needed when the window ends exactly at column w) and 3 pad floats

__device__ __forceinline__ void dp_backtrack(const float* __restrict__ mplane, size_t m_pitch, int w, int h, int x,
                                             int* __restrict__ seam, int* __restrict__ seam_log,
                                             float (*win)[32][DP_WINP], int lane)
{
    const float INF = __int_as_float(0x7f800000);
    if (lane == 0) { seam[h - 1] = x; if (seam_log) seam_log[h - 1] = x; }
    const int max_base = (int) m_pitch - DP_WIN;
    // window rows of the batch that starts below image row ytop: window row i = image row ytop-1-i, column c of the
    // plane at win[buf][i][4 + c - base]
    auto stage = [&](int ytop, int xs, int buf) -> int {
        int base = (xs - 66) & ~3;
        base = base < 0 ? 0 : (base > max_base ? max_base : base);
        if (ytop >= 1) {
            for (int idx = lane; idx < 32 * (DP_WIN / 4); idx += 32) {
                const int i = idx / (DP_WIN / 4), c = idx - i * (DP_WIN / 4);
                const int yy = ytop - 1 - i;
                if (yy >= 0) {
                    const float* src = mplane + (size_t) yy * m_pitch + base + 4 * c;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t) __cvta_generic_to_shared(&win[buf][i][4 + 4 * c])), "l"(src) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        return base;
    };
    int buf = 0;
    int base = stage(h - 1, x, 0);
    for (int ytop = h - 1; ytop >= 1; ytop -= 32) {
        const int base_next = stage(ytop - 32, x, buf ^ 1);     // lands while this batch is walked
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        // +inf sentinels left of column 0 and at column w (range clipping), so that the walk needs no bounds checks
        if (base == 0) win[buf][lane][3] = INF;
        if (w - base >= 0 && w - base <= DP_WIN) win[buf][lane][4 + w - base] = INF;   // <=: w == base + DP_WIN lands in the row's tail pad
        __syncwarp();
        const int steps = ytop < 32 ? ytop : 32;
        // the walk: three shared-memory reads at immediate offsets, two compares, a few selects per row; lane i remembers
        // the column of step i and the 32 columns are written with one coalesced store at the end of the batch
        int mine = 0;
        const float* wp = &win[buf][0][4] + (x - base);        // wp[0] = cumulative value at (window row, column x)
        auto walk = [&](int i) {
            const float a = wp[-1], b = wp[0], c = wp[1];
            const bool p1 = b < a;
            const float best = p1 ? b : a;
            const bool p2 = c < best;
            const int d = p2 ? 1 : (p1 ? 0 : -1);
            x += d;
            wp += DP_WINP + d;
            if (lane == i) mine = x;
        };
        if (steps == 32) {
#pragma unroll
            for (int i = 0; i < 32; i++) walk(i);
        } else {
            for (int i = 0; i < steps; i++) walk(i);
        }
        if (lane < steps) { seam[ytop - 1 - lane] = mine; if (seam_log) seam_log[ytop - 1 - lane] = mine; }
        __syncwarp();
        base = base_next;
        buf ^= 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// TMA = true: the energies of a strip are staged nst chunks of DP_SR rows ahead into shared memory with 1-D bulk
// async copies (cp.async.bulk, completion on an mbarrier per warp and stage) instead of a register ring of plain
// loads: nothing the cluster barrier's memory fence has to wait for, and 48 rows in flight per warp.
constexpr int DP_NST_MAX = 4;     // stages per warp: as many (2..4) as fit into shared memory, chosen by the host

__device__ __forceinline__ uint32_t dp_smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
// bounded parity wait: a protocol error traps after ~2^24 polls instead of hanging the GPU
__device__ __forceinline__ void dp_mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .u32 n;\n"
        "mov.u32 n, 0;\n"
        "DP_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DP_DONE;\n"
        "add.u32 n, n, 1;\n"
        "setp.lt.u32 p, n, 0x1000000;\n"
        "@p bra DP_WAIT;\n"
        "trap;\n"
        "DP_DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}

// compile-time loop: f(std::integral_constant<int, 0>) ... f(std::integral_constant<int, N - 1>)
template <int N, int I = 0, typename F>
__device__ __forceinline__ void dp_unroll(F&& f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        dp_unroll<N, I + 1>(f);
    }
}

// DP_P float4 groups per lane: strip = 128 * DP_P columns, 128 * DP_P - 2 * DP_R of them published.
// DP_SR: rows per staged chunk of the bulk-copy variant (32 when two stages per warp fit, else 16), 0 = register ring.
template <int DP_P, int DP_SR>
__global__ void __cluster_dims__(DP_CL, 1, 1) __launch_bounds__(DP_MAXW * 32) dctc_seam_dp_kernel(const float* __restrict__ en, size_t en_pitch, int w, int h,
                                                                    float* __restrict__ mplane, size_t m_pitch,
                                                                    int* __restrict__ seam, int* __restrict__ seam_log,
                                                                    int* __restrict__ run_flag, int nst,
                                                                    const __grid_constant__ CUtensorMap tmap, int use_tmap,
                                                                    int* __restrict__ xlast)
{
    DCTC_PDL_PROLOGUE();
    // fallback of the incremental update (dctc_seam_incr_kernel): runs only when that kernel asked for a rebuild
    if (run_flag && *run_flag == 0) return;       // uniform over the cluster, before any cluster barrier
    if (run_flag && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(run_flag + 1, 1);   // rebuilds since the session was loaded
    constexpr bool TMA = DP_SR > 0;
    constexpr int SRD = TMA ? DP_SR : DP_R;       // (divisor that is never zero)
    constexpr int CPL = 4 * DP_P;                 // cells per lane
    constexpr int STRIP = 32 * CPL;               // columns a warp computes
    constexpr int WOUT = STRIP - 2 * DP_R;        // columns a warp publishes
    constexpr int HL = DP_R / CPL;                // halo lanes on each side (DP_R is a multiple of CPL)
    constexpr int PF = TMA ? 1 : 16 / DP_P;       // energy rows in flight per lane (register ring of the non-TMA variant)
    extern __shared__ __align__(128) float xrow[]; // two exchange rows of (warps per CTA * WOUT) floats, double buffered;
                                                  // TMA: followed by warps x nst stages x DP_SR rows x STRIP energies
    __shared__ __align__(8) unsigned long long ebar[DP_MAXW][DP_NST_MAX];
    __shared__ __align__(8) unsigned long long hbar[DP_MAXW][2];   // halo exchange: per warp and slot, counts the neighbours' bytes
    __shared__ __align__(8) unsigned long long rbar;               // CTA 0: counts the strips' last-row minima (8 bytes each)
    __shared__ __align__(8) int2 red_p[DP_CL * DP_MAXW];           // CTA 0: (value bits, column) per strip
    __shared__ float red_v[DP_CL * DP_MAXW];      // per-strip minima, gathered in CTA 0 through distributed shared memory
    __shared__ int red_i[DP_CL * DP_MAXW];
    __shared__ __align__(16) float win[2][32][DP_WINP];
    const int tid = threadIdx.x, lane = tid & 31;
    const int wpc = blockDim.x >> 5;                        // warps (strips) per CTA
    const uint32_t rank = dp_cta_rank();
    const int warp = (int) rank * wpc + (tid >> 5);         // strip index across the cluster
    const int nwarps = DP_CL * wpc;
    const int xlen = wpc * WOUT;
    const float INF = __int_as_float(0x7f800000);
    const int c0 = warp * WOUT - DP_R + lane * CPL;   // first column of this lane (may be < 0 or >= w)
    const bool central = lane >= HL && lane < 32 - HL;

    // Per-lane constants of the strip geometry: which groups start inside the image, which of their cells lie right of
    // it.  Cells outside the image must stay +inf in every row (range clipping); the patch is applied to the computed
    // row, never to a freshly loaded value -- touching the loaded registers right after the load would turn the
    // 8-rows-ahead prefetch into a synchronous load (measured: 425 clk per row instead of ~40).
    bool ld_ok[DP_P];
    unsigned inf_mask[DP_P];
#pragma unroll
    for (int g = 0; g < DP_P; g++) {
        const int x = c0 + 4 * g;
        ld_ok[g] = x >= 0 && x < w;              // x is a multiple of 4 and the row pitch covers round_up(w, 4)
        inf_mask[g] = !ld_ok[g] ? 15u : ((x + 1 >= w ? 2u : 0u) | (x + 2 >= w ? 4u : 0u) | (x + 3 >= w ? 8u : 0u));
    }
    auto patch = [&](float4& v, unsigned m) {
        if (m & 1u) v.x = INF;
        if (m & 2u) v.y = INF;
        if (m & 4u) v.z = INF;
        if (m & 8u) v.w = INF;
    };
    // Every lane loads unconditionally (lanes / rows outside the image re-read a valid address and their cells are
    // patched to +inf): a predicated load is compiled into "load to a temporary + predicated move", i.e. a wait for the
    // data right after the issue, which again serialises the prefetch.
    const float* en_g[DP_P];
#pragma unroll
    for (int g = 0; g < DP_P; g++) en_g[g] = en + (ld_ok[g] ? c0 + 4 * g : 0);
    const int hm1 = h - 1;

    // TMA staging of this warp's strip: columns [cs, cs + STRIP) clipped to the row, DP_R rows per stage
    const int cs = warp * WOUT - DP_R;
    const int tx0 = max(cs, 0), tx1 = min(cs + STRIP, (int) en_pitch);
    const int tbytes = tx1 > tx0 ? (tx1 - tx0) * 4 : 0;
    float* estage = xrow + 2 * xlen + (size_t) (tid >> 5) * nst * SRD * STRIP;
    const int nchunks = (h - 1 + SRD - 1) / SRD;
    auto issue_chunk = [&](int ch, int st) {   // rows 1 + ch * DP_SR .. (clamped to h - 1) -> stage st = ch % nst
        const uint32_t bar = dp_smem_u32(&ebar[tid >> 5][st]);
        if (use_tmap) {
            // one 2-D tensor copy (box = STRIP columns x DP_SR rows; columns left of the image and rows below it are
            // zero-filled by the TMA unit) instead of DP_SR row copies, which the hardware issues one elected lane at a time
            if (lane == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t) (SRD * STRIP * 4)) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(dp_smem_u32(estage + (size_t) st * SRD * STRIP)), "l"(&tmap), "r"(cs), "r"(1 + ch * SRD), "r"(bar) : "memory");
            }
            __syncwarp();
            return;
        }
        if (lane == 0) {
            if (tbytes > 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t) (tbytes * SRD)) : "memory");
            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
        }
        __syncwarp();
        if (lane < SRD && tbytes > 0) {
            const int y = min(1 + ch * SRD + lane, hm1);
            const float* src = en + (size_t) y * en_pitch + tx0;
            const uint32_t dst = dp_smem_u32(estage + ((size_t) st * SRD + lane) * STRIP + (tx0 - cs));
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(src), "r"((uint32_t) tbytes), "r"(bar) : "memory");
        }
    };
    if (lane == 0) {
        if (TMA) {
#pragma unroll
            for (int st = 0; st < DP_NST_MAX; st++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dp_smem_u32(&ebar[tid >> 5][st])) : "memory");
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dp_smem_u32(&hbar[tid >> 5][0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dp_smem_u32(&hbar[tid >> 5][1])) : "memory");
        if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dp_smem_u32(&rbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (TMA) {
        for (int ch = 0; ch < nst && ch < nchunks; ch++) issue_chunk(ch, ch);
    }
    dp_cluster_sync();   // once: every CTA's halo mbarriers exist before a neighbour's first st.async can arrive

    float4 cur[DP_P], e[PF][DP_P];
#pragma unroll
    for (int g = 0; g < DP_P; g++) {
        cur[g] = __ldg(reinterpret_cast<const float4*>(en_g[g]));
        const int x = c0 + 4 * g;
        if (central && ld_ok[g]) *reinterpret_cast<float4*>(mplane + x) = cur[g];
        patch(cur[g], inf_mask[g]);
        if (!TMA) {
#pragma unroll
            for (int k = 0; k < PF; k++)               // e[k] holds row y with (y - 1) % PF == k
                e[k][g] = __ldg(reinterpret_cast<const float4*>(en_g[g] + (size_t) min(1 + k, hm1) * en_pitch));
        }
    }
#ifdef DCTC_SYNC_DEBUG
    const long long dbg_t0 = clock64();
#endif
    // The row chain is bound by the number of instructions a lone warp per scheduler has to issue (the ncu source view
    // shows ~6 clk per instruction, mostly fixed-latency "wait" stalls), so the loop body is kept minimal: staged
    // energies are read with immediate offsets from a shared-space base register, cells outside the image are made
    // +inf once per chunk in the staged copy (inf + min = inf, no per-row patch), the cumulative plane is written
    // through one running pointer, and whole blocks run without per-row bounds checks.
    float* mptr = mplane + m_pitch + c0;             // this lane's first cell of row y
    unsigned any_mask = 0;
#pragma unroll
    for (int g = 0; g < DP_P; g++) any_mask |= inf_mask[g];
    int xb = 0;
    int hround = 0;                                  // halo exchanges done: slot = hround & 1, mbarrier parity = (hround >> 1) & 1
    int c_ch = 0, c_st = 0;                          // chunk being consumed, its stage and mbarrier parity
    uint32_t c_par = 0;
    uint32_t stg_s = 0;                              // shared-space address of this lane's cells in the current chunk
    auto chunk_begin = [&]() {                       // a new chunk of staged rows starts: wait for it, patch its +inf cells
        dp_mbar_wait(dp_smem_u32(&ebar[tid >> 5][c_st]), c_par);
        float* sp = estage + (size_t) c_st * SRD * STRIP + lane * CPL;
        stg_s = dp_smem_u32(sp);
        if (any_mask) {
            for (int rr = 0; rr < SRD; rr++) {
#pragma unroll
                for (int g = 0; g < DP_P; g++) {
                    float* q = sp + rr * STRIP + 4 * g;
                    if (inf_mask[g] & 1u) q[0] = INF;
                    if (inf_mask[g] & 2u) q[1] = INF;
                    if (inf_mask[g] & 4u) q[2] = INF;
                    if (inf_mask[g] & 8u) q[3] = INF;
                }
            }
        }
    };
    auto chunk_end = [&]() {                         // this warp is done with the chunk: refill its stage nst chunks ahead
        __syncwarp();
        if (c_ch + nst < nchunks) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_chunk(c_ch + nst, c_st);
        }
        c_ch++;
        if (++c_st == nst) { c_st = 0; c_par ^= 1u; }
    };
    auto row = [&](auto kk_c, int y) {
        constexpr int kk = decltype(kk_c)::value;
        constexpr int k = kk % PF;
        if (TMA) {
#pragma unroll
            for (int g = 0; g < DP_P; g++)
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(e[0][g].x), "=f"(e[0][g].y), "=f"(e[0][g].z), "=f"(e[0][g].w)
                             : "r"(stg_s + (uint32_t) (((kk % SRD) * STRIP + 4 * g) * 4)));   // base register + immediate
        }
        // neighbours across lanes; the strip's outermost lanes see +inf (their cells are never published)
        float l = __shfl_up_sync(0xffffffffu, cur[DP_P - 1].w, 1);
        float r = __shfl_down_sync(0xffffffffu, cur[0].x, 1);
        if (lane == 0) l = INF;
        if (lane == 31) r = INF;
        float4 o[DP_P];
#pragma unroll
        for (int g = 0; g < DP_P; g++) {
            const float lg = g == 0 ? l : cur[g > 0 ? g - 1 : 0].w;
            const float rg = g == DP_P - 1 ? r : cur[g < DP_P - 1 ? g + 1 : g].x;
            o[g].x = e[k][g].x + fminf(fminf(lg, cur[g].x), cur[g].y);
            o[g].y = e[k][g].y + fminf(fminf(cur[g].x, cur[g].y), cur[g].z);
            o[g].z = e[k][g].z + fminf(fminf(cur[g].y, cur[g].z), cur[g].w);
            o[g].w = e[k][g].w + fminf(fminf(cur[g].z, cur[g].w), rg);
        }
        // register-ring variant: refill the slot just consumed with the row PF ahead (issued after the last use of the
        // slot, so the load targets the slot's registers directly; they are not touched again for PF rows)
#if !(DP_EXP & 1)
        if (!TMA) {
#pragma unroll
            for (int g = 0; g < DP_P; g++)
                e[k][g] = __ldg(reinterpret_cast<const float4*>(en_g[g] + (size_t) min(y + PF, hm1) * en_pitch));
        }
#endif
#pragma unroll
        for (int g = 0; g < DP_P; g++) {
#if DP_EXP & 2
            if (o[g].x == 12345.678f)
#endif
            if (central && ld_ok[g]) *reinterpret_cast<float4*>(mptr + 4 * g) = o[g];
            if (!TMA) patch(o[g], inf_mask[g]);      // cells outside the image stay +inf (TMA: their staged energy is +inf)
            cur[g] = o[g];
        }
        mptr += m_pitch;
    };
    for (int yb = 1; yb < h; yb += DP_R) {
        if (yb + DP_R <= h) {                         // whole block: no per-row bounds checks
            dp_unroll<DP_R>([&](auto kk_c) {
                constexpr int kk = decltype(kk_c)::value;
                if (TMA && kk % SRD == 0) chunk_begin();
                row(kk_c, yb + kk);
                if (TMA && kk % SRD == SRD - 1) chunk_end();
            });
        } else {
            dp_unroll<DP_R>([&](auto kk_c) {
                constexpr int kk = decltype(kk_c)::value;
                const int y = yb + kk;
                if (TMA && kk % SRD == 0 && y < h) chunk_begin();
                if (y < h) row(kk_c, y);              // uniform across the CTA
                if (TMA && kk % SRD == SRD - 1) chunk_end();
            });
        }
        // Exchange the last row of the block with the two neighbouring strips only: the outer 32 published cells on each
        // side go straight into the neighbour warp's halo slot (st.async into its CTA's shared memory, completion counted
        // on ITS mbarrier), then this warp waits for the 2 x 128 bytes it expects itself.  No cluster-wide barrier and no
        // memory fence: the release/acquire cluster barrier used before carried MEMBAR.ALL.GPU + CCTL.IVALL, i.e. it also
        // waited for the block's global stores of the cumulative plane (40 % of the kernel's stall samples).
        // Slot = block parity: a neighbour can be at most one block ahead, it needs this warp's row to go further.
        if (yb + DP_R < h && !(DP_EXP & 4)) {
            const int wl = tid >> 5;                          // warp inside the CTA
            float* halo = xrow + (wl * 2 + xb) * 64;          // [0,32): cells left of the published range, [32,64): right
            const uint32_t my_bar = dp_smem_u32(&hbar[wl][xb]);
            if (lane == 0) {
                const uint32_t expect = (warp > 0 ? 128u : 0u) + (warp < nwarps - 1 ? 128u : 0u);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(my_bar), "r"(expect) : "memory");
            }
            if (lane >= HL && lane < 2 * HL && warp > 0) {                        // my leftmost 32 published cells
                const int nw = warp - 1;
                const uint32_t nrank = (uint32_t) (nw / wpc);
                const int nwl = nw - (int) nrank * wpc;
                const uint32_t dst = dp_map(xrow + (nwl * 2 + xb) * 64 + 32 + (lane - HL) * CPL, nrank);
                const uint32_t bar = dp_map(&hbar[nwl][xb], nrank);
#pragma unroll
                for (int g = 0; g < DP_P; g++) dp_st_async_v4(dst + 16u * g, cur[g], bar);
            }
            if (lane >= 32 - 2 * HL && lane < 32 - HL && warp < nwarps - 1) {     // my rightmost 32 published cells
                const int nw = warp + 1;
                const uint32_t nrank = (uint32_t) (nw / wpc);
                const int nwl = nw - (int) nrank * wpc;
                const uint32_t dst = dp_map(xrow + (nwl * 2 + xb) * 64 + (lane - (32 - 2 * HL)) * CPL, nrank);
                const uint32_t bar = dp_map(&hbar[nwl][xb], nrank);
#pragma unroll
                for (int g = 0; g < DP_P; g++) dp_st_async_v4(dst + 16u * g, cur[g], bar);
            }
            dp_mbar_wait(my_bar, (uint32_t) ((hround >> 1) & 1));
            hround++;
            if (!central) {
                const float* hp = halo + (lane < HL ? lane * CPL : 32 + (lane - (32 - HL)) * CPL);
#pragma unroll
                for (int g = 0; g < DP_P; g++) {
                    const int x = c0 + 4 * g;
                    if (x >= 0 && x < w) cur[g] = *reinterpret_cast<const float4*>(hp + 4 * g);   // cells right of the image arrive as +inf
                    else cur[g] = make_float4(INF, INF, INF, INF);
                }
            }
            xb ^= 1;
        }
    }
#ifdef DCTC_SYNC_DEBUG
    const long long dbg_t1 = clock64();
#endif
    // leftmost minimum of the last row (central cells only: the others are either duplicates or contaminated)
    float bv = INF;
    int bi = 0x7fffffff;
    if (central) {
#pragma unroll
        for (int g = 0; g < DP_P; g++) {
            const float c4[4] = {cur[g].x, cur[g].y, cur[g].z, cur[g].w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int x = c0 + 4 * g + k;
                if (x >= 0 && x < w && (c4[k] < bv || bi == 0x7fffffff)) { bv = c4[k]; bi = x; }   // ascending x: strict < keeps the leftmost
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    if (xlast) {
        // parallel back-track follows in other kernels (the launch boundary publishes the cumulative plane): the strips'
        // minima travel to CTA 0 like the halos (st.async + mbarrier), no cluster barrier and no memory fence at the end
        if (lane == 0)
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
                         ::"r"(dp_map(&red_p[warp], 0u)), "r"(__float_as_uint(bv)), "r"(bi), "r"(dp_map(&rbar, 0u)) : "memory");
        if (rank == 0 && tid < 32) {
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dp_smem_u32(&rbar)), "r"((uint32_t) nwarps * 8u) : "memory");
            dp_mbar_wait(dp_smem_u32(&rbar), 0u);
        }
    } else {
        if (lane == 0) { dp_st_cluster_f32(dp_map(&red_v[warp], 0u), bv); dp_st_cluster_s32(dp_map(&red_i[warp], 0u), bi); }
        dp_cluster_sync();   // release/acquire at cluster scope: also orders every CTA's global writes of the cumulative plane
    }
    if (rank == 0 && tid < 32) {
        bv = INF;
        bi = 0x7fffffff;
        for (int i = tid; i < nwarps; i += 32) {   // strips in ascending column order: strict < keeps the leftmost
            const float rv = xlast ? __int_as_float(red_p[i].x) : red_v[i];
            const int ri = xlast ? red_p[i].y : red_i[i];
            if (ri != 0x7fffffff && (bi == 0x7fffffff || rv < bv)) { bv = rv; bi = ri; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi != 0x7fffffff && (bi == 0x7fffffff || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
        }
        if (xlast) {
            // parallel back-track: the jump and trace kernels below take over from the seam's last-row column
            if (lane == 0) { *xlast = bi; seam[h - 1] = bi; if (seam_log) seam_log[h - 1] = bi; }
        } else {
            dp_backtrack(mplane, m_pitch, w, h, bi, seam, seam_log, win, lane);   // serial walk by warp 0 of CTA 0
        }
#ifdef DCTC_SYNC_DEBUG
        if (tid == 0) {
            unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            printf("dp kernel w %d h %d: chain %lld clk, reduce+backtrack %lld clk, globaltimer %llu ns\n", w, h, dbg_t1 - dbg_t0, clock64() - dbg_t1, gt);
        }
#endif
    }
}

// ---- parallel back-track (build_vpath without the h-step dependent chain) ---------------------------------------
// The serial walk above is a chain of h dependent shared-memory reads (133 clk per row: 73 us of a 1080-row seam).  The
// parent choice of a cell depends only on the cumulative map, not on the path, so the rows are cut into blocks of JB
// rows (counted upwards from the last row) and, for EVERY column x of a block's bottom row, the column its path has at
// the block's top row is computed in parallel: top-down, P_top[x] = x, P_y[x] = P_{y-1}[x + d_y(x)] with d_y(x) the
// parent direction of (y, x) -- neighbours only, so a warp keeps a strip of 128 columns in registers and exchanges
// the outer cells by shuffles, exactly like the seam DP (JB = 32 contaminated columns on each side, 64 published).
// dctc_seam_jump_kernel: one warp per (strip, block), the whole grid busy for ~1 us; it stores the jump as an int8
// offset.  dctc_seam_trace_kernel: one CTA per block; it follows the ceil((h-1)/32) jumps from the seam's last-row
// column down to its own block (the jump entries a path can reach form a cone of 64*j+1 columns in block j, staged
// into shared memory first), then walks its 32 rows like dp_backtrack does.  ~100 dependent steps instead of h.
constexpr int JB = 32;

__global__ void __launch_bounds__(128) dctc_seam_jump_kernel(const float* __restrict__ mplane, size_t m_pitch, int w, int h,
                                                             int8_t* __restrict__ jump, size_t j_pitch, int strips)
{
    DCTC_PDL_PROLOGUE();
    const int lane = threadIdx.x & 31;
    const int strip = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (strip >= strips) return;
    const int k = blockIdx.y;
    const int yb = h - 1 - JB * k;                    // bottom row of the block
    const int yt = max(yb - JB, 0);                   // top row
    const int steps = yb - yt;
    const float INF = __int_as_float(0x7f800000);
    const int c0 = strip * 64 - JB + 4 * lane;        // this lane's four columns
    const bool in_row = c0 >= 0 && c0 < w;            // (c0 is a multiple of 4 and the pitch covers round_up(w, 4))
    const float* src = mplane + (size_t) yt * m_pitch + (in_row ? c0 : 0);
    int P[4] = {c0, c0 + 1, c0 + 2, c0 + 3};
    auto row = [&](float4 v) {                        // v = cumulative values of row y-1 -> P of row y
        float m[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (!in_row || c0 + j >= w) m[j] = INF;   // outside the image: never a parent (range clipping)
        float l = __shfl_up_sync(0xffffffffu, m[3], 1), r = __shfl_down_sync(0xffffffffu, m[0], 1);
        int pl = __shfl_up_sync(0xffffffffu, P[3], 1), pr = __shfl_down_sync(0xffffffffu, P[0], 1);
        if (lane == 0) l = INF;
        if (lane == 31) r = INF;
        int Q[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float a = j == 0 ? l : m[j - 1], b = m[j], c = j == 3 ? r : m[j + 1];
            const int pa = j == 0 ? pl : P[j - 1], pb = P[j], pc = j == 3 ? pr : P[j + 1];
            const bool p1 = b < a;                    // first strict minimum scanning x-1, x, x+1
            const float best = p1 ? b : a;
            const bool p2 = c < best;
            Q[j] = p2 ? pc : (p1 ? pb : pa);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) P[j] = Q[j];
    };
    for (int i0 = 0; i0 < steps; i0 += 8) {           // eight independent loads in flight, then eight dependent steps
        float4 pv[8];
#pragma unroll
        for (int j = 0; j < 8; j++)
            pv[j] = __ldcg(reinterpret_cast<const float4*>(src + (size_t) min(i0 + j, steps - 1) * m_pitch));
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (i0 + j < steps) row(pv[j]);
    }
    if (lane >= JB / 4 && lane < 32 - JB / 4 && in_row) {
        int8_t* dst = jump + (size_t) k * j_pitch + c0;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (c0 + j < w) dst[j] = (int8_t) (P[j] - (c0 + j));
    }
}

constexpr int TR_WIN = 72;                            // walk window: columns x-34 .. x+37 around the block's bottom column
constexpr int TR_THREADS = 256;

__global__ void __launch_bounds__(TR_THREADS) dctc_seam_trace_kernel(const float* __restrict__ mplane, size_t m_pitch, int w, int h,
                                                              const int8_t* __restrict__ jump, size_t j_pitch,
                                                              const int* __restrict__ xlast, int* __restrict__ seam,
                                                              int* __restrict__ seam_log)
{
    extern __shared__ __align__(16) unsigned char tr_smem[];
    __shared__ __align__(16) float win[JB][TR_WIN + 8];
    __shared__ int xb_s;
    DCTC_PDL_PROLOGUE();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int k = blockIdx.x;
    const int x0 = *xlast;
    // Cone of block j: its bottom column can be x0 - 32 j .. x0 + 32 j.  The entries are staged as aligned 16-byte
    // chunks (at most 4 j + 3 per row, at byte offset 32 j^2 + 16 j of the cone buffer); warp w takes the rows
    // j = w, w + 8, ... and issues all chunks of a row before it stores any of them, so the loads of a row (and of the
    // other warps' rows) are in flight together instead of one L2 round trip per iteration.
    for (int j = wid; j < k; j += TR_THREADS / 32) {
        const int a0 = max(0, x0 - JB * j) & ~15;
        const int last = min(w - 1, x0 + JB * j);
        const int nchunk = ((last - a0) >> 4) + 1;                          // <= 4 j + 3
        const uint4* srcj = reinterpret_cast<const uint4*>(jump + (size_t) j * j_pitch + a0);
        uint4* dstj = reinterpret_cast<uint4*>(tr_smem + 32 * j * j + 16 * j);
        for (int c0 = 0; c0 < nchunk; c0 += 160) {                          // (one pass up to 1080 rows, two beyond)
            uint4 v[5];
#pragma unroll
            for (int i = 0; i < 5; i++)
                if (c0 + lane + 32 * i < nchunk) v[i] = __ldcg(srcj + c0 + lane + 32 * i);
#pragma unroll
            for (int i = 0; i < 5; i++)
                if (c0 + lane + 32 * i < nchunk) dstj[c0 + lane + 32 * i] = v[i];
        }
    }
    __syncthreads();
    if (tid == 0) {
        const int8_t* cone = reinterpret_cast<const int8_t*>(tr_smem);
        int x = x0;
        for (int j = 0; j < k; j++) x += cone[32 * j * j + 16 * j + (x - (max(0, x0 - JB * j) & ~15))];
        xb_s = x;
    }
    __syncthreads();
    int x = xb_s;
    const int yb = h - 1 - JB * k, yt = max(yb - JB, 0), steps = yb - yt;
    // window rows: win[i] = cumulative row yb-1-i, columns base .. base+TR_WIN-1 at win[i][4 + c - base]
    const int max_base = (int) m_pitch - TR_WIN;
    int base = (x - 34) & ~3;
    base = base < 0 ? 0 : (base > max_base ? max_base : base);
    for (int idx = tid; idx < steps * (TR_WIN / 4); idx += TR_THREADS) {
        const int i = idx / (TR_WIN / 4), c = idx - i * (TR_WIN / 4);
        const float* src = mplane + (size_t) (yb - 1 - i) * m_pitch + base + 4 * c;
        *reinterpret_cast<float4*>(&win[i][4 + 4 * c]) = __ldcg(reinterpret_cast<const float4*>(src));
    }
    __syncthreads();
    if (tid >= 32) return;
    const float INF = __int_as_float(0x7f800000);
    // +inf sentinels left of column 0 and at column w (range clipping), so that the walk needs no bounds checks
    if (lane < steps) {
        if (base == 0) win[lane][3] = INF;
        if (w - base >= 0 && w - base <= TR_WIN) win[lane][4 + w - base] = INF;
    }
    __syncwarp();
    int mine = 0;
    const float* wp = &win[0][4] + (x - base);
    for (int i = 0; i < steps; i++) {
        const float a = wp[-1], b = wp[0], c = wp[1];
        const bool p1 = b < a;
        const float best = p1 ? b : a;
        const bool p2 = c < best;
        const int d = p2 ? 1 : (p1 ? 0 : -1);
        x += d;
        wp += TR_WIN + 8 + d;
        if (lane == i) mine = x;
    }
    if (lane < steps) { seam[yb - 1 - lane] = mine; if (seam_log) seam_log[yb - 1 - lane] = mine; }
}

// ---- incremental cumulative map (liblqr's update_mmap) -----------------------------------------------------------
// After a seam is removed only part of the cumulative map changes: the cells inside the energy band of the removed seam
// (their energy, or the identity of their parents, changed) and, row after row, the cells below a cell whose value
// changed.  liblqr's update_mmap walks the rows with exactly that shrinking / growing column range instead of
// rebuilding the map (lqr_carver_resize, src/render.c:377).  Here one warp does the walk: the cumulative plane was
// compacted over the seam by the carve kernel, row y is recomputed for x in
//     R(y) = band(y)  U  [first changed cell of row y-1  - 1,  last changed cell of row y-1  + 1]
// with the same FP32 formula as the full rebuild, so the plane (and every later seam) is bit-identical to it.
// R(y) grows by at most one column per side and row, so the energies and old map values a row can need are known
// INCR_D rows ahead: they are staged with cp.async into a shared-memory ring, and the walk itself touches shared memory
// only (lane l owns the columns lo + l + 32k; the recomputed row is handed to the next one through a double buffer with
// two unchanged cells on each side).  R(y) wider than INCR_CAP columns raises the rebuild flag and the full DP kernel
// takes over for this seam.  The walk ends with the leftmost minimum of the last row and the same back-track as the
// full kernel.
// MEASURED (1920x1080 noise, 480 seams): the walk recomputes ~150 columns per row on average and needs the full rebuild
// for 2 of 480 seams, but ONE warp spends ~900 clk per row on it (phases per row: 250-470 clk row arithmetic, ~180 clk
// side cells + next range, ~340 clk staging; a lone warp has no other warp to hide its 4-6 clk dependent-issue
// latencies behind), i.e. 918 us per seam against 199 us for the cluster-wide rebuild.  It is therefore OFF by default
// (dctc_carver_set_incremental) and kept as the bit-identical reference point for a multi-warp block version.
#ifndef DCTC_INCR_ABL
#define DCTC_INCR_ABL 0   // timing ablations only
#endif
constexpr int INCR_CAP = 512;                           // widest recomputed range per row
constexpr int INCR_D = 8;                               // rows staged ahead
constexpr int INCR_T = 8;                               // rows between two tightenings of the recomputed range
constexpr int INCR_SW = INCR_CAP + 2 * INCR_D + 16;     // floats per staged row and plane

// one row of the walk with NK columns per lane: every shared-memory load first, then the arithmetic and the stores
template <int NK>
__device__ __forceinline__ void incr_row(int y, int lane, int lo, int hi, int w, int plo, const float* __restrict__ re,
                                         const float* __restrict__ rm, int xs, const float* __restrict__ prev,
                                         float* __restrict__ next, float* __restrict__ mrow, int& clo, int& chi)
{
    const float INF = __int_as_float(0x7f800000);
    float e[NK], mo[NK], pl[NK], pc[NK], pr[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) {
        const int xc = min(lo + lane + 32 * k, hi);
        e[k] = re[xc - xs];
        mo[k] = rm[xc - xs];
        const int i = y > 0 ? xc - plo + 2 : 1;
        pl[k] = prev[i - 1];
        pc[k] = prev[i];
        pr[k] = prev[i + 1];
    }
#pragma unroll
    for (int k = 0; k < NK; k++) {
        const int x = lo + lane + 32 * k;
        if (x <= hi) {
            float mn = e[k];
            if (y > 0) mn = e[k] + fminf(fminf(x > 0 ? pl[k] : INF, pc[k]), x < w - 1 ? pr[k] : INF);
            next[x - lo + 2] = mn;
            if (__float_as_uint(mn) != __float_as_uint(mo[k])) {
#if DCTC_INCR_ABL != 1
                mrow[x] = mn;
#endif
                clo = min(clo, x);
                chi = max(chi, x);
            }
        }
    }
}

__global__ void __launch_bounds__(32) dctc_seam_incr_kernel(const float* __restrict__ en, size_t en_pitch, int w, int h,
                                                            float* __restrict__ mplane, size_t m_pitch, int r,
                                                            int* __restrict__ seam, int* __restrict__ seam_log,
                                                            int* __restrict__ rebuild_flag)
{
    extern __shared__ __align__(16) int incr_sm[];
    float* ring = reinterpret_cast<float*>(incr_sm);                       // INCR_D slots x (energy row, map row) x INCR_SW
    float* buf = ring + INCR_D * 2 * INCR_SW;                              // two hand-over rows of INCR_CAP + 4 cells
    int* xsb = reinterpret_cast<int*>(buf + 2 * (INCR_CAP + 4));           // first staged column per slot
    int* bmin = xsb + INCR_D;                                              // band of the removed seam per row
    int* bmax = bmin + h;
    __shared__ __align__(16) float win[2][32][DP_WINP];
    const int lane = threadIdx.x;
    const float INF = __int_as_float(0x7f800000);
    if (lane == 0) *rebuild_flag = 0;
    for (int y = lane; y < h; y += 32) {
        int lo, hi;
        dctc_band_limits(seam, y, h, w, r, &lo, &hi);
        bmin[y] = lo;
        bmax[y] = hi;
    }
    __syncwarp();
    const int xmax4 = ((w + 3) & ~3) - 1;             // last column of the 16-byte chunk that holds column w-1
    // stage row ys (columns that R(ys) and its two neighbours on each side can reach from the range [lo, hi] of a row
    // at most INCR_D above it) into ring slot ys % INCR_D
    auto stage = [&](int ys, int lo, int hi) {
        if (ys < h) {
            const int slot = ys % INCR_D;
            const int s0 = max(0, lo - INCR_D - 2) & ~3, s1 = min(xmax4, (hi + INCR_D + 2) | 3);
            if (lane == 0) xsb[slot] = s0;
            const float* ge = en + (size_t) ys * en_pitch + s0;
            const float* gm = mplane + (size_t) ys * m_pitch + s0;
            float* de = ring + slot * 2 * INCR_SW;
            float* dm = de + INCR_SW;
            const int nch = (s1 - s0 + 1) >> 2;
            for (int c = lane; c < nch; c += 32) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t) __cvta_generic_to_shared(de + 4 * c)), "l"(ge + 4 * c) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t) __cvta_generic_to_shared(dm + 4 * c)), "l"(gm + 4 * c) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int lo = bmin[0], hi = bmax[0];
    if (hi - lo + 1 > INCR_CAP) {
        if (lane == 0) *rebuild_flag = 1;
        return;
    }
    for (int ys = 0; ys < INCR_D; ys++) stage(ys, lo, hi);
    float* prev = buf;
    float* next = buf + INCR_CAP + 4;
    int plo = 0;                                      // prev[i] holds the cell x = plo - 2 + i of row y-1
#ifdef DCTC_INCR_STATS
    int stat_w = 0;
    long long st[5] = {0, 0, 0, 0, 0};
    const long long stat_t0 = clock64();
#endif
    for (int y = 0; y < h; y++) {
        const int width = hi - lo + 1;
#ifdef DCTC_INCR_STATS
        stat_w += width;
#endif
        if (width > INCR_CAP) {                       // uniform
            if (lane == 0) *rebuild_flag = 1;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            return;
        }
#ifdef DCTC_INCR_STATS
        const long long c0 = clock64();
#endif
        asm volatile("cp.async.wait_group %0;" ::"n"(INCR_D - 1) : "memory");   // row y has landed (own copies)
        __syncwarp();                                                            // ... and everybody's
#ifdef DCTC_INCR_STATS
        const long long c1 = clock64();
#endif
        const int slot = y % INCR_D;
        const float* re = ring + slot * 2 * INCR_SW;
        const float* rm = re + INCR_SW;
        const int xs = xsb[slot];
        float* mrow = mplane + (size_t) y * m_pitch;
        int clo = 0x7fffffff, chi = -1;
        if (width <= 32) incr_row<1>(y, lane, lo, hi, w, plo, re, rm, xs, prev, next, mrow, clo, chi);
        else if (width <= 64) incr_row<2>(y, lane, lo, hi, w, plo, re, rm, xs, prev, next, mrow, clo, chi);
        else if (width <= 128) incr_row<4>(y, lane, lo, hi, w, plo, re, rm, xs, prev, next, mrow, clo, chi);
        else if (width <= 256) incr_row<8>(y, lane, lo, hi, w, plo, re, rm, xs, prev, next, mrow, clo, chi);
        else incr_row<16>(y, lane, lo, hi, w, plo, re, rm, xs, prev, next, mrow, clo, chi);
#ifdef DCTC_INCR_STATS
        const long long c2 = clock64();
#endif
        if (lane < 4) {   // two unchanged cells on each side of the range, from the staged (old) map row
            const int xe = lane < 2 ? lo - 2 + lane : hi + lane - 1;
            const int idx = lane < 2 ? lane : width + lane;
            next[idx] = (xe >= 0 && xe < w) ? rm[xe - xs] : INF;
        }
        // Next range.  Finding the changed cells is a warp-wide reduction on the row-to-row critical path, so it is done
        // only every INCR_T rows; in between the range simply grows by one column per side (a superset of R(y+1):
        // cells recomputed without need come out bit-identical, and band(y+1) lies inside band(y) +- 1).
        int nlo = max(lo - 1, 0), nhi = min(hi + 1, w - 1);
        if ((y & (INCR_T - 1)) == INCR_T - 1 && y + 1 < h) {
            clo = __reduce_min_sync(0xffffffffu, clo);
            chi = __reduce_max_sync(0xffffffffu, chi);
            nlo = bmin[y + 1];
            nhi = bmax[y + 1];
            if (chi >= 0) { nlo = min(nlo, clo - 1); nhi = max(nhi, chi + 1); }
            nlo = max(nlo, 0);
            nhi = min(nhi, w - 1);
        }
        plo = lo;
        float* t = prev; prev = next; next = t;
        lo = nlo;
        hi = nhi;
#ifdef DCTC_INCR_STATS
        const long long c3 = clock64();
#endif
        __syncwarp();                                 // hand-over row written, ring slot of row y read by everybody
#ifdef DCTC_INCR_STATS
        const long long c4 = clock64();
#endif
#if DCTC_INCR_ABL == 2
        asm volatile("cp.async.commit_group;" ::: "memory");
#else
        if (hi - lo + 1 <= INCR_CAP) stage(y + INCR_D, lo, hi);                  // refills the slot row y sat in
        else asm volatile("cp.async.commit_group;" ::: "memory");
#endif
#ifdef DCTC_INCR_STATS
        st[0] += c1 - c0; st[1] += c2 - c1; st[2] += c3 - c2; st[3] += c4 - c3; st[4] += clock64() - c4;
#endif
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#ifdef DCTC_INCR_STATS
    if (lane == 0) { atomicAdd(rebuild_flag + 2, stat_w / 16); atomicAdd(rebuild_flag + 3, (int) ((clock64() - stat_t0) >> 10));
        if (w == 1900) printf("walk phases (clk per row): wait %lld, row %lld, ext+range %lld, syncwarp %lld, stage %lld\n", st[0] / h, st[1] / h, st[2] / h, st[3] / h, st[4] / h); }
#endif
    __threadfence();
    __syncwarp();
    // leftmost minimum of the last row (from L2: part of it was just rewritten)
    const float* last = mplane + (size_t) (h - 1) * m_pitch;
    float bv = INF;
    int bi = 0x7fffffff;
    for (int x = lane; x < w; x += 32) {
        const float v = __ldcg(last + x);
        if (v < bv || bi == 0x7fffffff) { bv = v; bi = x; }     // ascending x per lane: strict < keeps the leftmost
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ov < bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    dp_backtrack(mplane, m_pitch, w, h, bi, seam, seam_log, win, lane);
}

// ---- visibility map + seam display ---------------------------------------------------
// liblqr's update_vsmap: the pixel removed from row y by the k-th seam is original column raw[y][s]; its entry of the
// visibility map becomes k (1-based).  raw is compacted over the seam like the image (one CTA per row).
__global__ void __launch_bounds__(256) dctc_vs_update_kernel(int* __restrict__ raw, int* __restrict__ vs, int w0,
                                                             const int* __restrict__ seam, int w_old, int order)
{
    DCTC_PDL_PROLOGUE();
    const int y = blockIdx.x;
    int* rrow = raw + (size_t) y * w0;
    const int s = seam[y];
    if (threadIdx.x == 0) vs[(size_t) y * w0 + rrow[s]] = order;
    for (int base = s; base < w_old - 1; base += 256 * 4) {
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int x = base + k * 256 + threadIdx.x;
            v[k] = x < w_old - 1 ? rrow[x + 1] : 0;
        }
        __syncthreads();   // all loads of this chunk (incl. rrow[s] above) before any store; also orders chunk after chunk
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int x = base + k * 256 + threadIdx.x;
            if (x < w_old - 1) rrow[x] = v[k];
        }
    }
}

__global__ void __launch_bounds__(256) dctc_iota_rows_kernel(int* __restrict__ raw, int w0, int h)
{
    const size_t i = (size_t) blockIdx.x * 256 + threadIdx.x;
    if (i < (size_t) w0 * h) raw[i] = (int) (i % (size_t) w0);
}

// display_carver_seams (src/render.c:204-240): for x < w-1, y < h-1, every removed pixel becomes (0, 255*vis/depth, 0)
__global__ void __launch_bounds__(256) dctc_paint_seams_kernel(uint8_t* __restrict__ img, size_t pitch, int channels, int w, int h,
                                                               const int* __restrict__ vs, int depth)
{
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= w - 1 || y >= h - 1) return;
    const int vis = vs[(size_t) y * w + x];
    if (vis == 0) return;
    uint8_t* p = img + (size_t) y * pitch + (size_t) x * channels;
    p[0] = 0;
    if (channels > 1) p[1] = (uint8_t) (255.0 * ((double) vis) / ((double) depth));
    if (channels > 2) p[2] = 0;
}

// ---- seam enlarging (lqr_carver_resize to a larger size, src/render.c:357-363,377) ----------------------------------
// liblqr's lqr_carver_inflate [from memory: PARITY UNPINNED]: every pixel the first n seams went through is doubled;
// the new pixel sits on its left and holds the integer mean (a + b) / 2 of the pixel and its left neighbour in the
// original row (a copy in column 0).  One CTA per row: block-wide prefix sum of the "seam pixel" flags gives every
// original pixel its output column.
__global__ void __launch_bounds__(256) dctc_inflate_rows_kernel(const uint8_t* __restrict__ orig, size_t pitch, int channels, int w0,
                                                                const int* __restrict__ vs, int n, uint8_t* __restrict__ out, size_t out_pitch)
{
    __shared__ int wsum[8];
    __shared__ int running;
    const int y = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint8_t* src = orig + (size_t) y * pitch;
    uint8_t* dst = out + (size_t) y * out_pitch;
    const int* vrow = vs + (size_t) y * w0;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < w0; base += 256) {
        const int x = base + tid;
        const int v = x < w0 ? vrow[x] : 0;
        const int flag = (v > 0 && v <= n) ? 1 : 0;
        int inc = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        int before = running;
        for (int k = 0; k < wid; k++) before += wsum[k];
        const int pos = x + before + inc - flag;          // output column of the (possibly inserted) first pixel
        if (x < w0) {
            const uint8_t* p = src + (size_t) x * channels;
            uint8_t* q = dst + (size_t) pos * channels;
            if (flag) {
                for (int c = 0; c < channels; c++) q[c] = x > 0 ? (uint8_t) (((int) p[c - channels] + (int) p[c]) / 2) : p[c];
                q += channels;
            }
            for (int c = 0; c < channels; c++) q[c] = p[c];
        }
        __syncthreads();
        if (tid == 255) running = before + inc;
        __syncthreads();
    }
}

void dctc_carver_release(dctc_context* ctx)
{
    if (ctx->c_img) cudaFree(ctx->c_img);
    if (ctx->c_en) cudaFree(ctx->c_en);
    if (ctx->c_m) cudaFree(ctx->c_m);
    if (ctx->c_dir) cudaFree(ctx->c_dir);
    if (ctx->c_seam_log) cudaFree(ctx->c_seam_log);
    if (ctx->c_raw) cudaFree(ctx->c_raw);
    if (ctx->c_vs) cudaFree(ctx->c_vs);
    if (ctx->c_seam) cudaFree(ctx->c_seam);
    if (ctx->c_band) cudaFree(ctx->c_band);
    if (ctx->c_band_vals) cudaFree(ctx->c_band_vals);
    if (ctx->h_mirror) cudaFreeHost(ctx->h_mirror);
    if (ctx->h_band) cudaFreeHost(ctx->h_band);
    ctx->c_img = nullptr; ctx->c_en = nullptr; ctx->c_m = nullptr; ctx->c_dir = nullptr; ctx->c_seam_log = nullptr; ctx->c_seam_log_cap = 0; ctx->c_raw = nullptr; ctx->c_vs = nullptr; ctx->c_vs_depth = 0; ctx->c_vs_w = 0; ctx->c_seam = nullptr; ctx->c_band = nullptr;
    ctx->c_band_vals = nullptr; ctx->h_mirror = nullptr; ctx->h_band = nullptr;
    ctx->c_w0 = ctx->c_w = ctx->c_h = ctx->c_ch = 0;
    ctx->c_m_valid = false;
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

// Full energy map of the session's current image.  A carver session keeps one arithmetic for the whole seam loop: the
// per-seam band updates run in the FP32 tile kernel, so full maps use the bit-identical FP32 march kernel rather than
// the tensor-core kernel (liblqr: update_emap must reproduce what build_emap would give on the carved image).
static int carver_full_energy(dctc_context* ctx)
{
    DctcK1Args a;
    carver_args(ctx, a);
    const int saved_kernel = ctx->kernel;
    if (ctx->kernel == DCTC_KERNEL_AUTO || ctx->kernel == DCTC_KERNEL_TC_SPLIT || ctx->kernel == DCTC_KERNEL_FP32_STREAM) ctx->kernel = DCTC_KERNEL_FP32_MARCH;
    const int rc = dctc_run_k1(ctx, a, 1, ctx->stream);
    ctx->kernel = saved_kernel;
    return rc;
}

// dctc_set_params with a session loaded (liblqr: lqr_carver_set_energy_function invalidates the energy map): the
// resident map is rebuilt with the new operator and the cumulative plane of the seam DP is dropped.
int dctc_carver_params_changed(dctc_context* ctx)
{
    if (!ctx->c_img) return DCTC_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    ctx->mirror_valid = false;
    ctx->c_m_valid = false;
    const int rc = carver_full_energy(ctx);
    if (rc) return rc;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return DCTC_OK;
}

static int carver_load_impl(dctc_context* ctx, const uint8_t* img, int w, int h, int channels, size_t pitch)
{
    // img may be host or device memory (cudaMemcpyDefault: the enlarge path reloads the session from a device buffer)
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
    CK(ctx, cudaMemcpy2DAsync(ctx->c_img, ctx->c_pitch, img, pitch, (size_t) w * channels, h, cudaMemcpyDefault,
                              ctx->stream));
    int rc = carver_full_energy(ctx);
    if (rc) return rc;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return DCTC_OK;
}

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
    // a half-built session (an allocation or the upload failed) must not look loaded to the next call
    const int rc = carver_load_impl(ctx, img, w, h, channels, pitch);
    if (rc) dctc_carver_release(ctx);
    return rc;
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
        // connected seams only (delta_x = 1, src/render.c:313): the band width 4*(b/2) and the band tiles rely on it
        if (y > 0 && (seam_x[y] - seam_x[y - 1] > 1 || seam_x[y - 1] - seam_x[y] > 1)) return DCTC_ERR_INVALID;
        h_seam[y] = seam_x[y];
    }
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMemcpyAsync(ctx->c_seam, h_seam, sizeof(int) * h, cudaMemcpyHostToDevice, ctx->stream));
    dctc_carve_rows_kernel<256, 4><<<h, 256, 0, ctx->stream>>>(ctx->c_img, ctx->c_pitch, ctx->c_ch, ctx->c_en,
                                                               ctx->c_en_pitch, ctx->c_seam, w_old, nullptr, 0);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    ctx->c_w = w_old - 1;
    ctx->mirror_valid = false;
    ctx->c_m_valid = false;     // the host picked this seam: the device's cumulative plane no longer matches
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
    if (ctx->c_w > DP_CL * DP_MAXW * (512 - 2 * DP_R)) return DCTC_ERR_UNSUPPORTED;   // wider than the cluster's strips: use the host seam loop
    CK(ctx, cudaSetDevice(ctx->device));
    if (ctx->c_dump_vmaps && !ctx->c_raw) {
        if (ctx->c_w != ctx->c_w0) return DCTC_ERR_STATE;   // the map must be requested before the first seam is removed
        const size_t n = (size_t) ctx->c_w0 * h;
        CK(ctx, cudaMalloc((void**) &ctx->c_raw, sizeof(int) * n));
        CK(ctx, cudaMalloc((void**) &ctx->c_vs, sizeof(int) * n));
        CK(ctx, cudaMemsetAsync(ctx->c_vs, 0, sizeof(int) * n, ctx->stream));
        dctc_iota_rows_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->c_raw, ctx->c_w0, h);
        CK(ctx, cudaGetLastError());
        ctx->launches++;
        ctx->c_vs_depth = 0;
        ctx->c_vs_w = ctx->c_w0;
    }
    // cumulative-map plane, rows padded so that the back-track windows (DP_WIN floats) stay inside
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
    // strip geometry: 128*P columns per warp, 128*P - 2*DP_R published; the strips are spread over a cluster of DP_CL CTAs
    const int wcur = ctx->c_w;
    const int cap = DP_CL * DP_MAXW;
    int P = wcur <= cap * (128 - 2 * DP_R) ? 1 : wcur <= cap * (256 - 2 * DP_R) ? 2 : 4;
    if (const char* ep = getenv("DCTC_DP_P")) {   // timing experiments: wider strips than the width needs
        const int want = atoi(ep);
        if ((want == 2 || want == 4) && want > P) P = want;
    }
    const int wout = 128 * P - 2 * DP_R;
    const int strips = (wcur + wout - 1) / wout;
    const int wpc = (strips + DP_CL - 1) / DP_CL;
    // energies staged by bulk async copies when at least two stages (DP_SR rows each) per warp fit into shared memory
    // (the kernel also has ~38 KB of static shared memory: back-track windows, reduction scratch, mbarriers)
    const size_t xrow_smem = sizeof(float) * 2 * (size_t) wpc * wout;
    int nst = 0, sr = 0;
    if (!getenv("DCTC_DP_NO_TMA")) {
        const size_t row_smem = sizeof(float) * ((size_t) wpc * 128 * P);      // one staged row of every warp
        const size_t budget = 180 * 1024 - xrow_smem;
        if (2 * 32 * row_smem <= budget) { sr = 32; nst = 2; }                 // whole 32-row blocks, two in flight
        else {
            sr = 16;
            for (nst = DP_NST_MAX; nst >= 2 && (size_t) nst * 16 * row_smem > budget; nst--) {}
            if (nst < 2) { sr = 0; nst = 0; }
        }
    }
    const size_t dp_smem = xrow_smem + sizeof(float) * ((size_t) nst * sr * wpc * 128 * P);
    using dp_fn = void (*)(const float*, size_t, int, int, float*, size_t, int*, int*, int*, int, const CUtensorMap, int, int*);
    static const dp_fn table[3][3] = {{dctc_seam_dp_kernel<1, 0>, dctc_seam_dp_kernel<1, 16>, dctc_seam_dp_kernel<1, 32>},
                                      {dctc_seam_dp_kernel<2, 0>, dctc_seam_dp_kernel<2, 16>, dctc_seam_dp_kernel<2, 32>},
                                      {dctc_seam_dp_kernel<4, 0>, dctc_seam_dp_kernel<4, 16>, dctc_seam_dp_kernel<4, 32>}};
    const dp_fn dp = table[P == 1 ? 0 : P == 2 ? 1 : 2][sr / 16];
    // tensor map of the energy plane for the staged variant (strips of up to 256 columns: the TMA box limit)
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int use_tmap = 0;
    if (sr > 0 && 128 * P <= 256 && !getenv("DCTC_DP_NO_TENSORMAP")) {
        typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) == cudaSuccess && fp &&
            qres == cudaDriverEntryPointSuccess) {
            const cuuint64_t gdim[2] = {(cuuint64_t) ctx->c_en_pitch, (cuuint64_t) h};
            const cuuint64_t gstr[1] = {(cuuint64_t) ctx->c_en_pitch * sizeof(float)};
            const cuuint32_t box[2] = {(cuuint32_t) (128 * P), (cuuint32_t) sr};
            const cuuint32_t estr[2] = {1, 1};
            if (((encode_fn) fp)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ctx->c_en, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                use_tmap = 1;
        } else {
            (void) cudaGetLastError();
        }
    }
    // always opt in: the kernel's static shared memory (back-track windows) plus the dynamic part can exceed 48 KB
    CK(ctx, cudaFuncSetAttribute(dp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (dp_smem > 1024 ? dp_smem : 1024)));
    // incremental update (update_mmap): one warp; needs the band table of h rows twice in shared memory
    const size_t incr_smem = sizeof(float) * (INCR_D * 2 * INCR_SW + 2 * (INCR_CAP + 4)) + sizeof(int) * (INCR_D + 2 * (size_t) h);
    const bool incr_ok = ctx->c_incremental && incr_smem <= 160 * 1024;
    if (incr_ok)   // dynamic + static shared memory can exceed the 48 KB default
        CK(ctx, cudaFuncSetAttribute(dctc_seam_incr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) incr_smem));
    if (!ctx->c_band) {
        CK(ctx, cudaMalloc((void**) &ctx->c_band, sizeof(int) * 8));
        CK(ctx, cudaMemsetAsync(ctx->c_band, 0, sizeof(int) * 8, ctx->stream));
    }
    // parallel back-track (jump + trace kernels): needs the jump plane and a cone of jump entries in shared memory that
    // grows with the square of the number of 32-row blocks; very tall images keep the serial walk of the DP kernel
    const int nb = (h - 1 + JB - 1) / JB;
    const size_t cone_smem = nb > 0 ? (size_t) 32 * (nb - 1) * (nb - 1) + 16 * (size_t) (nb - 1) + 16 : 16;   // rows j < nb-1: 16 (4 j + 3) bytes each
    const size_t j_pitch = (m_pitch + 15) & ~(size_t) 15;    // jump rows start 16-byte aligned (staged in 16-byte chunks)
    const bool par_bt = nb > 0 && cone_smem <= 200 * 1024 && !getenv("DCTC_SERIAL_BACKTRACK");
    int* const xlast = par_bt ? ctx->c_band + 4 : nullptr;
    if (par_bt) {
        if (!ctx->c_dir) CK(ctx, cudaMalloc((void**) &ctx->c_dir, (size_t) nb * j_pitch));
        if (cone_smem > 40 * 1024)
            CK(ctx, cudaFuncSetAttribute(dctc_seam_trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) cone_smem));
    }
    const bool pdl = !getenv("DCTC_NO_PDL");
    for (int s = 0; s < n_seams; s++) {
        const int w_old = ctx->c_w;
        int* log_s = ctx->c_seam_log + (size_t) s * h;
        if (ctx->c_m_valid && incr_ok) {
            // update_mmap + build_vpath: walk the changed cells only; the full rebuild below runs only if the walk gave up
            dctc_seam_incr_kernel<<<1, 32, incr_smem, ctx->stream>>>(ctx->c_en, ctx->c_en_pitch, w_old, h, ctx->c_m, m_pitch, r,
                                                                     ctx->c_seam, log_s, ctx->c_band);
            dp<<<DP_CL, wpc * 32, dp_smem, ctx->stream>>>(ctx->c_en, ctx->c_en_pitch, w_old, h, ctx->c_m, m_pitch, ctx->c_seam, log_s, ctx->c_band, nst, tmap, use_tmap, nullptr);
            ctx->launches++;
        } else {
            // build_mmap, then build_vpath: jump maps of all 32-row blocks in parallel, one trace CTA per block
            // (programmatic dependent launches: each kernel of the loop is scheduled while its predecessor drains and
            // blocks in DCTC_PDL_PROLOGUE until that one's results are visible)
            CK(ctx, dctc_launch_pdl(dp, dim3(DP_CL), dim3(wpc * 32), dp_smem, ctx->stream, pdl, ctx->c_en, ctx->c_en_pitch, w_old, h, ctx->c_m, m_pitch,
                                    ctx->c_seam, log_s, (int*) nullptr, nst, tmap, use_tmap, xlast));
            if (par_bt) {
                const int js = (w_old + 63) / 64;
                CK(ctx, dctc_launch_pdl(dctc_seam_jump_kernel, dim3((js + 3) / 4, nb), dim3(128), 0, ctx->stream, pdl, ctx->c_m, m_pitch, w_old, h,
                                        ctx->c_dir, j_pitch, js));
                CK(ctx, dctc_launch_pdl(dctc_seam_trace_kernel, dim3(nb), dim3(TR_THREADS), cone_smem, ctx->stream, pdl, ctx->c_m, m_pitch, w_old, h,
                                        ctx->c_dir, j_pitch, xlast, ctx->c_seam, log_s));
                ctx->launches += 2;
            }
        }
#ifdef DCTC_SYNC_DEBUG
        { cudaError_t e_ = cudaStreamSynchronize(ctx->stream); if (e_ != cudaSuccess) { printf("seam %d: dp kernel failed: %s (w %d)\n", s, cudaGetErrorString(e_), w_old); return dctc_fail_cuda(ctx, e_); } }
#endif
        if (ctx->c_dump_vmaps) {   // update_vsmap
            CK(ctx, dctc_launch_pdl(dctc_vs_update_kernel, dim3(h), dim3(256), 0, ctx->stream, pdl, ctx->c_raw, ctx->c_vs, ctx->c_vs_w, ctx->c_seam, w_old,
                                    ++ctx->c_vs_depth));
            ctx->launches++;
        }
        // carve: compact image and energy rows over the seam
        CK(ctx, dctc_launch_pdl(dctc_carve_rows_kernel<256, 4>, dim3(h), dim3(256), 0, ctx->stream, pdl, ctx->c_img, ctx->c_pitch, ctx->c_ch, ctx->c_en,
                                ctx->c_en_pitch, ctx->c_seam, w_old, incr_ok ? ctx->c_m : (float*) nullptr, m_pitch));
        CK(ctx, cudaGetLastError());
        ctx->launches += 2;
        ctx->c_w = w_old - 1;
        ctx->c_m_valid = incr_ok;   // the plane now holds the map of the image before this removal, moved with the pixels
        // update_emap: K1 in band mode around the removed seam
        DctcK1Args a;
        carver_args(ctx, a);
        a.seam = ctx->c_seam; a.band_r = r; a.band_vals = ctx->c_band_vals; a.band_stride = bs;
        a.pdl = pdl ? 1 : 0;
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

// lqr_carver_resize to a LARGER width (seams_number > 0, src/render.c:357-363): the n seams a shrink would remove,
// then the inflate kernel; the session continues on the enlarged image (energy map rebuilt), the visibility map of the
// start frame stays readable through dctc_carver_vmap.
int dctc_carver_enlarge_width(dctc_context* ctx, int n_seams, int* seams_out)
{
    if (!ctx || n_seams < 0) return DCTC_ERR_INVALID;
    if (!ctx->c_img) return DCTC_ERR_STATE;
    if (n_seams == 0) return DCTC_OK;
    if (n_seams >= ctx->c_w || ctx->c_w != ctx->c_w0) return DCTC_ERR_STATE;   // enlarge starts from a freshly loaded frame
    const int w0 = ctx->c_w, h = ctx->c_h, ch = ctx->c_ch, w1 = w0 + n_seams;
    CK(ctx, cudaSetDevice(ctx->device));
    uint8_t* d_orig = nullptr;
    uint8_t* d_new = nullptr;
    const size_t p0 = ctx->c_pitch, p1 = ((size_t) w1 * ch + 15) & ~(size_t) 15;
    CK(ctx, cudaMalloc((void**) &d_orig, p0 * h));
    cudaError_t e = cudaMalloc((void**) &d_new, p1 * h);
    if (e != cudaSuccess) { cudaFree(d_orig); return dctc_fail_cuda(ctx, e); }
    e = cudaMemcpyAsync(d_orig, ctx->c_img, p0 * h, cudaMemcpyDeviceToDevice, ctx->stream);
    int rc = e == cudaSuccess ? DCTC_OK : dctc_fail_cuda(ctx, e);
    const bool dump_saved = ctx->c_dump_vmaps;
    if (rc == DCTC_OK) {
        ctx->c_dump_vmaps = true;                      // the seams' order per original pixel drives the pixel synthesis
        rc = dctc_carver_resize_width(ctx, n_seams, seams_out);
        ctx->c_dump_vmaps = dump_saved;
    }
    if (rc == DCTC_OK) {
        dctc_inflate_rows_kernel<<<h, 256, 0, ctx->stream>>>(d_orig, p0, ch, w0, ctx->c_vs, n_seams, d_new, p1);
        e = cudaGetLastError();
        ctx->launches++;
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = dctc_fail_cuda(ctx, e);
    }
    if (rc == DCTC_OK) {
        // continue the session on the enlarged image; keep the visibility map of the start frame
        int* vs = ctx->c_vs;
        const int depth = ctx->c_vs_depth, vs_w = ctx->c_vs_w;
        ctx->c_vs = nullptr;
        dctc_carver_release(ctx);
        rc = carver_load_impl(ctx, d_new, w1, h, ch, p1);
        if (rc != DCTC_OK) dctc_carver_release(ctx);
        if (rc == DCTC_OK) { ctx->c_vs = vs; ctx->c_vs_depth = depth; ctx->c_vs_w = vs_w; }
        else cudaFree(vs);
    }
    cudaFree(d_orig);
    cudaFree(d_new);
    return rc;
}

int dctc_carver_set_incremental(dctc_context* ctx, int on)
{
    if (!ctx) return DCTC_ERR_INVALID;
    ctx->c_incremental = on != 0;
    return DCTC_OK;
}

int dctc_carver_rebuild_count(dctc_context* ctx)
{
    if (!ctx || !ctx->c_img) return -1;
    if (!ctx->c_band) return 0;
    int v[4] = {0, 0, 0, 0};
    if (cudaSetDevice(ctx->device) != cudaSuccess) return -1;
    if (cudaMemcpyAsync(v, ctx->c_band, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
#ifdef DCTC_INCR_STATS
    fprintf(stderr, "incremental map: rebuilds %d, sum of range widths / 16 = %d, walk kclk = %d\n", v[1], v[2], v[3]);
#endif
    return v[1];
}

int dctc_carver_set_dump_vmaps(dctc_context* ctx, int on)
{
    if (!ctx) return DCTC_ERR_INVALID;
    ctx->c_dump_vmaps = on != 0;
    return DCTC_OK;
}

int dctc_carver_vmap(dctc_context* ctx, int* vmap_out, int* depth_out)
{
    if (!ctx) return DCTC_ERR_INVALID;
    if (!ctx->c_img || !ctx->c_vs) return DCTC_ERR_STATE;
    CK(ctx, cudaSetDevice(ctx->device));
    if (vmap_out)
        CK(ctx, cudaMemcpyAsync(vmap_out, ctx->c_vs, sizeof(int) * (size_t) ctx->c_vs_w * ctx->c_h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (depth_out) *depth_out = ctx->c_vs_depth;
    return DCTC_OK;
}

int dctc_carver_paint_seams(dctc_context* ctx, uint8_t* img, int channels, size_t pitch)
{
    if (!ctx || !img || channels < 1 || channels > 4) return DCTC_ERR_INVALID;
    if (!ctx->c_img || !ctx->c_vs || ctx->c_vs_depth <= 0) return DCTC_ERR_STATE;
    const int w = ctx->c_vs_w, h = ctx->c_h;
    if (pitch < (size_t) w * channels) return DCTC_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    uint8_t* d = nullptr;
    const size_t dp = ((size_t) w * channels + 15) & ~(size_t) 15;
    CK(ctx, cudaMalloc((void**) &d, dp * h));
    cudaError_t e = cudaMemcpy2DAsync(d, dp, img, pitch, (size_t) w * channels, h, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        dctc_paint_seams_kernel<<<dim3((w + 255) / 256, h), 256, 0, ctx->stream>>>(d, dp, channels, w, h, ctx->c_vs, ctx->c_vs_depth);
        e = cudaGetLastError();
        ctx->launches++;
    }
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(img, pitch, d, dp, (size_t) w * channels, h, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t es = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e == cudaSuccess) e = es;
    if (e != cudaSuccess) return dctc_fail_cuda(ctx, e);
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
