// tc_probe2b.cu — 2 CTAs per SM variant of tc_probe2.cu (256 TMEM columns each): do MMAs of co-resident CTAs interleave for free? tcgen05.mma issue-rate probe with the MMAs issued from a converged warp behind elect.sync
// (tc_probe.cu issued them under `if (tid == 0)`, which ptxas turns into an ELECT/BRA.U.ANY loop per MMA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe2 tools/tc_probe2.cu && ./tc_probe2
// Reports SM cycles per MMA for TS (A in TMEM) and SS (A in smem) at M=128, K=16, N in {64,128,256}, and the
// latency of one tcgen05.st.x8 + tcgen05.ld.x64 round trip.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W;\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t) ((addr & 0x3FFFF) >> 4);
    d |= (uint64_t) ((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t) ((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t) 1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24); }
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

struct alignas(128) Smem {
    __half b[256 * 16];
    __half a[128 * 16];
    uint64_t bar;
    uint32_t tmem_base;
};

// mode 0: TS, 1: SS.  pattern 0: every MMA overwrites one D tile; 1: groups of 3 (first overwrites, two accumulate),
// rotating over 4 D tiles (the energy kernel's pattern).
template <int N>
__global__ void __launch_bounds__(160) probe(int mode, int pattern, int iters, long long* cycles)
{
    __shared__ Smem s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) { mbar_init(smem_u32(&s.bar), 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    for (int i = tid; i < 256 * 16; i += 160) s.b[i] = __float2half(0.001f * (i % 97));
    for (int i = tid; i < 128 * 16; i += 160) s.a[i] = __float2half(0.002f * (i % 89));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s.tmem_base;
    if (warp < 4) {
        uint32_t r[8];
        for (int j = 0; j < 8; j++) r[j] = 0x3C003C00u;
        const uint32_t lane_base = (uint32_t) (warp * 32) << 16;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem + lane_base + 0), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (warp == 4) {
        const uint64_t bdesc = make_smem_desc(smem_u32(s.b), 128, 256);
        const uint64_t adesc = make_smem_desc(smem_u32(s.a), 2048, 128);
        const uint32_t idesc = make_idesc(128, N);
        long long t0 = clock64();
        if (elect_one()) {
            if (pattern == 0) {
                for (int it = 0; it < iters; it++) {
                    if (mode == 0) mma_ts(tmem + 128, tmem, bdesc, idesc, 0);
                    else mma_ss(tmem + 128, adesc, bdesc, idesc, 0);
                }
            } else {
                for (int it = 0; it < iters; it += 3) {
                    const uint32_t d = tmem + 128 + (N <= 64 ? (uint32_t) ((it / 3) & 1) * 64u : 0u);
                    if (mode == 0) { mma_ts(d, tmem, bdesc, idesc, 0); mma_ts(d, tmem + 8, bdesc, idesc, 1); mma_ts(d, tmem, bdesc, idesc, 1); }
                    else { mma_ss(d, adesc, bdesc, idesc, 0); mma_ss(d, adesc, bdesc, idesc, 1); mma_ss(d, adesc, bdesc, idesc, 1); }
                }
            }
            mma_commit(smem_u32(&s.bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&s.bar), 0);
        long long t1 = clock64();
        if ((tid & 31) == 0) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// latency of tcgen05.st x8 + wait and tcgen05.ld x64 + wait, per warp, measured by lane 0 of warp 0
__global__ void __launch_bounds__(128) ldst_probe(long long* out)
{
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base + ((uint32_t) (warp * 32) << 16);
    uint32_t r[8];
    for (int j = 0; j < 8; j++) r[j] = tid + j;
    long long t0 = clock64();
    for (int it = 0; it < 64; it++) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem + (it & 7) * 8), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    long long t1 = clock64();
    uint32_t v[64];
    uint32_t acc = 0;
    for (int it = 0; it < 64; it++) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
                     : "r"(tmem + (it & 1) * 64));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 64; j++) acc ^= v[j];
    }
    long long t2 = clock64();
    if (tid == 0) { out[0] = (t1 - t0) / 64; out[1] = (t2 - t1) / 64; out[2] = acc; }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
}

template <int N>
void run(long long* dC)
{
    for (int mode = 0; mode < 2; mode++)
        for (int pattern = 0; pattern < 2; pattern++)
            for (int grid : {148, 296}) {
                const int iters = 3072;
                probe<N><<<grid, 160>>>(mode, pattern, iters, dC);
                CHECK(cudaDeviceSynchronize());
                long long c[296];
                CHECK(cudaMemcpy(c, dC, sizeof(c), cudaMemcpyDeviceToHost));
                long long mx = 0;
                for (int i = 0; i < grid; i++) if (c[i] > mx) mx = c[i];
                printf("  %s M128 N%-3d K16 %s, %3d CTAs: %6.1f cycles/MMA\n", mode == 0 ? "TS" : "SS", N, pattern ? "3-accumulate groups over 4 tiles" : "overwrite one tile              ", grid, (double) mx / iters);
            }
}

int main()
{
    setvbuf(stdout, nullptr, _IONBF, 0);
    long long* dC;
    CHECK(cudaMalloc(&dC, 296 * 8));
    run<64>(dC);


    ldst_probe<<<1, 128>>>(dC);
    CHECK(cudaDeviceSynchronize());
    long long c[3];
    CHECK(cudaMemcpy(c, dC, sizeof(c), cudaMemcpyDeviceToHost));
    printf("  tcgen05.st.x8+wait: %lld cycles, tcgen05.ld.x64+wait(+64 xor): %lld cycles\n", c[0], c[1]);
    printf("TC_PROBE2 DONE\n");
    return 0;
}
