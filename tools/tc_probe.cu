// tc_probe.cu — stand-alone probe of the tcgen05 building blocks the tensor-core energy kernel relies on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tools/tc_probe.cu && ./tc_probe
// Checks (against a CPU matmul) one M128 x N64 x K16 fp16 MMA with FP32 accumulation in both operand modes
//   TS: A written to TMEM with tcgen05.st (row m in lane m, two fp16 per 32-bit column), B from shared memory
//   SS: A and B from shared memory in the no-swizzle K-major canonical layout
// and measures cycles per MMA when many are issued back to back (is N=64 shared-memory-bound in SS mode?).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    for (long long spin = 0; !mbar_try_wait(bar, parity); spin++)
        if (spin > 4000000LL) { printf("mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t) ((addr & 0x3FFFF) >> 4);
    d |= (uint64_t) ((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t) ((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t) 1 << 46;  // descriptor version for sm_100
    return d;                 // layout_type = 0 (no swizzle), base_offset = 0
}

// kind::f16, D=F32, A=B=F16, both K-major, M=128, N
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

#define TMEM_LD_X32(taddr, r)                                                                                          \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
                 : "r"(taddr))

struct alignas(128) Smem {
    __half b[256 * 16];      // canonical K-major no-swizzle: [n/8][k/8][n%8][k%8] (rows >= 64 only used by the timing runs)
    __half a[128 * 16];      // canonical K-major no-swizzle: [k/8][m/8][m%8][k%8]
    uint64_t bar;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(128) probe_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D, int mode, int iters, long long* cycles, int N, int same_tile, int issuers = 1, int M = 128)
{
    __shared__ Smem s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        mbar_init(smem_u32(&s.bar), issuers);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    for (int i = tid; i < 256 * 16; i += 128) {
        const int n = i / 16, k = i % 16;
        s.b[(n / 8) * 128 + (k / 8) * 64 + (n % 8) * 8 + (k % 8)] = B[i % (64 * 16)];
    }
    for (int i = tid; i < 128 * 16; i += 128) {
        const int m = i / 16, k = i % 16;
        s.a[(k / 8) * 1024 + (m / 8) * 64 + (m % 8) * 8 + (k % 8)] = A[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s.tmem_base;
    const uint32_t lane_base = (uint32_t) (warp * 32) << 16;
    // A row of this thread -> TMEM columns [64, 72): 8 x 32-bit = 16 fp16
    {
        uint32_t r[8];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(A + tid * 16);
        for (int j = 0; j < 8; j++) r[j] = src[j];
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem + lane_base + 64), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint64_t bdesc = make_smem_desc(smem_u32(s.b), 128, 256);
    const uint64_t adesc = make_smem_desc(smem_u32(s.a), 2048, 128);
    const uint32_t idesc = make_idesc(M, N);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        t0 = clock64();
        for (int it = 0; it < iters; it++) {
            // rotate over 4 accumulator tiles so consecutive MMAs are independent; tile 0 holds the checked result
            const uint32_t dcol = (iters == 1) ? 0u : (same_tile ? 256u : 256u * (uint32_t) (it & 1));
            const uint32_t acc = same_tile && it > 0;
            if (mode == 0) mma_ts(tmem + dcol, tmem + 64, bdesc, idesc, acc);
            else mma_ss(tmem + dcol, adesc, bdesc, idesc, acc);
        }
        mma_commit(smem_u32(&s.bar));
    }
    mbar_wait(smem_u32(&s.bar), 0);
    if (tid == 0) { t1 = clock64(); if (cycles) cycles[blockIdx.x] = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t v[32];
    for (int half = 0; half < 2; half++) {
        TMEM_LD_X32(tmem + lane_base + half * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (blockIdx.x == 0)
            for (int j = 0; j < 32; j++) D[tid * 64 + half * 32 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main()
{
    setvbuf(stdout, nullptr, _IONBF, 0);
    std::vector<__half> A(128 * 16), B(64 * 16);
    std::vector<float> Af(128 * 16), Bf(64 * 16), ref(128 * 64), out(128 * 64);
    srand(1);
    for (int i = 0; i < 128 * 16; i++) { float v = (rand() % 2001 - 1000) / 64.0f; A[i] = __float2half(v); Af[i] = __half2float(A[i]); }
    for (int i = 0; i < 64 * 16; i++) { float v = (rand() % 2001 - 1000) / 2048.0f; B[i] = __float2half(v); Bf[i] = __half2float(B[i]); }
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < 64; n++) {
            double s = 0;
            for (int k = 0; k < 16; k++) s += (double) Af[m * 16 + k] * Bf[n * 16 + k];
            ref[m * 64 + n] = (float) s;
        }
    __half *dA, *dB; float* dD; long long* dC;
    CHECK(cudaMalloc(&dA, A.size() * 2)); CHECK(cudaMalloc(&dB, B.size() * 2)); CHECK(cudaMalloc(&dD, out.size() * 4)); CHECK(cudaMalloc(&dC, 148 * 8));
    CHECK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
    int bad_total = 0;
    for (int mode = 0; mode < 2; mode++) {
        CHECK(cudaMemset(dD, 0, out.size() * 4));
        probe_kernel<<<1, 128>>>(dA, dB, dD, mode, 1, nullptr, 64, 0);
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0; int bad = 0;
        for (int i = 0; i < 128 * 64; i++) { double e = fabs((double) out[i] - ref[i]); if (e > maxerr) maxerr = e; if (e > 1e-3) bad++; }
        printf("mode %s: max abs err %.3e, mismatches %d / 8192  (D[0][0]=%f ref %f, D[5][9]=%f ref %f)\n", mode == 0 ? "TS" : "SS", maxerr, bad, out[0], ref[0], out[5 * 64 + 9], ref[5 * 64 + 9]);
        bad_total += bad;
        for (int N : {64, 128, 256})
            for (int same : {0, 1})
                for (int grid : {148}) {
                    const int iters = 1024;
                    probe_kernel<<<grid, 128>>>(dA, dB, dD, mode, iters, dC, N, same);
                    CHECK(cudaDeviceSynchronize());
                    long long c[148];
                    CHECK(cudaMemcpy(c, dC, sizeof(c), cudaMemcpyDeviceToHost));
                    long long mx = 0; for (int i = 0; i < grid; i++) if (c[i] > mx) mx = c[i];
                    printf("  %s %d MMAs M128 N%d K16, %s, %d CTAs: %.1f cycles/MMA\n", mode == 0 ? "TS" : "SS", iters, N, same ? "accumulate into one tile" : "alternate two tiles", grid, (double) mx / iters);
                }
    }
    for (int issuers : {1, 2})
        for (int N : {64}) {
            probe_kernel<<<148, 128>>>(dA, dB, dD, 0, 1024, dC, N, 1, issuers, 128);
            CHECK(cudaDeviceSynchronize());
            long long c[148];
            CHECK(cudaMemcpy(c, dC, sizeof(c), cudaMemcpyDeviceToHost));
            long long mx = 0; for (int i = 0; i < 148; i++) if (c[i] > mx) mx = c[i];
            printf("  TS M128 N%d: %d issuing warps x 1024 MMAs: %.1f cycles per MMA (aggregate)\n", N, issuers, (double) mx / (1024.0 * issuers));
        }
    printf("TC_PROBE %s\n", bad_total == 0 ? "PASS" : "FAIL");
    return bad_total == 0 ? 0 : 1;
}
