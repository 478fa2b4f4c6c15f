// tc_probe3.cu — tcgen05.ld throughput by shape, all four lane quarters busy (4 warps) on every SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe3 tools/tc_probe3.cu && ./tc_probe3
// The K1-TC consumers read 64 FP32 accumulators per pixel (256 B/px) out of TMEM; this measures how many bytes per
// clock an SM can move TMEM -> registers with the 32x32b shape the kernel uses and with the wider 16x256b shape.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

#define R8(v, o) "=r"(v[o]), "=r"(v[o+1]), "=r"(v[o+2]), "=r"(v[o+3]), "=r"(v[o+4]), "=r"(v[o+5]), "=r"(v[o+6]), "=r"(v[o+7])
#define P8(o) "%" #o
#define LD32(shape, taddr, v)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned." shape ".b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
                 : R8(v, 0), R8(v, 8), R8(v, 16), R8(v, 24) : "r"(taddr))
#define LD64(shape, taddr, v)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned." shape ".b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];" \
                 : R8(v, 0), R8(v, 8), R8(v, 16), R8(v, 24), R8(v, 32), R8(v, 40), R8(v, 48), R8(v, 56) : "r"(taddr))

// mode 0: 32x32b.x64 (one instruction = 32 lanes x 64 columns = 8 KB)
// mode 1: 32x32b.x32 twice
// mode 2: 16x256b.x8 twice (16 lanes x 64 columns each = 4 KB)
// mode 3: 16x128b.x16 twice? (16 lanes x 128b x16 = 16 lanes x 64 columns) -- 32 registers
template <int MODE>
__global__ void __launch_bounds__(128) probe(int iters, long long* cycles, uint32_t* sink)
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
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const uint32_t col = (uint32_t) (it & 3) * 64u;
        if (MODE == 0) {
            uint32_t v[64];
            LD64("32x32b.x64", tmem + col, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= v[0] ^ v[63];
        } else if (MODE == 1) {
            uint32_t v[32], u[32];
            LD32("32x32b.x32", tmem + col, v);
            LD32("32x32b.x32", tmem + col + 32, u);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= v[0] ^ u[31];
        } else if (MODE == 2) {
            uint32_t v[32], u[32];
            LD32("16x256b.x8", tmem + col, v);
            LD32("16x256b.x8", tmem + col + (16u << 16), u);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= v[0] ^ u[31];
        } else {
            uint32_t v[32], u[32];
            LD32("16x128b.x16", tmem + col, v);
            LD32("16x128b.x16", tmem + col + (16u << 16), u);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= v[0] ^ u[31];
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * 128 + tid] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
}

template <int MODE>
void run(const char* name, long long* dC, uint32_t* dS)
{
    const int iters = 2048;
    for (int grid : {1, 148, 296}) {
        probe<MODE><<<grid, 128>>>(iters, dC, dS);
        CHECK(cudaDeviceSynchronize());
        long long c[296];
        CHECK(cudaMemcpy(c, dC, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (int i = 0; i < grid; i++) if (c[i] > mx) mx = c[i];
        // per iteration every warp moves 32 lanes x 64 columns x 4 B = 8 KB; 4 warps per CTA
        const double ctas_per_sm = grid <= 148 ? 1.0 : 2.0;
        printf("  %-28s %3d CTAs: %7.1f clk per 8 KB per warp, %6.1f B/clk per SM (4 warps x %.0f CTAs)\n", name, grid, (double) mx / iters,
               4.0 * ctas_per_sm * 8192.0 * iters / (double) mx, ctas_per_sm);
    }
}

int main()
{
    setvbuf(stdout, nullptr, _IONBF, 0);
    long long* dC; uint32_t* dS;
    CHECK(cudaMalloc(&dC, 296 * 8));
    CHECK(cudaMalloc(&dS, 296 * 128 * 4));
    run<0>("32x32b.x64", dC, dS);
    run<1>("32x32b.x32 x2", dC, dS);
    run<2>("16x256b.x8 x2", dC, dS);
    run<3>("16x128b.x16 x2", dC, dS);
    printf("TC_PROBE3 DONE\n");
    return 0;
}
