// fp32x2_probe.cu - throughput of scalar FFMA vs packed FFMA2 vs a 2:1 mix on B200 (per-SM lanes/clk).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int MODE>  // 0: 12 FFMA chains, 1: 12 FFMA2 chains, 2: 8 FFMA2 + 8 FFMA chains, 3: 8 FFMA2 + 4 FFMA
__global__ void __launch_bounds__(256) k(float* out, int iters, long long* cyc)
{
    float a[16]; float2 b[12];
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 0.001f + i;
    for (int i = 0; i < 12; i++) b[i] = make_float2(threadIdx.x * 0.002f + i, i * 0.5f);
    const float c = 0.999f, d = 0.001f;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 12; i++) a[i] = fmaf(a[i], c, d);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 12; i++) b[i] = __ffma2_rn(b[i], make_float2(c, c), make_float2(d, d));
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) { b[i] = __ffma2_rn(b[i], make_float2(c, c), make_float2(d, d)); a[i] = fmaf(a[i], c, d); }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) { b[2 * i] = __ffma2_rn(b[2 * i], make_float2(c, c), make_float2(d, d)); b[2 * i + 1] = __ffma2_rn(b[2 * i + 1], make_float2(c, c), make_float2(d, d)); a[i] = fmaf(a[i], c, d); }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 16; i++) s += a[i];
    for (int i = 0; i < 12; i++) s += b[i].x + b[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double flops_per_iter_per_thread)
{
    float* out; long long* cyc;
    const int blocks = 148 * 4, iters = 4096;
    CHECK(cudaMalloc(&out, blocks * 256 * 4)); CHECK(cudaMalloc(&cyc, blocks * 8));
    k<MODE><<<blocks, 256>>>(out, iters, cyc);
    CHECK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, iters, cyc);
    cudaEventRecord(e1);
    CHECK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148 * 4]; CHECK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < blocks; i++) if (h[i] > mx) mx = h[i];
    // 4 blocks of 256 threads per SM resident (1024 threads): FMA lane-ops per SM per cycle
    double lane_ops = flops_per_iter_per_thread * iters * 1024.0;
    printf("%-34s %.3f ms, max %lld cycles/block -> %.1f FMA lane-ops per clk per SM (%.1f TFLOP/s)\n", name, ms, mx, lane_ops / mx,
           2.0 * flops_per_iter_per_thread * iters * blocks * 256.0 / (ms * 1e-3) / 1e12);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    setvbuf(stdout, nullptr, _IONBF, 0);
    run<0>("scalar FFMA x12", 12);
    run<1>("packed FFMA2 x12", 24);
    run<2>("8 FFMA2 + 8 FFMA", 24);
    run<3>("8 FFMA2 + 4 FFMA", 20);
    return 0;
}
