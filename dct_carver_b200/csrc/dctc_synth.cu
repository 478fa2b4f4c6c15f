// Synthetic image generator on the device (bench + tests); see dctc_synth_px in dctc_common.cuh.
#include "dctc_common.cuh"
#include "dctc_launch.h"

__global__ void dctc_synth_kernel(uint8_t* __restrict__ img, size_t frame_stride, int w, int h, int channels,
                                  size_t pitch, uint32_t seed, int pattern, int first_frame, int y_offset)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w) return;
    uint8_t* p = img + (size_t) blockIdx.z * frame_stride + (size_t) y * pitch + (size_t) x * channels;
    for (int c = 0; c < channels; c++)
        p[c] = dctc_synth_px(seed, (uint32_t) (first_frame + blockIdx.z), (uint32_t) (y + y_offset), (uint32_t) x, (uint32_t) c, pattern);
}

cudaError_t dctc_launch_synth(uint8_t* d_img, int n_frames, size_t frame_stride, int w, int h, int channels,
                              size_t pitch, uint32_t seed, int pattern, int first_frame, int y_offset,
                              cudaStream_t stream)
{
    if (w <= 0 || h <= 0 || n_frames <= 0) return cudaSuccess;
    if (h > 65535 || n_frames > 65535) return cudaErrorInvalidConfiguration;
    dim3 block(256), grid((w + 255) / 256, h, n_frames);
    dctc_synth_kernel<<<grid, block, 0, stream>>>(d_img, frame_stride, w, h, channels, pitch, seed, pattern, first_frame, y_offset);
    return cudaGetLastError();
}

extern "C" uint8_t dctc_synth_byte(uint32_t seed, uint32_t frame, uint32_t y, uint32_t x, uint32_t c, int pattern)
{
    return dctc_synth_px(seed, frame, y, x, c, pattern);
}
