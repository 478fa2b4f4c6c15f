// Internal: kernel launchers and the context layout shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include "../../include/dctc.h"

struct DctcK1Args;

// Programmatic dependent launch (the kernels of the device seam loop): a kernel launched with the attribute may be
// scheduled while its predecessor in the stream is still running; its CTAs then block in DCTC_PDL_PROLOGUE until the
// predecessor has completed and its memory operations are visible.  The macro is a no-op in a normal launch.
#ifdef DCTC_PDL_NO_EARLY
#define DCTC_PDL_PROLOGUE() asm volatile("griddepcontrol.wait;" ::: "memory")
#else
#define DCTC_PDL_PROLOGUE()                                                                                            \
    do {                                                                                                               \
        asm volatile("griddepcontrol.launch_dependents;");                                                            \
        asm volatile("griddepcontrol.wait;" ::: "memory");                                                            \
    } while (0)
#endif

template <typename... KArgs, typename... Args>
static inline cudaError_t dctc_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                                          Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

cudaError_t dctc_launch_k1_tile(const DctcK1Args& a, int blocksize, int n_frames, bool uniform, cudaStream_t stream);
cudaError_t dctc_launch_k1_small(const DctcK1Args& a, int blocksize, int n_frames, bool uniform, int sm_count, cudaStream_t stream);
cudaError_t dctc_launch_k1_march8(const DctcK1Args& a, int n_frames, bool uniform, cudaStream_t stream);
cudaError_t dctc_launch_k1_tc8(const DctcK1Args& a, int n_frames, bool uniform, int* counter, int sm_count, cudaStream_t stream);
cudaError_t dctc_launch_k1_tc16(const DctcK1Args& a, int n_frames, bool uniform, int* counter, int sm_count, cudaStream_t stream);
cudaError_t dctc_launch_synth(uint8_t* d_img, int n_frames, size_t frame_stride, int w, int h, int channels,
                              size_t pitch, uint32_t seed, int pattern, int first_frame, int y_offset,
                              cudaStream_t stream);

constexpr int DCTC_SLOTS = 4;   // device staging slots of the host-buffer batch call (H2D, kernel and D2H of different frames overlap)
constexpr int DCTC_TC_COUNTERS = 64;

struct dctc_context {
    int device = 0;
    cudaStream_t stream = nullptr;      // compute
    cudaStream_t s_in = nullptr;        // H2D
    cudaStream_t s_out = nullptr;       // D2H
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_in[DCTC_SLOTS] = {}, ev_k[DCTC_SLOTS] = {}, ev_out[DCTC_SLOTS] = {};
    float edges = 0.5f, textures = 0.5f;  // defaults src/main.c:30-33
    int blocksize = 8;
    int kernel = DCTC_KERNEL_AUTO;
    int last_cuda = 0;
    unsigned long long launches = 0;
    int sm_count = 0;
    int* tc_counters = nullptr;         // work-item counters of the persistent tensor-core kernel (device)
    unsigned tc_next = 0;
    // staging slots of the host-buffer API
    uint8_t* d_in[DCTC_SLOTS] = {};
    float* d_out[DCTC_SLOTS] = {};
    size_t d_in_cap[DCTC_SLOTS] = {}, d_out_cap[DCTC_SLOTS] = {};
    // carver session (K2)
    uint8_t* c_img = nullptr;      // device image, pitch c_pitch, compacted in place per seam
    float* c_en = nullptr;         // device energy, pitch c_en_pitch floats (multiple of 4: 16-byte aligned rows)
    size_t c_en_pitch = 0;
    float* c_m = nullptr;          // cumulative map of the device seam DP (rebuilt per seam, read back by the back-track)
    int8_t* c_dir = nullptr;       // jump plane of the parallel back-track: per 32-row block and bottom column, the path's column offset at the block top
    int* c_seam_log = nullptr;     // seams of dctc_carver_resize_width, n_seams * h
    size_t c_seam_log_cap = 0;
    // visibility map (lqr_carver_set_dump_vmaps, src/render.c:374): raw = original column of every current pixel,
    // vs = removal order of every original pixel (0 = never removed); both w0 x h ints, only when requested
    int* c_raw = nullptr;
    int* c_vs = nullptr;
    int c_vs_depth = 0;
    int c_vs_w = 0;                // width of the frame the visibility map belongs to (the session's width when it was requested)
    bool c_dump_vmaps = false;
    int* c_seam = nullptr;         // h entries
    int* c_band = nullptr;         // c_band[0]: 'rebuild the cumulative map from scratch' flag of the incremental seam DP; c_band[4]: last-row column of the current seam (device)
    bool c_incremental = false;    // device seam loop: update the cumulative map incrementally (dctc_carver_set_incremental)
    bool c_m_valid = false;        // the cumulative plane holds the map of the image before the last removal, compacted over that seam
    float* c_band_vals = nullptr;  // packed band values
    float* h_mirror = nullptr;     // host mirror of the energy map (pinned)
    int* h_band = nullptr;         // pinned 2*h
    int c_w0 = 0, c_w = 0, c_h = 0, c_ch = 0;
    size_t c_pitch = 0;
    bool mirror_valid = false;
    // K3 energy-image export
    unsigned int* k3_lohi = nullptr;   // device (min, max) of the compressed energies, as float bit patterns
    uint8_t* k3_img = nullptr;
    size_t k3_img_cap = 0;
};

int dctc_fail_cuda(dctc_context* ctx, cudaError_t e);
void dctc_carver_release(dctc_context* ctx);
int dctc_carver_params_changed(dctc_context* ctx);   // rebuilds the resident energy map of a loaded carver session
