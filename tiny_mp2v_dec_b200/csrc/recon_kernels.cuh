// Device-side interface between recon_api.cu (context, scheduling) and recon_kernels.cu (kernels).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "mp2v_recon.h"

namespace mp2v {

constexpr int kMaxBatch = 128;       // pictures fused into one launch (descriptors travel as kernel arguments: 104 B each, 13 KB of the 32 KB limit)
constexpr int kMaxConvertBatch = 96; // frames per output-conversion launch (32-byte descriptors as kernel arguments: 3 KB)
constexpr int kCtaThreads = 128;

// A warp owns `mbs_per_warp` consecutive macroblocks and walks them in batches of <= MP2V_SLOTS coded
// blocks.  The run length is a multiple of the all-coded batch size (24 slots: 4:2:0 4 MBs, 4:2:2 3,
// 4:4:4 2) and is chosen per launch so that the grid is at least ~6 waves of the resident warps
// (148 SMs x 8 CTAs x 4 warps): a grid of 1.1 waves costs two, which measured as a 40 % loss.
#ifndef MP2V_SLOTS
#define MP2V_SLOTS 24
#endif
#ifndef MP2V_WAVES
#define MP2V_WAVES 6
#endif
inline int choose_mbs_per_warp(int cf, long long total_mbs) {
    const int unit = MP2V_SLOTS / (cf == 1 ? 6 : cf == 2 ? 8 : 12);
    long long run = total_mbs / ((long long)MP2V_WAVES * 148 * 8 * 4);
    if (run > 60) run = 60;
    if (run < unit) run = unit;          // small launches (one picture): as many CTAs as possible
    return (int)(run / unit * unit);
}

struct pic_desc_t {
    const mp2v_pic_params_t* params;   // device copy of the picture parameters (W, alternate_scan)
    const mp2v_mb_info_t* mb;
    const mp2v_coef_t* coef;
    uint8_t* dst[3];
    const uint8_t* l0[3];
    const uint8_t* l1[3];
    int32_t l0_id, l1_id;              // the same references as frame ids: the z coordinate of the TMA boxes (-1: none)
};

struct batch_desc_t {
    pic_desc_t pic[kMaxBatch];
    int32_t n_pics;
    int32_t mbw, mbh, mb_count;
    int32_t stride[3];
    int32_t ctas_per_pic;
    int32_t mbs_per_warp;
    // (optional) device counters: [0] batches transformed, [1] of them with the exact (saturating) arithmetic in pass 2,
    // [2] in pass 1 as well -- how often the range analysis fails to prove a batch safe
    unsigned long long* counters;
};

// TMA descriptors of the frame pool, one per plane: a [frame][row][pixel] tensor of bytes whose boxes are the
// (w + 1) x (h + 1) reference windows of one macroblock (recon_kernels.cu)
struct alignas(64) recon_tmaps_t { CUtensorMap plane[3]; };
// frames: address of frame 0 plane 0; frames lie frame_alloc bytes apart.  Fails (cudaErrorNotSupported /
// cudaErrorInvalidValue) when the driver cannot encode the maps.
cudaError_t make_frame_tmaps(int chroma_format, uint8_t* frames, const mp2v_frame_layout_t& lay, size_t frame_alloc, int n_frames, recon_tmaps_t* out);

// grid = n_pics * ctas_per_pic; returns the CUDA error of the launch
cudaError_t launch_recon(int chroma_format, const batch_desc_t& batch, const recon_tmaps_t& tm, cudaStream_t stream);

// planar frames -> NV12 / P010 / UYVY (MP2V_OUT_*, mp2v_recon.h) in the caller's device buffers (convert_kernel.cu)
struct convert_frame_t { const uint8_t* y; const uint8_t* cb; const uint8_t* cr; uint8_t* dst; };
struct convert_batch_t {
    convert_frame_t frame[kMaxConvertBatch];
    int32_t n_frames, width, height, stride_y, stride_c, dst_pitch;
};
cudaError_t launch_convert(int format, const convert_batch_t& batch, cudaStream_t stream);

// registers / shared memory of the kernels as compiled, for DESIGN.md and the occupancy report
cudaError_t recon_kernel_attributes(int chroma_format, cudaFuncAttributes* out);

}  // namespace mp2v
