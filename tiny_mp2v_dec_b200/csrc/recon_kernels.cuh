// Device-side interface between recon_api.cu (context, scheduling) and recon_kernels.cu (kernels).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "mp2v_recon.h"

namespace mp2v {

constexpr int kMaxBatch = 32;        // pictures fused into one launch (descriptors travel as kernel arguments)
constexpr int kCtaThreads = 128;

// macroblocks per CTA, chosen so that an all-coded group fills the 128 IDCT threads:
// 4:2:0 20*6 = 120 blocks, 4:2:2 16*8 = 128, 4:4:4 10*12 = 120
__host__ __device__ constexpr int mbs_per_cta(int cf) { return cf == 1 ? 20 : cf == 2 ? 16 : 10; }

struct pic_desc_t {
    const mp2v_pic_params_t* params;   // device copy of the picture parameters (W, alternate_scan)
    const mp2v_mb_info_t* mb;
    const mp2v_coef_t* coef;
    uint8_t* dst[3];
    const uint8_t* l0[3];
    const uint8_t* l1[3];
    int32_t cta_begin;                 // first CTA of this picture inside the launch
    int32_t pad;
};

struct batch_desc_t {
    pic_desc_t pic[kMaxBatch];
    int32_t n_pics;
    int32_t mbw, mbh, mb_count;
    int32_t stride[3];
    int32_t ctas_per_pic;
};

// grid = n_pics * ctas_per_pic; returns the CUDA error of the launch
cudaError_t launch_recon(int chroma_format, const batch_desc_t& batch, cudaStream_t stream);

// registers / shared memory of the kernels as compiled, for DESIGN.md and the occupancy report
cudaError_t recon_kernel_attributes(int chroma_format, cudaFuncAttributes* out);

}  // namespace mp2v
