// Output-path kernel: planar 4:2:0 frame in the device pool -> NV12 in a caller's device
// buffer, so a GPU consumer (encoder, renderer, inference pre-processing) takes frames without the D2H of
// planar YUV that bounds the end-to-end decode rate.  The reference only ever writes planar
// Y, Cb, Cr from host memory (tiny_decoder/tiny_mp2v_dec.cpp:11-17); NV12 = the same Y plane followed by
// one plane of interleaved Cb/Cr pairs.
//
// Pure byte mover, HBM-bound: up to 32 frames per launch (blockIdx.y), every thread moves aligned 16-byte units (a 128-bit load + store for
// luma; two 64-bit loads, two byte permutes and one 128-bit store for chroma); algorithmic bytes =
// 2 x frame bytes (read + write).
#include "recon_kernels.cuh"

namespace mp2v {

namespace {

constexpr int kNv12Threads = 128, kNv12Rows = 8;

// A CTA walks groups of 8 output rows; a thread owns 16-byte column units of those rows: all 8 (chroma: 16) loads
// are issued before the first store, no divisions, streaming stores (the consumer, not this kernel, re-reads the data).
__global__ void __launch_bounds__(kNv12Threads) planar420_to_nv12_kernel(const __grid_constant__ nv12_batch_t b) {
    const nv12_frame_t& f = b.frame[blockIdx.y];
    const int units_per_row = b.width >> 4;                   // width is a multiple of 16
    const int rows = b.height + (b.height >> 1);
    for (int r0 = blockIdx.x * kNv12Rows; r0 < rows; r0 += gridDim.x * kNv12Rows) {
        for (int xu = threadIdx.x; xu < units_per_row; xu += kNv12Threads) {
            const int x = xu << 4;
            uint4 v[kNv12Rows];
#pragma unroll
            for (int k = 0; k < kNv12Rows; k++) {
                const int r = r0 + k;
                if (r < b.height) {
                    v[k] = __ldg(reinterpret_cast<const uint4*>(f.y + (size_t)r * b.stride_y + x));
                } else if (r < rows) {
                    const size_t off = (size_t)(r - b.height) * b.stride_c + (x >> 1);
                    const uint2 p = __ldg(reinterpret_cast<const uint2*>(f.cb + off)), q = __ldg(reinterpret_cast<const uint2*>(f.cr + off));
                    v[k] = make_uint4(__byte_perm(p.x, q.x, 0x5140), __byte_perm(p.x, q.x, 0x7362), __byte_perm(p.y, q.y, 0x5140), __byte_perm(p.y, q.y, 0x7362));
                }
            }
#pragma unroll
            for (int k = 0; k < kNv12Rows; k++)
                if (r0 + k < rows) __stcs(reinterpret_cast<uint4*>(f.dst + (size_t)(r0 + k) * b.dst_pitch + x), v[k]);
        }
    }
}

}  // namespace

cudaError_t launch_nv12(const nv12_batch_t& batch, cudaStream_t stream) {
    if (batch.n_frames < 1 || batch.n_frames > kMaxNv12Batch) return cudaErrorInvalidValue;
    const int groups = (batch.height + (batch.height >> 1) + kNv12Rows - 1) / kNv12Rows;
    int ctas = groups;
    const int cap = (148 * 16 + batch.n_frames - 1) / batch.n_frames;     // 16 resident CTAs of 128 threads per SM over the whole launch
    if (ctas > cap) ctas = cap < 1 ? 1 : cap;                              // (one CTA per row group measured 9 % slower)
    planar420_to_nv12_kernel<<<dim3((unsigned)ctas, (unsigned)batch.n_frames), kNv12Threads, 0, stream>>>(batch);
    return cudaGetLastError();
}

}  // namespace mp2v
