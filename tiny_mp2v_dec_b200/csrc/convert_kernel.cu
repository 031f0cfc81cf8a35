// Output-path kernels: planar frames in the device pool -> the packed / semi-planar layouts GPU
// consumers (encoders, renderers, inference pre-processing) take, written into a caller's device buffer, so that
// the D2H of planar YUV that bounds the end-to-end decode rate is optional.  The reference only ever
// writes planar Y, Cb, Cr from host memory (tiny_decoder/tiny_mp2v_dec.cpp:11-17).
//   NV12  (4:2:0)  the Y plane, then rows of interleaved Cb/Cr pairs                            1.5 bytes / pixel
//   P010  (4:2:0)  the same layout with 16-bit samples, the 8 decoded bits in the high byte      3 bytes / pixel
//   UYVY  (4:2:2)  packed Cb Y0 Cr Y1 per pixel pair                                              2 bytes / pixel
// Pure byte movers, HBM-bound: up to 96 frames per launch (blockIdx.y); a CTA walks groups of output rows, a thread
// owns 16-pixel column units of those rows and issues all its loads before the first (streaming) store; no divisions.
// Algorithmic bytes = frame bytes read + output bytes written.
#include <algorithm>

#include "recon_kernels.cuh"

namespace mp2v {

namespace {

constexpr int kCvtThreads = 128;

// 4 bytes -> 4 samples of 16 bits with the byte in the high half (two words)
__device__ __forceinline__ uint2 widen4(uint32_t v) { return make_uint2(__byte_perm(v, 0u, 0x1404), __byte_perm(v, 0u, 0x3424)); }

template <int FMT, int ROWS>
__global__ void __launch_bounds__(kCvtThreads) convert_kernel(const __grid_constant__ convert_batch_t b) {
    const convert_frame_t& f = b.frame[blockIdx.y];
    constexpr int PX = 16;                                    // pixels per thread and row (8 with one 16-byte P010 store measured 53 % instead of 71 %)
    const int units_per_row = b.width / PX;                    // width is a multiple of 16
    const int rows = FMT == MP2V_OUT_UYVY ? b.height : b.height + (b.height >> 1);
    constexpr int W = FMT == MP2V_OUT_NV12 ? 1 : 2;           // 16-byte stores per unit and row
    for (int r0 = blockIdx.x * ROWS; r0 < rows; r0 += gridDim.x * ROWS) {
        for (int xu = threadIdx.x; xu < units_per_row; xu += kCvtThreads) {
            const int x = xu * PX;
            uint4 v[ROWS][W];
#pragma unroll
            for (int k = 0; k < ROWS; k++) {
                const int r = r0 + k;
                if (r >= rows) continue;
                if (FMT == MP2V_OUT_UYVY) {
                    const uint4 y = __ldg(reinterpret_cast<const uint4*>(f.y + (size_t)r * b.stride_y + x));
                    const size_t off = (size_t)r * b.stride_c + (x >> 1);
                    const uint2 p = __ldg(reinterpret_cast<const uint2*>(f.cb + off)), q = __ldg(reinterpret_cast<const uint2*>(f.cr + off));
                    const uint32_t uv0 = __byte_perm(p.x, q.x, 0x5140), uv1 = __byte_perm(p.x, q.x, 0x7362);      // u0 v0 u1 v1 | u2 v2 u3 v3
                    const uint32_t uv2 = __byte_perm(p.y, q.y, 0x5140), uv3 = __byte_perm(p.y, q.y, 0x7362);
                    v[k][0] = make_uint4(__byte_perm(uv0, y.x, 0x5140), __byte_perm(uv0, y.x, 0x7362), __byte_perm(uv1, y.y, 0x5140), __byte_perm(uv1, y.y, 0x7362));
                    v[k][W - 1] = make_uint4(__byte_perm(uv2, y.z, 0x5140), __byte_perm(uv2, y.z, 0x7362), __byte_perm(uv3, y.w, 0x5140), __byte_perm(uv3, y.w, 0x7362));
                } else {
                    uint4 s;                                   // 16 output samples of 8 bits: luma, or interleaved Cb/Cr pairs
                    if (r < b.height) {
                        s = __ldg(reinterpret_cast<const uint4*>(f.y + (size_t)r * b.stride_y + x));
                    } else {
                        const size_t off = (size_t)(r - b.height) * b.stride_c + (x >> 1);
                        const uint2 p = __ldg(reinterpret_cast<const uint2*>(f.cb + off)), q = __ldg(reinterpret_cast<const uint2*>(f.cr + off));
                        s = make_uint4(__byte_perm(p.x, q.x, 0x5140), __byte_perm(p.x, q.x, 0x7362), __byte_perm(p.y, q.y, 0x5140), __byte_perm(p.y, q.y, 0x7362));
                    }
                    if (FMT == MP2V_OUT_NV12) v[k][0] = s;
                    else {
                        const uint2 a = widen4(s.x), c = widen4(s.y), d = widen4(s.z), e = widen4(s.w);
                        v[k][0] = make_uint4(a.x, a.y, c.x, c.y);
                        v[k][W - 1] = make_uint4(d.x, d.y, e.x, e.y);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < ROWS; k++) {
                if (r0 + k >= rows) continue;
                uint4* dst = reinterpret_cast<uint4*>(f.dst + (size_t)(r0 + k) * b.dst_pitch + (size_t)xu * (16 * W));
#pragma unroll
                for (int j = 0; j < W; j++) __stcs(dst + j, v[k][j]);
            }
        }
    }
}

}  // namespace

cudaError_t launch_convert(int format, const convert_batch_t& batch, cudaStream_t stream) {
    if (batch.n_frames < 1 || batch.n_frames > kMaxConvertBatch) return cudaErrorInvalidValue;
    const int rows = format == MP2V_OUT_UYVY ? batch.height : batch.height + (batch.height >> 1);
    const int rows_per_group = format == MP2V_OUT_NV12 ? 8 : 4;
    // one wave: 16 resident CTAs of 128 threads per SM over the whole launch, and every CTA of a frame the same number of
    // row groups (1080p NV12: 204 groups of 8 rows = 68 CTAs x 3; an uneven split measured 5 % slower, one CTA per group 9 %)
    const int groups = (rows + rows_per_group - 1) / rows_per_group;
    const int cap = std::max(1, (148 * 16 + batch.n_frames - 1) / batch.n_frames);
    const int rounds = (groups + cap - 1) / cap;
    const int ctas = (groups + rounds - 1) / rounds;
    const dim3 grid((unsigned)ctas, (unsigned)batch.n_frames);
    switch (format) {
        case MP2V_OUT_NV12: convert_kernel<MP2V_OUT_NV12, 8><<<grid, kCvtThreads, 0, stream>>>(batch); break;
        case MP2V_OUT_P010: convert_kernel<MP2V_OUT_P010, 4><<<grid, kCvtThreads, 0, stream>>>(batch); break;
        case MP2V_OUT_UYVY: convert_kernel<MP2V_OUT_UYVY, 4><<<grid, kCvtThreads, 0, stream>>>(batch); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace mp2v
