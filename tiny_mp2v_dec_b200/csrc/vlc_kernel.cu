// Device-side slice parser (SURVEY.md 8(f) "what comes next": the entropy decode in front of the
// reconstruction kernel).
//
// A slice is serial by construction -- every code's position depends on every code before it -- and
// MPEG-2 offers no entry points below the slice, so the unit of parallelism is the slice: one thread
// walks one slice with exactly the code the host parser runs (host/slice_core.h, compiled for both
// sides), writing the same records straight into the picture's device arenas.  What the GPU brings is
// breadth, not speed per symbol: every slice of every picture in flight is parsed at once (pictures
// have no parse-time dependencies, only reconstruction does), while the host only finds start codes.
//
// Mapping: `lanes` threads of each warp are active (1 by default).  Threads of a warp that sit on
// different slices diverge at every block and macroblock boundary, so a warp's time is the SUM of
// its lanes' paths; with one slice per warp the walk is a pure latency chain (table look-up -> shift
// -> look-up) and the warps of all slices hide one another's latency.  The parser's issue-slot cost
// is a few percent of the machine; the reconstruction kernel keeps the rest.
//
// Memory: tables (~155 KB, pointer-free, copied once) are read through the read-only path and stay
// L1/L2 resident; the bitstream is read 12 aligned bytes at a time (bitreader_t::refill); records are
// written sequentially by the owning thread and merge into full sectors in L2.
#include "vlc_kernel.cuh"

namespace mp2v {

namespace {

constexpr int kVlcCtaThreads = 128;

__global__ void __launch_bounds__(kVlcCtaThreads)
parse_slices_kernel(const uint8_t* __restrict__ staged, const vlc_decode_tables_t* __restrict__ tables, mp2v_mb_info_t* __restrict__ mb,
                    mp2v_coef_t* __restrict__ coef, vlc_slice_status_t* __restrict__ status, int lanes) {
    const vlc_pic_header_t& hdr = *reinterpret_cast<const vlc_pic_header_t*>(staged + kVlcParamsBytes);
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * kVlcCtaThreads + threadIdx.x) >> 5;
    const int slice = warp * lanes + lane;
    if (lane >= lanes || slice >= (int)hdr.n_slices) return;
    const vlc_slice_t sl = reinterpret_cast<const vlc_slice_t*>(staged + kVlcParamsBytes + sizeof(vlc_pic_header_t))[slice];
    const slice_syntax_t sx = hdr.sx;
    const uint32_t base = (uint32_t)slice * hdr.slice_region;
    uint32_t n = 0;
    int first_mbx = 0, last_mbx = -1, mb_row = 0;
    const int err = parse_slice_core<true>(staged + hdr.data_off + sl.byte_off, sl.code, sx, *tables, mb, coef + base, base, &n, &first_mbx, &last_mbx, &mb_row);
    uint32_t coded = 0, dirs = 0;
    if (mb_row >= 0 && mb_row < sx.mbh) {          // (the host checked the row before staging the slice)
        mp2v_mb_info_t* row = mb + (size_t)mb_row * sx.mbw;
        // accounting over the records this slice completed ...
        for (int x = first_mbx; x <= last_mbx; x++) {
            const uint32_t bits = row[x].bits;
            coded += __popc(MP2V_MB_CBP(bits));
            dirs += ((bits & MP2V_MB_FWD) ? 1u : 0u) + ((bits & MP2V_MB_BWD) ? 1u : 0u);
        }
        // ... and blank intra macroblocks (no coded block: reconstruct to 0) for the rest of its row
        const uint4 blank = make_uint4(0u, MP2V_MB_BITS(0, 1, 0, MP2V_MB_INTRA), 0u, 0u);
        if (last_mbx < first_mbx) { first_mbx = 0; last_mbx = -1; }
        for (int x = 0; x < first_mbx; x++) reinterpret_cast<uint4*>(row)[x] = blank;
        for (int x = last_mbx + 1; x < sx.mbw; x++) reinterpret_cast<uint4*>(row)[x] = blank;
    }
    reinterpret_cast<uint4*>(status)[slice] = make_uint4((uint32_t)err, n, coded, dirs);
}

}  // namespace

cudaError_t vlc_upload_tables(void** d_tables) {
    const vlc_decode_tables_t& t = vlc_decode_tables();
    cudaError_t e = cudaMalloc(d_tables, sizeof(vlc_decode_tables_t));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*d_tables, &t, sizeof(vlc_decode_tables_t), cudaMemcpyHostToDevice);
}

cudaError_t vlc_kernel_attributes(cudaFuncAttributes* out) { return cudaFuncGetAttributes(out, parse_slices_kernel); }

cudaError_t launch_vlc(const uint8_t* d_staged, const void* d_tables, mp2v_mb_info_t* mb, mp2v_coef_t* coef, vlc_slice_status_t* status,
                       int n_slices, int lanes, cudaStream_t stream) {
    if (n_slices <= 0) return cudaSuccess;
    const int warps = (n_slices + lanes - 1) / lanes;
    const int ctas = (warps * 32 + kVlcCtaThreads - 1) / kVlcCtaThreads;
    parse_slices_kernel<<<ctas, kVlcCtaThreads, 0, stream>>>(d_staged, static_cast<const vlc_decode_tables_t*>(d_tables), mb, coef, status, lanes);
    return cudaGetLastError();
}

}  // namespace mp2v
