// Device-side slice parser and start-code scan (the entropy decode in front of the
// reconstruction kernel).
//
// A slice is serial by construction -- every code's position depends on every code before it -- and
// MPEG-2 offers no entry points below the slice, so the unit of parallelism is the slice: one thread
// walks one slice with exactly the code the host parser runs (host/slice_core.h, compiled for both
// sides), writing the same records straight into the picture's device arenas.  What the GPU brings is
// breadth, not speed per symbol: every slice of every picture in flight is parsed at once (pictures
// have no parse-time dependencies, only reconstruction does), while the host only finds start codes.
//
// Mapping: `lanes` threads of each warp are active, each on its own slice.  Threads of a warp that sit on
// different slices serialise wherever their paths differ (they wait for one another at block ends), so a
// warp's time grows with its lanes; but a lane's registers are held by its whole warp, and a lot of
// single-lane warps fills the register files and keeps the reconstruction launches of the previous lot out.
// The launcher (recon_api.cu: launch_parse_batch) therefore puts one slice in a warp for small lots and two
// for large ones -- measured best on the whole decode.  A slice's walk is a latency chain
// (look-up -> shift -> look-up); the warps of all slices hide one another's latency, and the cost of the
// kernel is its instruction count: two thirds of the issue slots of a resident decode, the
// reconstruction kernel has the rest.
//
// Memory: the run/level fast and long-code tables (24 KB) are copied to shared memory by every CTA; the other
// tables (pointer-free, copied to the device once) are read through the read-only path and stay L1/L2
// resident; the bitstream is read as aligned 32-bit words with a one-word look-ahead (bitreader.h); records
// are written sequentially by the owning thread and merge into full sectors in L2.
//
// The same file holds the start-code scan of the stream-resident front end (scan_*_kernel below).
#include "vlc_kernel.cuh"

namespace mp2v {

namespace {

#ifndef MP2V_VLC_CTA
#define MP2V_VLC_CTA 128
#endif
constexpr int kVlcCtaThreads = MP2V_VLC_CTA;

// the run/level tables of the symbol loop (B.14, B.15: 8 KB of fast entries and 4 KB of long codes each) live in
// shared memory: a look-up per coefficient
__device__ __forceinline__ void load_fast_tables(uint32_t* s_tab, const vlc_decode_tables_t* __restrict__ tables) {
    const uint32_t* src[4] = {tables->b14.gpu_fast, tables->b15.gpu_fast, tables->b14.gpu_long, tables->b15.gpu_long};
    constexpr int words[4] = {1 << kFastBits, 1 << kFastBits, 1 << coef_vlc_t::kLongBits, 1 << coef_vlc_t::kLongBits};
    uint32_t* d = s_tab;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        for (int i = threadIdx.x; i < words[t]; i += kVlcCtaThreads) d[i] = __ldg(src[t] + i);
        d += words[t];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kVlcCtaThreads, 1024 / kVlcCtaThreads)
parse_slices_kernel(const uint8_t* __restrict__ staged, const vlc_decode_tables_t* __restrict__ tables, mp2v_mb_info_t* __restrict__ mb,
                    mp2v_coef_t* __restrict__ coef, vlc_slice_status_t* __restrict__ status, int lanes) {
    __shared__ __align__(16) uint32_t s_fast[kDevTableWords];
    load_fast_tables(s_fast, tables);
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
    const int err = parse_slice_core<true>(staged + hdr.data_off + sl.byte_off, sl.code, sx, *tables, mb, coef + base, base, &n, &first_mbx, &last_mbx, &mb_row, (uint32_t)__cvta_generic_to_shared(s_fast));
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

// ---- stream-resident variant: grid.y = picture of the batch, one warp per slice, descriptors in device memory
__global__ void __launch_bounds__(kVlcCtaThreads, 1024 / kVlcCtaThreads)
parse_stream_slices_kernel(const uint8_t* __restrict__ stream, const uint8_t* __restrict__ desc_base, size_t desc_stride,
                           const vlc_decode_tables_t* __restrict__ tables, int lanes) {
    __shared__ __align__(16) uint32_t s_fast[kDevTableWords];
    load_fast_tables(s_fast, tables);
    const vlc_stream_pic_t& d = *reinterpret_cast<const vlc_stream_pic_t*>(desc_base + (size_t)blockIdx.y * desc_stride);
    // the picture's parameter block (W, scan, frame ids) travels in the descriptor: the first CTA drops it where the
    // reconstruction kernel reads it (that launch is ordered behind this one)
    if (blockIdx.x == 0)
        for (unsigned i = threadIdx.x; i < sizeof(mp2v_pic_params_t) / 4; i += kVlcCtaThreads)
            reinterpret_cast<uint32_t*>(d.params_out)[i] = reinterpret_cast<const uint32_t*>(&d.params)[i];
    const int lane = threadIdx.x & 31;
    const int slice = ((blockIdx.x * kVlcCtaThreads + threadIdx.x) >> 5) * lanes + lane;
    if (lane >= lanes || slice >= (int)d.n_slices) return;
    const slice_syntax_t sx = d.sx;
    const uint8_t* sc = stream + d.slice_off[slice];             // 00 00 01 <slice_start_code> payload...
    const uint32_t base = (uint32_t)slice * d.slice_region;
    uint32_t n = 0;
    int first_mbx = 0, last_mbx = -1, mb_row = 0;
    const int err = parse_slice_core<true>(sc + 4, (int)sc[3], sx, *tables, d.mb, d.coef + base, base, &n, &first_mbx, &last_mbx, &mb_row, (uint32_t)__cvta_generic_to_shared(s_fast));
    uint32_t coded = 0, dirs = 0;
    if (mb_row >= 0 && mb_row < sx.mbh) {
        mp2v_mb_info_t* row = d.mb + (size_t)mb_row * sx.mbw;
        for (int x = first_mbx; x <= last_mbx; x++) {
            const uint32_t bits = row[x].bits;
            coded += __popc(MP2V_MB_CBP(bits));
            dirs += ((bits & MP2V_MB_FWD) ? 1u : 0u) + ((bits & MP2V_MB_BWD) ? 1u : 0u);
        }
        const uint4 blank = make_uint4(0u, MP2V_MB_BITS(0, 1, 0, MP2V_MB_INTRA), 0u, 0u);
        if (last_mbx < first_mbx) { first_mbx = 0; last_mbx = -1; }
        for (int x = 0; x < first_mbx; x++) reinterpret_cast<uint4*>(row)[x] = blank;
        for (int x = last_mbx + 1; x < sx.mbw; x++) reinterpret_cast<uint4*>(row)[x] = blank;
    }
    reinterpret_cast<uint4*>(d.status)[slice] = make_uint4((uint32_t)err, n, coded, dirs);
}

// ---- start-code scan.  A CTA owns kScanChunk bytes; a thread 16 of them (plus two bytes of look-ahead).
constexpr int kScanThreads = 256, kScanChunk = kScanThreads * 16;

__device__ __forceinline__ uint32_t start_code_mask(const uint8_t* s, size_t pos, size_t len) {
    // bit i: a 00 00 01 prefix starts at pos + i (i < 16); the buffer is readable (zero padded) past len
    const uint4 a = *reinterpret_cast<const uint4*>(s + pos);
    const uint32_t nx = *reinterpret_cast<const uint32_t*>(s + pos + 16);
    const uint32_t w[5] = {a.x, a.y, a.z, a.w, nx};
    uint32_t zero = 0, one = 0;          // bit i: byte i is 0x00 / 0x01 (18 bytes)
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const uint32_t z = __vcmpeq4(w[k], 0u), o = __vcmpeq4(w[k], 0x01010101u);
        zero |= ((((z >> 7) & 0x01010101u) * 0x01020408u) >> 24 & 0xfu) << (4 * k);
        one |= ((((o >> 7) & 0x01010101u) * 0x01020408u) >> 24 & 0xfu) << (4 * k);
    }
    uint32_t m = zero & (zero >> 1) & (one >> 2) & 0xffffu;
    if (pos + 16 > len) m &= pos < len ? (1u << (len - pos)) - 1u : 0u;
    return m;
}

__global__ void __launch_bounds__(kScanThreads) scan_count_kernel(const uint8_t* __restrict__ s, size_t len, uint32_t* __restrict__ counts) {
    const size_t pos = (size_t)blockIdx.x * kScanChunk + (size_t)threadIdx.x * 16;
    const int c = __popc(start_code_mask(s, pos, len));
    __shared__ uint32_t part[kScanThreads / 32];
    const uint32_t wsum = __reduce_add_sync(0xffffffffu, (uint32_t)c);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = wsum;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int i = 0; i < kScanThreads / 32; i++) t += part[i]; counts[blockIdx.x] = t; }
}

// exclusive prefix of the per-CTA counts, in place; counts[n] = total (one CTA walks the array in tiles)
__global__ void __launch_bounds__(1024) scan_prefix_kernel(uint32_t* __restrict__ counts, uint32_t n, uint32_t* __restrict__ total) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? counts[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, x, d); if ((threadIdx.x & 31) >= d) x += t; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t y = wsum[threadIdx.x];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, y, d); if (threadIdx.x >= d) y += t; }
            wsum[threadIdx.x] = y;
        }
        __syncthreads();
        const uint32_t before = carry_s + (threadIdx.x >= 32 ? wsum[(threadIdx.x >> 5) - 1] : 0u);
        if (i < n) counts[i] = before + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) { counts[n] = carry_s; *total = carry_s; }
}

__global__ void __launch_bounds__(kScanThreads) scan_write_kernel(const uint8_t* __restrict__ s, size_t len, const uint32_t* __restrict__ counts,
                                                                 uint32_t* __restrict__ codes, uint32_t cap, uint32_t base) {
    const size_t pos = (size_t)blockIdx.x * kScanChunk + (size_t)threadIdx.x * 16;
    uint32_t m = start_code_mask(s, pos, len);
    const uint32_t c = __popc(m);
    __shared__ uint32_t part[kScanThreads / 32];
    uint32_t x = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, x, d); if ((threadIdx.x & 31) >= d) x += t; }
    if ((threadIdx.x & 31) == 31) part[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t before = counts[blockIdx.x] + x - c;
    for (int i = 0; i < (int)(threadIdx.x >> 5); i++) before += part[i];
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        if (before < cap) codes[before] = base + (uint32_t)(pos + b);
        before++;
    }
}

}  // namespace

size_t vlc_scan_blocks(size_t len) { return (len + kScanChunk - 1) / kScanChunk; }

cudaError_t launch_start_code_scan(const uint8_t* d_stream, size_t len, uint32_t base, uint32_t* d_counts, uint32_t* d_codes, uint32_t cap, uint32_t* d_total,
                                   cudaStream_t stream) {
    const size_t nb = vlc_scan_blocks(len);
    if (nb == 0) return cudaMemsetAsync(d_total, 0, sizeof(uint32_t), stream);
    if (len > 0xfffffff0ull || (uint64_t)base + len > 0xfffffff0ull || ((uintptr_t)d_stream & 15)) return cudaErrorInvalidValue;      // 32-bit offsets, 16-byte loads
    scan_count_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(d_stream, len, d_counts);
    scan_prefix_kernel<<<1, 1024, 0, stream>>>(d_counts, (uint32_t)nb, d_total);
    scan_write_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(d_stream, len, d_counts, d_codes, cap, base);
    return cudaGetLastError();
}

cudaError_t launch_vlc_stream(const uint8_t* d_stream, const uint8_t* d_desc, size_t desc_stride, int n_pics, int max_slices, int lanes, const void* d_tables, cudaStream_t stream) {
    if (n_pics <= 0 || max_slices <= 0) return cudaSuccess;
    if (n_pics > kMaxStreamBatch || lanes < 1 || lanes > 32) return cudaErrorInvalidValue;
    const int warps = (max_slices + lanes - 1) / lanes;
    const dim3 grid((unsigned)((warps * 32 + kVlcCtaThreads - 1) / kVlcCtaThreads), (unsigned)n_pics);
    parse_stream_slices_kernel<<<grid, kVlcCtaThreads, 0, stream>>>(d_stream, d_desc, desc_stride, static_cast<const vlc_decode_tables_t*>(d_tables), lanes);
    return cudaGetLastError();
}

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
