// Device-side slice parser: launch interface (see vlc_kernel.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "host/slice_core.h"
#include "mp2v_recon.h"

namespace mp2v {

// One coded picture as staged for the device parser, a single H2D copy:
//   mp2v_pic_params_t (kVlcParamsBytes) | vlc_pic_header_t | vlc_slice_t[mbh] | (16-byte aligned) bitstream bytes + 16 bytes of zero padding
constexpr size_t kVlcParamsBytes = 512;
struct vlc_pic_header_t {
    slice_syntax_t sx;
    uint32_t n_slices;
    uint32_t slice_region;      // coefficient records reserved per slice: mbw * blocks * 64, the syntactic worst case
    uint32_t data_off;          // byte offset of the bitstream bytes from the start of the staged block
    uint32_t pad;
};
struct vlc_slice_t {
    uint32_t byte_off;          // payload offset inside the staged bitstream bytes
    int32_t code;               // slice_start_code value
};

// Result of parsing one slice, written by its thread with one plain 16-byte store into host memory the
// device can address (pinned + mapped): no read-back copy, no atomics; the host sums a picture's entries.
struct vlc_slice_status_t {
    uint32_t error;             // slice_error_t (0 = none)
    uint32_t n_coef;            // coefficient records written
    uint32_t coded_blocks;      // for the byte accounting of SURVEY.md 8(d)
    uint32_t ref_dirs;          // prediction directions summed over the slice's macroblocks
};

// ---- stream-resident front end: the whole elementary stream lies in device memory, a scan kernel lists its start
// codes, and pictures are handed over as descriptors (a few hundred bytes each, one H2D copy per parse launch) that
// name their slices by byte offset.  One launch parses every slice of up to kMaxStreamBatch pictures.
constexpr int kMaxStreamBatch = 128;
struct vlc_stream_pic_t {
    mp2v_pic_params_t params;           // copied into the slot's device parameter block by the kernel (the reconstruction kernel reads W from there)
    slice_syntax_t sx;
    uint32_t n_slices;
    uint32_t slice_region;              // coefficient records reserved per slice
    mp2v_pic_params_t* params_out;      // device addresses of the picture slot
    mp2v_mb_info_t* mb;
    mp2v_coef_t* coef;
    vlc_slice_status_t* status;         // pinned + mapped, one entry per slice
    uint32_t slice_off[1];              // n_slices byte offsets of the slices' start codes in the resident stream (the struct is over-allocated)
};
inline size_t vlc_stream_desc_bytes(int max_slices) { return (sizeof(vlc_stream_pic_t) + (size_t)max_slices * sizeof(uint32_t) + 15) & ~(size_t)15; }

// parse every slice of n_pics pictures described at desc (device copy, desc_stride bytes apart) out of the resident stream
cudaError_t launch_vlc_stream(const uint8_t* d_stream, const uint8_t* d_desc, size_t desc_stride, int n_pics, int max_slices, int lanes, const void* d_tables, cudaStream_t stream);

// start-code scan of a device-resident stream (the reference's scan_start_codes, start_codes_search.hpp:7-26): the offsets of
// every 00 00 01 prefix that STARTS in d_stream[0, len), ascending, each plus `base` (d_stream may be a 16-byte aligned part of
// a longer stream), into d_codes (capacity cap entries); *d_total = number found (may exceed cap: then only the first cap were
// stored).  d_counts: scratch of vlc_scan_blocks(len) + 1 entries.  d_stream must be readable, and hold the stream's bytes, up
// to len + 32.
size_t vlc_scan_blocks(size_t len);
cudaError_t launch_start_code_scan(const uint8_t* d_stream, size_t len, uint32_t base, uint32_t* d_counts, uint32_t* d_codes, uint32_t cap, uint32_t* d_total,
                                   cudaStream_t stream);

// the tables are copied to the device once per context
cudaError_t vlc_upload_tables(void** d_tables);
cudaError_t vlc_kernel_attributes(cudaFuncAttributes* out);

// parse every slice of one staged picture: macroblock records to mb[], coefficient records to coef[],
// one status entry per slice to status[]; macroblocks of a slice's row that the slice does not code
// are written as blank intra macroblocks
cudaError_t launch_vlc(const uint8_t* d_staged, const void* d_tables, mp2v_mb_info_t* mb, mp2v_coef_t* coef, vlc_slice_status_t* status,
                       int n_slices, int lanes, cudaStream_t stream);

}  // namespace mp2v
