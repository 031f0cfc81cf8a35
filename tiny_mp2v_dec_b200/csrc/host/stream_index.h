// One pass over an elementary stream: start codes -> coded pictures with their headers and slice
// payload pointers.  This is the role of mp2v_decoder_c::decode()'s start-code switch in the
// reference (decoder.cpp:278-329), separated from scheduling so that the GPU decoder, the host-only
// parse benchmark and the tests share it.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "mp2v_parser.h"
#include "mp2v_stream_headers.h"

namespace mp2v {

struct slice_ref_t {
    const uint8_t* payload;   // first byte after the 4-byte start code
    int code;                 // slice_start_code value 0x01..0xAF
    uint32_t bytes;           // payload length: up to the next start code prefix (or the end of the buffer)
};

struct coded_picture_t {
    picture_info_t info;
    sequence_info_t seq;      // sequence state in force for this picture
    std::vector<slice_ref_t> slices;
    int gop = 0;              // index of the closed GOP chain this picture belongs to (sharding unit)
};

// the sequence-level headers as the decode API publishes them (mp2v_decoder_c's public members): last occurrence of each
struct stream_headers_t {
    sequence_header_t sequence_header = {};
    sequence_extension_t sequence_extension = {};
    sequence_display_extension_t sequence_display_extension = {};
    sequence_scalable_extension_t sequence_scalable_extension = {};
    group_of_pictures_header_t group_of_pictures_header = {};
    bool have_display_extension = false, have_scalable_extension = false, have_gop_header = false;
    std::vector<uint8_t> user_data;          // every user_data() payload, stream order
};

struct stream_index_t {
    std::vector<coded_picture_t> pictures;   // coded order
    stream_headers_t headers;
    int n_gops = 0;
    std::string error;
};

// buffer must be followed by >= 64 readable bytes (decode() contract).  The start-code scan (the only
// part that touches every byte) is split over `threads` threads; header parsing stays serial.
bool index_stream(const uint8_t* buffer, size_t len, stream_index_t& out, int threads = 1);
// the same from a given list of start-code offsets (ascending; e.g. the device-side scan of mp2v_recon_stream_begin)
bool index_stream_from_codes(const uint8_t* buffer, size_t len, const uint32_t* codes, size_t n_codes, stream_index_t& out);

}  // namespace mp2v
