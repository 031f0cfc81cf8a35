// Host side of the reconstruction seam: MPEG-2 header and slice parsing that EMITS per-picture
// structure-of-arrays records (include/mp2v_recon.h) instead of calling IDCT / motion compensation
// per block as the reference does (mb_decoder.cpp:547-548, 610-615, 633-635).
//
// What stays on the host is what is serial by nature: VLC decoding, intra DC prediction
// (mb_decoder.cpp:46-72), motion-vector prediction (mb_decoder.cpp:447-519, 580-604), quantiser
// scale tracking and skipped-macroblock resolution (mb_decoder.cpp:541-550).  No dequantisation
// happens here: levels are shipped, products are formed on the device.
#pragma once
#include <atomic>
#include <cstdint>

#include "mp2v_recon.h"

namespace mp2v {

struct sequence_info_t {
    int horizontal_size = 0, vertical_size = 0;
    int chroma_format = 1;
    int progressive_sequence = 1;
    bool have_sequence_header = false, have_sequence_extension = false;
    // matrices in force (zig-zag order): defaults until a sequence header or a quant_matrix_extension loads them;
    // they persist until the next sequence header (ISO/IEC 13818-2 6.3.11), the chroma pair included
    uint8_t intra_matrix[64], non_intra_matrix[64], chroma_intra_matrix[64], chroma_non_intra_matrix[64];
    sequence_info_t();
};

struct picture_info_t {
    int temporal_reference = 0;
    int picture_coding_type = 0;       // 1 I, 2 P, 3 B
    int f_code[2][2] = {{15, 15}, {15, 15}};
    int intra_dc_precision = 0;
    int picture_structure = 3;
    int frame_pred_frame_dct = 1;
    int concealment_motion_vectors = 0;
    int q_scale_type = 0;
    int intra_vlc_format = 0;
    int alternate_scan = 0;
    bool have_coding_extension = false;
    uint8_t tx[4][64];                 // matrices in force for this picture, zig-zag order as transmitted
};

// Parsers return false on malformed / unsupported syntax (the reference returns true unconditionally
// and has undefined behaviour on bad input).
bool parse_sequence_header(const uint8_t* payload, sequence_info_t& seq);
bool parse_extension(const uint8_t* payload, sequence_info_t& seq, picture_info_t* pic);   // dispatches on the extension id
bool parse_picture_header(const uint8_t* payload, const sequence_info_t& seq, picture_info_t& pic);

// quantiser_matrices of mp2v_picture_c::init() (decoder.cpp:154-192) for all four sets
void build_picture_matrices(const picture_info_t& pic, uint8_t W[4][64]);

// Coefficient arena of one picture, shared by the slice parsers of that picture (one thread each).
// A slice is parsed into the calling thread's scratch buffer, then appended with ONE atomic add and
// one memcpy, so the arena stays dense: its high-water mark is exactly what the H2D copy moves.
struct coef_arena_t {
    mp2v_coef_t* base = nullptr;
    uint32_t capacity = 0;
    std::atomic<uint32_t> next{0};
    std::atomic<bool> overflow{false};
};

struct slice_result_t {
    int mbs = 0;                 // macroblock records written (skipped ones included)
    bool ok = true;
    const char* error = nullptr;
};

// Parse one slice (payload = first byte after the 4-byte start code; mb_row from the start code and,
// for tall pictures, slice_vertical_position_extension) and write its macroblock records to
// mb[mb_row * mbw + ...] and its coefficient records into the arena.
slice_result_t parse_slice(const uint8_t* payload, int slice_start_code, const sequence_info_t& seq, const picture_info_t& pic,
                           int mbw, int mbh, mp2v_mb_info_t* mb, coef_arena_t& arena);

// the POD view of the headers that the shared slice core (slice_core.h) and the GPU-side parser take
struct slice_syntax_t;
slice_syntax_t make_slice_syntax(const sequence_info_t& seq, const picture_info_t& pic, int mbw, int mbh);
bool picture_in_envelope(const picture_info_t& pic);

// locate the next start code prefix (00 00 01) in [p, end); returns end if none
const uint8_t* find_start_code(const uint8_t* p, const uint8_t* end);

}  // namespace mp2v
