// Header + slice parser emitting reconstruction records -- see mp2v_parser.h.
// Syntax per ISO/IEC 13818-2 6.2; behaviour checked against the reference's parsers
// (src/core/mp2v_hdr.cpp, mb_decoder.cpp) through tests/test_parser.py and the end-to-end parity tests.
#include "mp2v_parser.h"

#include <cstring>
#include <vector>

#include "bitreader.h"
#include "scan_tables.h"
#include "slice_core.h"
#include "vlc_decode.h"

namespace mp2v {

namespace {

// ISO/IEC 13818-2 6.3.11 default intra matrix, raster order
const uint8_t kDefaultIntraRaster[64] = {
     8, 16, 19, 22, 26, 27, 29, 34, 16, 16, 22, 24, 27, 29, 34, 37, 19, 22, 26, 27, 29, 34, 34, 38,
    22, 22, 26, 27, 29, 34, 37, 40, 22, 26, 27, 29, 32, 35, 40, 48, 26, 27, 29, 32, 35, 40, 48, 58,
    26, 27, 29, 34, 38, 46, 56, 69, 27, 29, 35, 38, 46, 56, 69, 83 };

void read_matrix(bitreader_t& br, uint8_t m[64]) {
    for (int i = 0; i < 64; i++) m[i] = (uint8_t)br.get(8);
}

}  // namespace

sequence_info_t::sequence_info_t() {
    const scan_tables_t& t = scan_tables();
    for (int i = 0; i < 64; i++) {
        intra_matrix[i] = chroma_intra_matrix[i] = kDefaultIntraRaster[t.shuffle[0][i]];
        non_intra_matrix[i] = chroma_non_intra_matrix[i] = 16;
    }
}

const uint8_t* find_start_code(const uint8_t* p, const uint8_t* end) {
    // start codes are sparse: let memchr find the 0x01 bytes, then look back
    p += 2;
    while (p < end) {
        const uint8_t* q = (const uint8_t*)memchr(p, 1, (size_t)(end - p));
        if (!q) break;
        if (q[-1] == 0 && q[-2] == 0) return q - 2;
        p = q + 1;
    }
    return end;
}

bool parse_sequence_header(const uint8_t* payload, sequence_info_t& seq) {
    bitreader_t br(payload);
    seq.horizontal_size = (int)br.get(12);
    seq.vertical_size = (int)br.get(12);
    br.get(4); br.get(4);                       // aspect_ratio_information, frame_rate_code
    br.get(18); br.get(1); br.get(10); br.get(1);   // bit_rate_value, marker, vbv_buffer_size_value, constrained_parameters_flag
    sequence_info_t defaults;
    memcpy(seq.intra_matrix, defaults.intra_matrix, 64);
    memcpy(seq.non_intra_matrix, defaults.non_intra_matrix, 64);
    if (br.get1()) read_matrix(br, seq.intra_matrix);
    if (br.get1()) read_matrix(br, seq.non_intra_matrix);
    // a sequence header resets all four: the chroma matrices take the luminance ones (6.3.11)
    memcpy(seq.chroma_intra_matrix, seq.intra_matrix, 64);
    memcpy(seq.chroma_non_intra_matrix, seq.non_intra_matrix, 64);
    seq.have_sequence_header = true;
    return seq.horizontal_size > 0 && seq.vertical_size > 0;
}

bool parse_picture_header(const uint8_t* payload, const sequence_info_t& seq, picture_info_t& pic) {
    bitreader_t br(payload);
    pic = picture_info_t();
    pic.temporal_reference = (int)br.get(10);
    pic.picture_coding_type = (int)br.get(3);
    br.get(16);                                 // vbv_delay
    if (pic.picture_coding_type == 2 || pic.picture_coding_type == 3) br.get(4);   // full_pel_forward_vector, forward_f_code (MPEG-1 fields)
    if (pic.picture_coding_type == 3) br.get(4);
    // matrices in force: sequence-level ones until a quant_matrix_extension replaces them (6.3.11)
    memcpy(pic.tx[0], seq.intra_matrix, 64); memcpy(pic.tx[2], seq.chroma_intra_matrix, 64);
    memcpy(pic.tx[1], seq.non_intra_matrix, 64); memcpy(pic.tx[3], seq.chroma_non_intra_matrix, 64);
    return pic.picture_coding_type >= 1 && pic.picture_coding_type <= 3;
}

bool parse_extension(const uint8_t* payload, sequence_info_t& seq, picture_info_t* pic) {
    bitreader_t br(payload);
    const int id = (int)br.get(4);
    switch (id) {
    case 1: {   // sequence_extension (mp2v_hdr.cpp:23-37)
        br.get(8);                                  // profile_and_level_indication
        seq.progressive_sequence = (int)br.get(1);
        seq.chroma_format = (int)br.get(2);
        seq.horizontal_size |= (int)br.get(2) << 12;
        seq.vertical_size |= (int)br.get(2) << 12;
        seq.have_sequence_extension = true;
        return seq.chroma_format >= 1 && seq.chroma_format <= 3;
    }
    case 8: {   // picture_coding_extension (mp2v_hdr.cpp:105-131)
        if (!pic) return false;
        for (int s = 0; s < 2; s++) for (int t = 0; t < 2; t++) pic->f_code[s][t] = (int)br.get(4);
        pic->intra_dc_precision = (int)br.get(2);
        pic->picture_structure = (int)br.get(2);
        br.get(1);                                  // top_field_first
        pic->frame_pred_frame_dct = (int)br.get(1);
        pic->concealment_motion_vectors = (int)br.get(1);
        pic->q_scale_type = (int)br.get(1);
        pic->intra_vlc_format = (int)br.get(1);
        pic->alternate_scan = (int)br.get(1);
        pic->have_coding_extension = true;
        return true;
    }
    case 3: {   // quant_matrix_extension (mp2v_hdr.cpp:133-152); applies to the picture that follows its header
        if (!pic) return false;
        if (br.get1()) { read_matrix(br, pic->tx[0]); memcpy(pic->tx[2], pic->tx[0], 64); }
        if (br.get1()) { read_matrix(br, pic->tx[1]); memcpy(pic->tx[3], pic->tx[1], 64); }
        if (br.get1()) read_matrix(br, pic->tx[2]);
        if (br.get1()) read_matrix(br, pic->tx[3]);
        // keep them in force for later pictures of the sequence as well
        memcpy(seq.intra_matrix, pic->tx[0], 64);
        memcpy(seq.non_intra_matrix, pic->tx[1], 64);
        memcpy(seq.chroma_intra_matrix, pic->tx[2], 64);
        memcpy(seq.chroma_non_intra_matrix, pic->tx[3], 64);
        return true;
    }
    default:
        return true;   // display / copyright / scalable extensions carry nothing the reconstruction needs
    }
}

void build_picture_matrices(const picture_info_t& pic, uint8_t W[4][64]) {
    for (int k = 0; k < 4; k++) build_scan_indexed_matrix(pic.tx[k], pic.alternate_scan, W[k]);
}

// ------------------------------------------------------------------------------------------------
namespace {

mp2v_coef_t* thread_scratch(size_t records) {
    static thread_local std::vector<mp2v_coef_t> scratch;
    if (scratch.size() < records) scratch.resize(records);
    return scratch.data();
}

}  // namespace

slice_syntax_t make_slice_syntax(const sequence_info_t& seq, const picture_info_t& pic, int mbw, int mbh) {
    slice_syntax_t sx{};
    sx.picture_coding_type = pic.picture_coding_type;
    for (int s = 0; s < 2; s++) for (int t = 0; t < 2; t++) sx.f_code[s][t] = pic.f_code[s][t];
    sx.intra_dc_precision = pic.intra_dc_precision;
    sx.q_scale_type = pic.q_scale_type;
    sx.intra_vlc_format = pic.intra_vlc_format;
    sx.chroma_format = seq.chroma_format;
    sx.vertical_size = seq.vertical_size;
    sx.mbw = mbw; sx.mbh = mbh;
    sx.field_dct_syntax = pic.frame_pred_frame_dct ? 0 : 1;
    return sx;
}

bool picture_in_envelope(const picture_info_t& pic) {
    return pic.picture_structure == 3 && !pic.concealment_motion_vectors;      // (frame_pred_frame_dct = 0: field DCT is decoded, field prediction is a slice error)
}

// Host side of parse_slice_core: the slice is parsed into the calling thread's scratch row, then
// appended to the picture arena with one atomic add + memcpy and its offsets rebased.
slice_result_t parse_slice(const uint8_t* payload, int slice_start_code, const sequence_info_t& seq, const picture_info_t& pic,
                           int mbw, int mbh, mp2v_mb_info_t* mb, coef_arena_t& arena) {
    slice_result_t res;
    auto fail = [&](const char* why) { res.ok = false; res.error = why; return res; };
    if (!picture_in_envelope(pic))
        return fail("only frame pictures without concealment vectors are supported (the reference's envelope)");
    const slice_syntax_t sx = make_slice_syntax(seq, pic, mbw, mbh);
    const int nblk = sx.chroma_format == 1 ? 6 : sx.chroma_format == 2 ? 8 : 12;
    mp2v_coef_t* const scratch = thread_scratch((size_t)mbw * nblk * 64u);
    uint32_t total = 0;
    int first_mbx = 0, last_mbx = -1, mb_row = 0;
    const int err = parse_slice_core<false>(payload, slice_start_code, sx, vlc_decode_tables(), mb, scratch, 0u, &total, &first_mbx, &last_mbx, &mb_row);
    if (err != SLICE_OK) return fail(slice_error_string(err));
    res.mbs = last_mbx - first_mbx + 1;
    const uint32_t base = arena.next.fetch_add(total, std::memory_order_relaxed);
    if ((uint64_t)base + total > arena.capacity) { arena.overflow.store(true, std::memory_order_relaxed); return fail("coefficient arena exhausted"); }
    memcpy(arena.base + base, scratch, (size_t)total * sizeof(mp2v_coef_t));
    mp2v_mb_info_t* row = mb + (size_t)mb_row * mbw;
    for (int x = first_mbx; x <= last_mbx; x++) row[x].coef_off += base;
    return res;
}

}  // namespace mp2v
