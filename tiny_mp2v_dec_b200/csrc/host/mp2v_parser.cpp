// Header + slice parser emitting reconstruction records -- see mp2v_parser.h.
// Syntax per ISO/IEC 13818-2 6.2; behaviour checked against the reference's parsers
// (src/core/mp2v_hdr.cpp, mb_decoder.cpp) through tests/test_parser.py and the end-to-end parity tests.
#include "mp2v_parser.h"

#include <cstring>
#include <vector>

#include "bitreader.h"
#include "scan_tables.h"
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

inline int quantiser_scale(int code, int q_scale_type) {   // decoder.cpp:140-145, mb_decoder.cpp:555-563
    if (!q_scale_type) return code << 1;
    if (code < 9) return code;
    if (code < 17) return (code - 4) << 1;
    if (code < 25) return (code - 10) << 2;
    return (code - 17) << 3;
}

}  // namespace

sequence_info_t::sequence_info_t() {
    const scan_tables_t& t = scan_tables();
    for (int i = 0; i < 64; i++) { intra_matrix[i] = kDefaultIntraRaster[t.shuffle[0][i]]; non_intra_matrix[i] = 16; }
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
    memcpy(pic.tx[0], seq.intra_matrix, 64); memcpy(pic.tx[2], seq.intra_matrix, 64);
    memcpy(pic.tx[1], seq.non_intra_matrix, 64); memcpy(pic.tx[3], seq.non_intra_matrix, 64);
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

#define MP2V_INLINE inline __attribute__((always_inline))

// Everything below works on the slice parser's LOCAL bit reader and write cursor, passed by reference
// into always-inlined helpers, so that both live in registers for the whole slice.

// one motion vector component, mb_decoder.cpp:447-503
MP2V_INLINE bool decode_mv_component(bitreader_t& br, const vlc_decode_tables_t& T, int f_code, int& pmv, int& out) {
    br.refill();
    const vlc_entry_t& e = T.motion.look(br.peek(10));
    if (!e.len) return false;
    br.skip(e.len);
    int delta = 0;
    if (e.val) {
        const int neg = (int)br.peek(1);
        br.skip(1);
        const int r_size = f_code - 1;
        delta = e.val;
        if (r_size) { delta = ((e.val - 1) << r_size) + (int)br.peek(r_size) + 1; br.skip(r_size); }
        if (neg) delta = -delta;
    }
    const int f16 = 16 << (f_code - 1);
    int v = pmv + delta;
    if (v < -f16) v += 2 * f16;
    if (v > f16 - 1) v -= 2 * f16;
    pmv = v; out = v;
    return true;
}

// one block: DC (intra) + run/level list; returns false on a syntax error
MP2V_INLINE bool parse_block(bitreader_t& br, mp2v_coef_t*& out, const vlc_decode_tables_t& T, const picture_info_t& pic,
                             uint16_t (&dc_pred)[3], int b, bool intra) {
    const uint32_t blk_bits = (uint32_t)b << 22;
    int i = 0;
    const coef_vlc_t* table = &T.b14;
    br.refill();
    if (intra) {
        const int comp = b < 4 ? 0 : 1 + (b & 1);
        int diff;
        const dc_fast_t f = T.dc_fast[comp ? 1 : 0][br.peek(kDcFastBits)];
        if (__builtin_expect(f.len != 0, 1)) { br.skip(f.len); diff = f.diff; }
        else {
            const vlc_entry_t& e = T.dcsize[comp ? 1 : 0].look(br.peek(10));
            if (!e.len) return false;
            br.skip(e.len);
            const int v = (int)br.peek(e.val);                 // here size >= 1 (size 0 always fits the fast table)
            br.skip(e.val);
            const int half = 1 << (e.val - 1);
            diff = v >= half ? v : v + 1 - 2 * half;           // mb_decoder.cpp:59-68
        }
        dc_pred[comp] = (uint16_t)(dc_pred[comp] + diff);
        const int16_t dc = (int16_t)(uint16_t)((uint32_t)dc_pred[comp] << (3 - pic.intra_dc_precision));
        *out++ = MP2V_COEF(dc, 0, b, MP2V_COEF_RAW);
        i = 1;
        if (pic.intra_vlc_format) table = &T.b15;
        br.refill();
    } else if (br.peek(1)) {                                   // first coefficient "1s" (mb_decoder.cpp:79-88)
        const int neg = (int)br.peek(2) & 1;
        br.skip(2);
        *out++ = MP2V_COEF(neg ? -1 : 1, 0, b, MP2V_COEF_FIRST);
        i = 1;
    }
    const coef_fast_t* fast = table->fast;
    for (;;) {
        // one refill (>= 56 bits) covers two symbols of any kind (escape = 24 bits); the first round
        // reuses the refill above (at most 22 bits were consumed since)
#pragma GCC unroll 2
        for (int rep = 0; rep < 2; rep++) {
            const coef_fast_t f = fast[br.peek(kFastBits)];
            int run, level;
            if (__builtin_expect(f.run < kFastEob, 1)) {
                br.skip(f.len);
                run = f.run; level = f.level;
            } else if (f.run == kFastEob) {
                br.skip(f.len);
                return true;
            } else {
                const coef_entry_t& e = table->look(br.peek(17));
                if (e.level > 0) {
                    br.skip(e.len);
                    const int neg = (int)br.peek(1);
                    br.skip(1);
                    run = e.run;
                    level = (e.level ^ -neg) + neg;
                } else if (e.level == kCoefEob && e.len) {
                    br.skip(e.len);
                    return true;
                } else if (e.level == kCoefEsc && e.len) {     // 6-bit run, 12-bit two's complement level
                    br.skip(6);
                    run = (int)br.peek(6); br.skip(6);
                    level = ((int)br.peek(12) ^ 0x800) - 0x800; br.skip(12);
                } else {
                    return false;
                }
            }
            i += run;
            if (__builtin_expect(i > 63, 0)) return false;
            *out++ = (uint32_t)(uint16_t)level | ((uint32_t)i << 16) | blk_bits;
            i++;
        }
        br.refill();
    }
}

}  // namespace

slice_result_t parse_slice(const uint8_t* payload, int slice_start_code, const sequence_info_t& seq, const picture_info_t& pic,
                           int mbw, int mbh, mp2v_mb_info_t* mb, coef_arena_t& arena) {
    slice_result_t res;
    auto fail = [&](const char* why) { res.ok = false; res.error = why; return res; };
    if (pic.picture_structure != 3 || !pic.frame_pred_frame_dct || pic.concealment_motion_vectors)
        return fail("only progressive frame pictures with frame prediction are supported (the reference's envelope)");
    const int cf = seq.chroma_format;
    const int nblk = cf == 1 ? 6 : cf == 2 ? 8 : 12;
    mp2v_coef_t* const scratch = thread_scratch((size_t)mbw * nblk * 64u);
    mp2v_coef_t* out = scratch;
    const vlc_decode_tables_t& T = vlc_decode_tables();
    bitreader_t br(payload);
    int pmv[2][2];
    uint16_t dc_pred[3];
    auto reset_dc = [&] { for (auto& d : dc_pred) d = (uint16_t)(1u << (pic.intra_dc_precision + 7)); };
    int mb_row = slice_start_code - 1;
    if (seq.vertical_size > 2800) mb_row += (int)br.get(3) << 7;        // slice_vertical_position_extension
    if (mb_row < 0 || mb_row >= mbh) return fail("slice row outside the picture");
    int qscale = quantiser_scale((int)br.get(5), pic.q_scale_type);
    if (br.get1()) {                                                    // intra_slice_flag (mp2v_hdr.h:352-360)
        br.get(8);
        while (br.get1()) br.get(8);
    }
    memset(pmv, 0, sizeof(pmv));
    reset_dc();
    const int pct = pic.picture_coding_type;
    mp2v_mb_info_t* row = mb + (size_t)mb_row * mbw;
    uint32_t prev_dirs = 0;
    int mbx = -1, first_mbx = 0;
    bool first = true;
    do {
        // ---- macroblock_address_increment (+ escapes)
        int inc = 0;
        for (;;) {
            br.refill();
            const vlc_entry_t& e = T.mba.look(br.peek(11));
            if (!e.len) return fail("bad macroblock_address_increment");
            br.skip(e.len);
            if (e.val) { inc += e.val; break; }
            inc += 33;
        }
        // first macroblock of a slice: the increment is its column (6.3.16); later ones: inc-1 skipped
        const int target = first ? inc - 1 : mbx + inc;
        if (target >= mbw) return fail("macroblock address past the end of the row");
        const int skipped = first ? 0 : inc - 1;
        if (first) { mbx = target - 1; first_mbx = target; first = false; }
        // ---- skipped macroblocks (mb_decoder.cpp:541-550)
        if (skipped > 0) {
            if (pct == 1) return fail("skipped macroblock in an I picture");
            if (pct == 2) memset(pmv, 0, sizeof(pmv));
            uint32_t dirs = pct == 2 ? MP2V_MB_FWD : prev_dirs;
            if (!dirs) dirs = MP2V_MB_FWD;                              // after an intra macroblock the reference predicts forward
            for (int k = 0; k < skipped; k++) {
                mp2v_mb_info_t& r = row[++mbx];
                r.coef_off = (uint32_t)(out - scratch);
                r.bits = MP2V_MB_BITS(0, qscale, 0, dirs);
                for (int s = 0; s < 2; s++) for (int t = 0; t < 2; t++)
                    r.mv[s][t] = (int16_t)((dirs & (s ? MP2V_MB_BWD : MP2V_MB_FWD)) ? pmv[s][t] : 0);
                res.mbs++;
            }
            reset_dc();
        }
        mp2v_mb_info_t& r = row[++mbx];
        // ---- macroblock_type
        br.refill();
        const vlc_entry_t& te = T.mbtype[pct].look(br.peek(6));
        if (!te.len) return fail("bad macroblock_type");
        br.skip(te.len);
        const int type = te.val;
        const bool intra = type & 0x02, fwd = type & 0x10, bwd = type & 0x08, pattern = type & 0x04;
        if (type & 0x20) qscale = quantiser_scale((int)br.peek(5), pic.q_scale_type), br.skip(5);
        // ---- motion vectors (frame prediction: one vector per direction)
        int mv[2][2] = {{0, 0}, {0, 0}};
        for (int s = 0; s < 2; s++) {
            if (!(s ? bwd : fwd)) continue;
            for (int t = 0; t < 2; t++) {
                const int fc = pic.f_code[s][t];
                if (fc < 1 || fc > 9) return fail("f_code out of range");
                if (!decode_mv_component(br, T, fc, pmv[s][t], mv[s][t])) return fail("bad motion_code");
            }
        }
        if (intra || (pct == 2 && !fwd)) memset(pmv, 0, sizeof(pmv));   // mb_decoder.cpp:599-603
        if (!intra) reset_dc();                                           // mb_decoder.cpp:623-626
        // ---- coded_block_pattern
        uint32_t cbp = 0;
        if (intra) cbp = (1u << nblk) - 1u;
        else if (pattern) {
            br.refill();
            const vlc_entry_t& ce = T.cbp.look(br.peek(9));
            if (!ce.len) return fail("bad coded_block_pattern");
            br.skip(ce.len);
            for (int i = 0; i < 6; i++) if (ce.val & (1 << (5 - i))) cbp |= 1u << i;          // mb_decoder.cpp:435-436
            if (cf == 2) { const uint32_t x = br.peek(2); br.skip(2); cbp |= ((x >> 1) & 1u) << 6 | (x & 1u) << 7; }
            if (cf == 3) { const uint32_t x = br.peek(6); br.skip(6); for (int i = 0; i < 6; i++) cbp |= ((x >> (5 - i)) & 1u) << (6 + i); }
        }
        // ---- blocks
        const uint32_t off = (uint32_t)(out - scratch);
        for (int b = 0; b < nblk; b++)
            if (cbp & (1u << b))
                if (!parse_block(br, out, T, pic, dc_pred, b, intra)) return fail("bad DCT coefficient syntax");
        uint32_t flags = 0;
        if (intra) flags = MP2V_MB_INTRA;
        else {
            if (fwd) flags |= MP2V_MB_FWD;
            if (bwd) flags |= MP2V_MB_BWD;
            if (!flags) flags = MP2V_MB_FWD;        // P picture "no MC": forward prediction with a zero vector (mb_decoder.cpp:329-338)
        }
        r.coef_off = off;
        r.bits = MP2V_MB_BITS((uint32_t)(out - scratch) - off, qscale, cbp, flags);
        for (int s = 0; s < 2; s++) for (int t = 0; t < 2; t++) r.mv[s][t] = (int16_t)(intra ? 0 : mv[s][t]);
        prev_dirs = flags & (MP2V_MB_FWD | MP2V_MB_BWD);
        res.mbs++;
        br.refill();
    } while (br.peek(23) != 0 && mbx < mbw - 1);
    // trailing macroblocks of the row that the slice did not code keep the caller's defaults.
    // append the slice's records to the picture arena and rebase the offsets written above
    const uint32_t total = (uint32_t)(out - scratch);
    const uint32_t base = arena.next.fetch_add(total, std::memory_order_relaxed);
    if ((uint64_t)base + total > arena.capacity) { arena.overflow.store(true, std::memory_order_relaxed); return fail("coefficient arena exhausted"); }
    memcpy(arena.base + base, scratch, (size_t)total * sizeof(mp2v_coef_t));
    for (int x = first_mbx; x <= mbx; x++) row[x].coef_off += base;
    return res;
}

}  // namespace mp2v
