// Slice-layer syntax parser shared by the host (mp2v_parser.cpp) and the device (csrc/vlc_kernel.cu).
//
// One function, parse_slice_core(), walks one slice (ISO/IEC 13818-2 6.2.4-6.2.6, frame pictures with
// frame prediction: the reference's envelope) and writes reconstruction records (include/mp2v_recon.h):
// a macroblock record per macroblock of the slice's row and the coefficient records behind `out`.
// It is the re-cut of the reference's parse_macroblock_template / parse_block (mb_decoder.cpp:74-155,
// 521-641): everything that is serial inside a slice -- VLC decoding, intra DC prediction
// (mb_decoder.cpp:46-72), motion vector prediction (:447-519, 580-604), quantiser_scale tracking
// (:555-563), skipped-macroblock resolution (:541-550) -- and nothing that touches pixels.
// The same source compiled for both sides is what keeps the two parsers bit-identical.
#pragma once
#include <cstdint>
#include <cstring>

#include "bitreader.h"
#include "mp2v_recon.h"
#include "vlc_decode.h"

namespace mp2v {

// what the slice layer needs of the picture / sequence headers (POD: also passed to the GPU)
struct slice_syntax_t {
    int32_t picture_coding_type;      // 1 I, 2 P, 3 B
    int32_t f_code[2][2];
    int32_t intra_dc_precision;
    int32_t q_scale_type;
    int32_t intra_vlc_format;
    int32_t chroma_format;            // 1, 2, 3
    int32_t vertical_size;            // > 2800: slices carry slice_vertical_position_extension
    int32_t mbw, mbh;
    int32_t field_dct_syntax;         // frame_pred_frame_dct = 0: macroblock_modes carry frame_motion_type and dct_type
};

enum slice_error_t {
    SLICE_OK = 0, SLICE_ERR_ROW, SLICE_ERR_MBA, SLICE_ERR_ADDRESS, SLICE_ERR_SKIP_IN_I, SLICE_ERR_MBTYPE, SLICE_ERR_FCODE,
    SLICE_ERR_MOTION, SLICE_ERR_CBP, SLICE_ERR_COEF, SLICE_ERR_CAPACITY, SLICE_ERR_MV_RANGE, SLICE_ERR_MOTION_TYPE
};

inline const char* slice_error_string(int e) {
    switch (e) {
        case SLICE_OK: return "ok";
        case SLICE_ERR_ROW: return "slice row outside the picture";
        case SLICE_ERR_MBA: return "bad macroblock_address_increment";
        case SLICE_ERR_ADDRESS: return "macroblock address past the end of the row";
        case SLICE_ERR_SKIP_IN_I: return "skipped macroblock in an I picture";
        case SLICE_ERR_MBTYPE: return "bad macroblock_type";
        case SLICE_ERR_FCODE: return "f_code out of range";
        case SLICE_ERR_MOTION: return "bad motion_code";
        case SLICE_ERR_CBP: return "bad coded_block_pattern";
        case SLICE_ERR_COEF: return "bad DCT coefficient syntax";
        case SLICE_ERR_CAPACITY: return "coefficient arena exhausted";
        case SLICE_ERR_MV_RANGE: return "motion vector points outside the reference frame";
        case SLICE_ERR_MOTION_TYPE: return "field prediction / dual prime in a frame picture (only frame-based prediction is supported)";
        default: return "slice parse error";
    }
}

#if defined(__CUDA_ARCH__)
#define MP2V_UNROLL2 _Pragma("unroll 2")
#elif defined(__CUDACC__)
#define MP2V_UNROLL2                       /* host pass of a .cu file: the host copy is never called from there */
#else
#define MP2V_UNROLL2 _Pragma("GCC unroll 2")
#endif

// (the two-trip loops over prediction direction / vector component: fully unrolled so that pmv[][] , mv[][] and
// f_code[][] are indexed by constants and live in registers -- the device compiler otherwise parks them in local memory)
#define MP2V_UNROLL_ALL MP2V_UNROLL2

#if defined(__CUDACC__)
#define MP2V_HDI __host__ __device__ __forceinline__
#else
#define MP2V_HDI inline __attribute__((always_inline))
#endif

MP2V_HDI uint32_t reverse_bits(uint32_t x, int n) {           // the low n bits of x, mirrored
#ifdef __CUDA_ARCH__
    return __brev(x) >> (32 - n);
#else
    uint32_t r = 0;
    for (int i = 0; i < n; i++) r |= ((x >> i) & 1u) << (n - 1 - i);
    return r;
#endif
}
MP2V_HDI int lowest_bit(uint32_t x) {                           // index of the lowest set bit (x != 0)
#ifdef __CUDA_ARCH__
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

MP2V_HDI int quantiser_scale_of(int code, int q_scale_type) {   // decoder.cpp:140-145, mb_decoder.cpp:555-563
    if (!q_scale_type) return code << 1;
    if (code < 9) return code;
    if (code < 17) return (code - 4) << 1;
    if (code < 25) return (code - 10) << 2;
    return (code - 17) << 3;
}

// Everything below works on the caller's LOCAL bit reader and write cursor, passed by reference into
// always-inlined helpers, so that both live in registers for the whole slice.

// one motion vector component, mb_decoder.cpp:447-503
MP2V_HDI bool decode_mv_component(bitreader_t& br, const vlc_decode_tables_t& T, int f_code, int& pmv, int& out) {
    br.refill();
    const vlc_entry_t e = T.motion.look(br.peek(10));
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

// intra DC predictors (mb_decoder.cpp:46-72) as three scalars: a block's component is only known at run time, and an
// array indexed by it would live in local memory on the device
struct dc_pred_t {
    uint32_t y, cb, cr;
    MP2V_HDI void reset(uint32_t v) { y = cb = cr = v; }
    MP2V_HDI uint32_t get(int comp) const { return comp == 0 ? y : comp == 1 ? cb : cr; }
    MP2V_HDI void set(int comp, uint32_t v) { if (comp == 0) y = v; else if (comp == 1) cb = v; else cr = v; }
};

constexpr int kDevTableWords = (2 << kFastBits) + (2 << coef_vlc_t::kLongBits);

#ifdef __CUDA_ARCH__
__device__ __forceinline__ uint32_t lds_word(uint32_t shared_base, uint32_t index) {      // word `index` of a table in shared memory
    uint32_t v;
    asm("{\n\t.reg .u32 a;\n\tmad.lo.u32 a, %2, 4, %1;\n\tld.shared.u32 %0, [a];\n\t}" : "=r"(v) : "r"(shared_base), "r"(index));
    return v;
}
#endif

// one block: DC (intra) + run/level list; returns false on a syntax error
// dev_fast: shared-window address of the device parser's copy of {b14.gpu_fast, b15.gpu_fast, b14.gpu_long, b15.gpu_long}
// (kDevTableWords words; unused on the host)
MP2V_HDI bool parse_block(bitreader_t& br, mp2v_coef_t*& out, const vlc_decode_tables_t& T, const slice_syntax_t& sx,
                          dc_pred_t& dc_pred, int b, bool intra, uint32_t mb_bits, uint32_t dev_fast) {
    const uint32_t blk_bits = ((uint32_t)b << 22) | mb_bits;      // block index + the macroblock column tag of every record
    int i = 0;
    const coef_vlc_t* table = &T.b14;
    br.refill();
    if (intra) {
        const int comp = b < 4 ? 0 : 1 + (b & 1);
        int diff;
        const dc_fast_t f = fetch_entry(&T.dc_fast[comp ? 1 : 0][br.peek(kDcFastBits)]);
        if (f.len != 0) { br.skip(f.len); diff = f.diff; }
        else {
            const vlc_entry_t e = T.dcsize[comp ? 1 : 0].look(br.peek(10));
            if (!e.len) return false;
            br.skip(e.len);
            const int v = (int)br.peek(e.val);                 // here size >= 1 (size 0 always fits the fast table)
            br.skip(e.val);
            const int half = 1 << (e.val - 1);
            diff = v >= half ? v : v + 1 - 2 * half;           // mb_decoder.cpp:59-68
        }
        const uint32_t pred = (dc_pred.get(comp) + (uint32_t)diff) & 0xffffu;
        dc_pred.set(comp, pred);
        const int16_t dc = (int16_t)(uint16_t)(pred << (3 - sx.intra_dc_precision));
        *out++ = MP2V_COEF(dc, 0, b, MP2V_COEF_RAW) | mb_bits;
        i = 1;
        if (sx.intra_vlc_format) table = &T.b15;
        br.refill();
    } else if (br.peek(1)) {                                   // first coefficient "1s" (mb_decoder.cpp:79-88)
        const int neg = (int)br.peek(2) & 1;
        br.skip(2);
        *out++ = MP2V_COEF(neg ? -1 : 1, 0, b, MP2V_COEF_FIRST) | mb_bits;
        i = 1;
    }
    // Symbol loop, one source for both sides.  A fast symbol is: look up one pre-assembled word (the device keeps the
    // tables in shared memory), add it to the running {position, block} word q -- that IS the record -- store, keep
    // its upper half as the next q.  q = ((i - 1) << 16 | block << 22 | macroblock tag) modulo 2^32: a position past
    // 63 carries into the block field and stays there, so the range check is ONE compare of q at the end of the block;
    // what bounds the writes meanwhile is the record count (a block has at most 64 coefficients and every symbol is
    // one).  Fast symbols are at most kFastBits = 11 bits long, so three of them fit the 33 bits a refill guarantees
    // on the device (the host's 56 bits hold them too): one refill check per three symbols.
    {
        const bool b15 = table == &T.b15;
#ifdef __CUDA_ARCH__
        const uint32_t fast = dev_fast + (b15 ? (4u << kFastBits) : 0u);
        const uint32_t long_codes = dev_fast + (8u << kFastBits) + (b15 ? (4u << coef_vlc_t::kLongBits) : 0u);
#define MP2V_FAST_ENTRY(idx) lds_word(fast, (idx))
#define MP2V_LONG_ENTRY(idx) lds_word(long_codes, (idx))
#else
        (void)dev_fast;
        const uint32_t* const fast = b15 ? T.b15.gpu_fast : T.b14.gpu_fast;
        const uint32_t* const long_codes = b15 ? T.b15.gpu_long : T.b14.gpu_long;
#define MP2V_FAST_ENTRY(idx) fast[(idx)]
#define MP2V_LONG_ENTRY(idx) long_codes[(idx)]
#endif
        uint32_t q = (((uint32_t)i << 16) | blk_bits) - 0x10000u;
        mp2v_coef_t* const o0 = out;
        uint32_t n = 0;                                        // records stored by this loop; 32-bit index, not a 64-bit pointer bump
        const uint32_t n_max = 64u - (uint32_t)i;
        bool ok = false;
        for (;;) {
            // tight loop over fast symbols only (its own loop so that the rare paths below do not shape its code)
            uint32_t e;
            for (;;) {
                br.refill();
#define MP2V_FAST_SYMBOL                                                                   \
                e = MP2V_FAST_ENTRY(br.peek(kFastBits));                                   \
                if ((int32_t)e < 0 || n == n_max) break;   /* not fast, or a 65th coefficient */ \
                br.skip((int)(e >> 24));                                                   \
                { const uint32_t rec = q + (e & 0x007fffffu); o0[n++] = rec; q = rec & 0xffff0000u; }
                MP2V_FAST_SYMBOL
                MP2V_FAST_SYMBOL
                MP2V_FAST_SYMBOL
#undef MP2V_FAST_SYMBOL
            }
            if ((int32_t)e >= 0) break;                        // count error inside the tight loop
            if (e & 0x40000000u) {                             // end of block
                br.skip((int)((e >> 24) & 15u));
                ok = true;
                break;
            }
            // the rare symbols: an escape (000001, 6-bit run, 12-bit two's complement level) or one of the long codes
            br.refill();                                       // up to 22 bits went since the loop's refill; an escape is 24
            const uint32_t w = br.peek(24);
            uint32_t inc, len;
            if ((w >> 18) == 1u) {
                inc = ((((w >> 12) & 63u) + 1u) << 16) | ((((w & 0xfffu) ^ 0x800u) - 0x800u) & 0xffffu);
                len = 24u;
            } else if ((w >> (24 - coef_vlc_t::kLongZeros)) == 0u) {
                const uint32_t l = MP2V_LONG_ENTRY((w >> (24 - coef_vlc_t::kLongZeros - coef_vlc_t::kLongBits)) & ((1u << coef_vlc_t::kLongBits) - 1u));
                if ((int32_t)l < 0) break;                     // no such code
                inc = l & 0x007fffffu;
                len = l >> 24;
            } else {
                break;
            }
            if (n == n_max) break;
            br.skip((int)len);
            const uint32_t rec = q + inc;
            o0[n++] = rec;
            q = rec & 0xffff0000u;
        }
#undef MP2V_FAST_ENTRY
#undef MP2V_LONG_ENTRY
        out = o0 + n;
        return ok && ((q ^ blk_bits) >> 22) == 0;              // i + run never passed 63
    }
}

// the reference does not clamp vectors: one that leaves the frame is an error here
MP2V_HDI bool mv_inside(int mbx, int mby, int mvx, int mvy, int width, int height) {
    const int x0 = mbx * 16 + (mvx >> 1), y0 = mby * 16 + (mvy >> 1);
    return x0 >= 0 && y0 >= 0 && x0 + 16 + (mvx & 1) <= width && y0 + 16 + (mvy & 1) <= height;
}

// Parse one slice.  payload = first byte after the 4-byte start code.  Macroblock records go to
// mb[row * mbw + x]; coefficient records to out_base[0 ...] with coef_off = coef_off_base + index.
// Returns a slice_error_t; *n_out = coefficient records written, [*first_mbx, *last_mbx] = the
// macroblocks of the row this slice wrote.  CHECK_MV: reject vectors that leave the frame while
// parsing (the device-side parser; the host path validates whole pictures in mp2v_recon_precheck).
template <bool CHECK_MV>
MP2V_HDI int parse_slice_core(const uint8_t* payload, int slice_start_code, const slice_syntax_t& sx, const vlc_decode_tables_t& T,
                              mp2v_mb_info_t* mb, mp2v_coef_t* out_base, uint32_t coef_off_base,
                              uint32_t* n_out, int* first_mbx_out, int* last_mbx_out, int* mb_row_out, uint32_t dev_fast = 0) {
    const int cf = sx.chroma_format, mbw = sx.mbw;
    const int nblk = cf == 1 ? 6 : cf == 2 ? 8 : 12;
    mp2v_coef_t* out = out_base;
    bitreader_t br(payload);
    int pmv[2][2] = {{0, 0}, {0, 0}};
    dc_pred_t dc_pred;
    const uint32_t dc_reset = 1u << (sx.intra_dc_precision + 7);
    dc_pred.reset(dc_reset);
    int mb_row = slice_start_code - 1;
    if (sx.vertical_size > 2800) mb_row += (int)br.get(3) << 7;        // slice_vertical_position_extension
    *n_out = 0; *first_mbx_out = 0; *last_mbx_out = -1; *mb_row_out = mb_row;
    if (mb_row < 0 || mb_row >= sx.mbh) return SLICE_ERR_ROW;
    int qscale = quantiser_scale_of((int)br.get(5), sx.q_scale_type);
    if (br.get1()) {                                                    // intra_slice_flag (mp2v_hdr.h:352-360)
        br.get(8);
        while (br.get1()) br.get(8);
    }
    const int pct = sx.picture_coding_type;
    mp2v_mb_info_t* row = mb + (size_t)mb_row * mbw;
    uint32_t prev_dirs = 0;
    int mbx = -1, first_mbx = 0, done_mbx = -1;      // done_mbx: last macroblock whose record is complete
    bool first = true;
    int err = SLICE_OK;
    do {
        // ---- macroblock_address_increment (+ escapes)
        int inc = 0;
        for (;;) {
            br.refill();
            const vlc_entry_t e = T.mba.look(br.peek(11));
            if (!e.len) { err = SLICE_ERR_MBA; break; }
            br.skip(e.len);
            if (e.val) { inc += e.val; break; }
            inc += 33;
        }
        if (err) break;
        // first macroblock of a slice: the increment is its column (6.3.16); later ones: inc-1 skipped
        const int target = first ? inc - 1 : mbx + inc;
        if (target >= mbw) { err = SLICE_ERR_ADDRESS; break; }
        const int skipped = first ? 0 : inc - 1;
        if (first) { mbx = target - 1; first_mbx = target; done_mbx = target - 1; first = false; }
        // ---- skipped macroblocks (mb_decoder.cpp:541-550)
        if (skipped > 0) {
            if (pct == 1) { err = SLICE_ERR_SKIP_IN_I; break; }
            if (pct == 2) { pmv[0][0] = pmv[0][1] = pmv[1][0] = pmv[1][1] = 0; }
            uint32_t dirs = pct == 2 ? MP2V_MB_FWD : prev_dirs;
            if (!dirs) dirs = MP2V_MB_FWD;                              // after an intra macroblock the reference predicts forward
            for (int k = 0; k < skipped && !err; k++) {
                if (CHECK_MV)
                    for (int s = 0; s < 2; s++)
                        if ((dirs & (s ? MP2V_MB_BWD : MP2V_MB_FWD)) && !mv_inside(mbx + 1, mb_row, pmv[s][0], pmv[s][1], mbw * 16, sx.mbh * 16)) err = SLICE_ERR_MV_RANGE;
                if (err) break;
                mp2v_mb_info_t& r = row[++mbx];
                r.coef_off = coef_off_base + (uint32_t)(out - out_base);
                r.bits = MP2V_MB_BITS(0, qscale, 0, dirs);
                for (int s = 0; s < 2; s++) for (int t = 0; t < 2; t++)
                    r.mv[s][t] = (int16_t)((dirs & (s ? MP2V_MB_BWD : MP2V_MB_FWD)) ? pmv[s][t] : 0);
                done_mbx = mbx;
            }
            if (err) break;
            dc_pred.reset(dc_reset);
        }
        mp2v_mb_info_t& r = row[++mbx];
        // ---- macroblock_type (no refill: the last increment code took at most 11 of the 33 bits, type, modes and quantiser 14 more)
        const vlc_entry_t te = T.mbtype[pct].look(br.peek(6));
        if (!te.len) { err = SLICE_ERR_MBTYPE; break; }
        br.skip(te.len);
        const int type = te.val;
        const bool intra = type & 0x02, fwd = type & 0x10, bwd = type & 0x08, pattern = type & 0x04;
        // frame pictures with frame_pred_frame_dct = 0 (parse_modes, mb_decoder.cpp:347-360): frame_motion_type, dct_type
        uint32_t field_dct = 0;
        if (sx.field_dct_syntax) {
            if (fwd || bwd) {
                const uint32_t fmt = br.peek(2);
                br.skip(2);
                if (fmt != 2u) { err = SLICE_ERR_MOTION_TYPE; break; }       // 1 field-based, 3 dual prime (the reference: TODO)
            }
            if (intra || pattern) { field_dct = br.peek(1) ? MP2V_MB_FIELD_DCT : 0u; br.skip(1); }
        }
        if (type & 0x20) { qscale = quantiser_scale_of((int)br.peek(5), sx.q_scale_type); br.skip(5); }
        // ---- motion vectors (frame prediction: one vector per direction)
        int mv[2][2] = {{0, 0}, {0, 0}};
        MP2V_UNROLL_ALL
        for (int s = 0; s < 2; s++) {
            if (err || !(s ? bwd : fwd)) continue;
            MP2V_UNROLL_ALL
            for (int t = 0; t < 2; t++) {
                if (err) continue;
                const int fc = sx.f_code[s][t];
                if (fc < 1 || fc > 9) { err = SLICE_ERR_FCODE; continue; }
                if (!decode_mv_component(br, T, fc, pmv[s][t], mv[s][t])) err = SLICE_ERR_MOTION;
            }
        }
        if (err) break;
        if (intra || (pct == 2 && !fwd)) { pmv[0][0] = pmv[0][1] = pmv[1][0] = pmv[1][1] = 0; }   // mb_decoder.cpp:599-603
        if (!intra) dc_pred.reset(dc_reset);                                // mb_decoder.cpp:623-626
        // ---- coded_block_pattern
        uint32_t cbp = 0;
        if (intra) cbp = (1u << nblk) - 1u;
        else if (pattern) {
            br.refill();
            const vlc_entry_t ce = T.cbp.look(br.peek(9));
            if (!ce.len) { err = SLICE_ERR_CBP; break; }
            br.skip(ce.len);
            cbp = reverse_bits((uint32_t)ce.val, 6);                                           // mb_decoder.cpp:435-436: block 0 is the code's MSB
            if (cf == 2) { const uint32_t x = br.peek(2); br.skip(2); cbp |= ((x >> 1) & 1u) << 6 | (x & 1u) << 7; }
            if (cf == 3) { const uint32_t x = br.peek(6); br.skip(6); cbp |= reverse_bits(x, 6) << 6; }
        }
        // ---- blocks
        const uint32_t off = (uint32_t)(out - out_base);
        // (coded blocks by the set bits of cbp, not by block number: slices that share a warp on the device then walk
        // their k-th coded block together whichever block that is)
        for (uint32_t left = cbp; left; left &= left - 1u)
            if (!parse_block(br, out, T, sx, dc_pred, lowest_bit(left), intra, MP2V_COEF_MB(mbx), dev_fast)) { err = SLICE_ERR_COEF; break; }
        if (err) break;
        uint32_t flags = 0;
        if (intra) flags = MP2V_MB_INTRA;
        else {
            if (fwd) flags |= MP2V_MB_FWD;
            if (bwd) flags |= MP2V_MB_BWD;
            if (!flags) flags = MP2V_MB_FWD;        // P picture "no MC": forward prediction with a zero vector (mb_decoder.cpp:329-338)
            if (CHECK_MV)
                for (int s = 0; s < 2; s++)
                    if ((flags & (s ? MP2V_MB_BWD : MP2V_MB_FWD)) && !mv_inside(mbx, mb_row, mv[s][0], mv[s][1], mbw * 16, sx.mbh * 16)) err = SLICE_ERR_MV_RANGE;
            if (err) break;
        }
        r.coef_off = (coef_off_base + off) | field_dct;
        r.bits = MP2V_MB_BITS((uint32_t)(out - out_base) - off, qscale, cbp, flags);
        for (int s = 0; s < 2; s++) for (int t = 0; t < 2; t++) r.mv[s][t] = (int16_t)(intra ? 0 : mv[s][t]);
        prev_dirs = flags & (MP2V_MB_FWD | MP2V_MB_BWD);
        done_mbx = mbx;
        br.refill();
    } while (br.peek(23) != 0 && mbx < mbw - 1);
    // trailing macroblocks of the row that the slice did not code keep the caller's defaults
    *n_out = (uint32_t)(out - out_base);
    *first_mbx_out = first_mbx;
    *last_mbx_out = done_mbx;
    return err;
}

}  // namespace mp2v
