// Synthetic MPEG-2 elementary stream generator -- see streamgen.h.
// Bit layouts follow ISO/IEC 13818-2 6.2 as the reference's parsers read them
// (src/core/mp2v_hdr.cpp:4-152, mp2v_hdr.h:345-363, mb_decoder.cpp:341-641).
#include "streamgen.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "scan_tables.h"
#include "vlc_tables.h"

namespace {

using namespace mp2v;

struct rng_t {   // splitmix64: same sequence on every box
    uint64_t s;
    explicit rng_t(uint64_t seed) : s(seed * 0x9e3779b97f4a7c15ull + 0x1234567ull) {}
    uint64_t next() {
        uint64_t z = (s += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    int below(int n) { return n <= 1 ? 0 : (int)(next() % (uint64_t)n); }
    int range(int lo, int hi) { return lo + below(hi - lo + 1); }   // inclusive
    bool pct(int p) { return below(100) < p; }
};

struct bitwriter_t {
    std::vector<uint8_t> bytes;
    uint64_t acc = 0;
    int n = 0;
    void put(uint32_t v, int len) {
        for (int i = len - 1; i >= 0; i--) {
            acc = (acc << 1) | ((v >> i) & 1u);
            if (++n == 8) { bytes.push_back((uint8_t)acc); acc = 0; n = 0; }
        }
    }
    void put(const char* bits) { for (; *bits; bits++) put(*bits == '1', 1); }
    void align() { while (n) put(0, 1); }
    std::vector<size_t> start_codes;   // byte offsets of the real start codes written
    void start_code(int code) { align(); start_codes.push_back(bytes.size()); put(0, 8); put(0, 8); put(1, 8); put((uint32_t)code, 8); }
};

// forward (encode) view of vlc_tables.h
struct enc_tables_t {
    const char* mba[34] = {};
    const char* mba_escape = nullptr;
    const char* mbtype[4][64] = {};
    const char* cbp[64] = {};
    const char* motion[17] = {};
    const char* dcsize[2][12] = {};
    const char* b14[64][41] = {};
    const char* b15[64][41] = {};
    enc_tables_t() {
        for (int i = 0; i < MP2V_COUNT(kTabMbAddrInc); i++) {
            if (kTabMbAddrInc[i].a) mba[kTabMbAddrInc[i].a] = kTabMbAddrInc[i].bits; else mba_escape = kTabMbAddrInc[i].bits;
        }
        for (int i = 0; i < MP2V_COUNT(kTabMbType); i++) mbtype[kTabMbType[i].b][kTabMbType[i].a] = kTabMbType[i].bits;
        for (int i = 0; i < MP2V_COUNT(kTabCbp); i++) cbp[kTabCbp[i].a] = kTabCbp[i].bits;
        for (int i = 0; i < MP2V_COUNT(kTabMotionCode); i++) motion[kTabMotionCode[i].a] = kTabMotionCode[i].bits;
        for (int i = 0; i < MP2V_COUNT(kTabDcSize); i++) dcsize[kTabDcSize[i].b][kTabDcSize[i].a] = kTabDcSize[i].bits;
        for (int i = 0; i < MP2V_COUNT(kTabCoefB14); i++) b14[kTabCoefB14[i].a][kTabCoefB14[i].b] = kTabCoefB14[i].bits;
        for (int i = 0; i < MP2V_COUNT(kTabCoefB15); i++) b15[kTabCoefB15[i].a][kTabCoefB15[i].b] = kTabCoefB15[i].bits;
    }
};
const enc_tables_t& enc() { static const enc_tables_t t; return t; }

struct picture_t {
    mp2v_pic_params_t params{};
    std::vector<mp2v_mb_info_t> mb;
    std::vector<mp2v_coef_t> coef;
    int display_index = 0, gop = 0, q_scale_type = 0, intra_dc_precision = 0;
    uint8_t tx[4][64] = {};
    int tx_loaded[4] = {};
};

int quantiser_scale_of(int code, int q_scale_type) {   // decoder.cpp:140-145
    if (!q_scale_type) return code << 1;
    if (code < 9) return code;
    if (code < 17) return (code - 4) << 1;
    if (code < 25) return (code - 10) << 2;
    return (code - 17) << 3;
}

static const uint8_t kDefaultIntra[64] = {   // ISO/IEC 13818-2 6.3.11 default intra matrix (raster)
     8, 16, 19, 22, 26, 27, 29, 34, 16, 16, 22, 24, 27, 29, 34, 37, 19, 22, 26, 27, 29, 34, 34, 38,
    22, 22, 26, 27, 29, 34, 37, 40, 22, 26, 27, 29, 32, 35, 40, 48, 26, 27, 29, 32, 35, 40, 48, 58,
    26, 27, 29, 34, 38, 46, 56, 69, 27, 29, 35, 38, 46, 56, 69, 83 };

}  // namespace

struct mp2v_gen {
    mp2v_gen_params_t p{};
    rng_t rng{0};
    bitwriter_t bw;
    std::vector<picture_t> pics;
    std::vector<size_t> gop_off;
    std::string err;
    int mbw = 0, mbh = 0, nblk = 0;

    // ---- per-slice coding state
    int pmv[2][2];
    int dc_pred[3];
    int qcode = 1, qscale = 2;
    uint32_t prev_flags = 0;
    int cur_mbx = 0;                   // column of the macroblock being coded (tag of its coefficient records)

    explicit mp2v_gen(const mp2v_gen_params_t& params) : p(params), rng(params.seed) {}

    // ------------------------------------------------------------------ headers
    void sequence_header() {
        bw.start_code(0xB3);
        bw.put(p.width & 0xfff, 12); bw.put(p.height & 0xfff, 12);
        bw.put(1, 4); bw.put(5, 4);                 // aspect 1:1, 30 fps
        bw.put(0x3ffff, 18); bw.put(1, 1);          // bit_rate_value, marker
        bw.put(112, 10); bw.put(0, 1);              // vbv_buffer_size_value, constrained_parameters_flag
        bw.put(0, 1); bw.put(0, 1);                 // no matrices here (the reference ignores them anyway)
        bw.start_code(0xB5);                        // sequence_extension
        bw.put(1, 4); bw.put(p.chroma_format == 1 ? 0x44 : 0x82, 8);
        bw.put(1, 1); bw.put(p.chroma_format, 2);   // progressive_sequence = 1
        bw.put((p.width >> 12) & 3, 2); bw.put((p.height >> 12) & 3, 2);
        bw.put(0, 12); bw.put(1, 1); bw.put(0, 8); bw.put(0, 1); bw.put(0, 2); bw.put(0, 5);
        if (p.user_data_bytes > 0) {                // user_data(): bytes that cannot emulate a start code
            bw.start_code(0xB2);
            for (int i = 0; i < p.user_data_bytes; i++) bw.put(0x80u | (uint32_t)rng.below(0x7f), 8);
        }
    }
    void gop_header(int closed) {
        bw.start_code(0xB8);
        bw.put(0, 25); bw.put(closed, 1); bw.put(0, 1);
    }

    // ------------------------------------------------------------------ one picture
    struct pic_hdr_t { int type, temporal_reference, f_code[2][2], alt_scan, q_scale_type, dc_prec; };

    uint8_t gop_tx[4][64] = {};       // matrices_once: the matrices loaded by the GOP's first picture
    bool gop_tx_valid = false;

    void picture_headers(const pic_hdr_t& h, picture_t& pic) {
        bw.start_code(0x00);
        bw.put(h.temporal_reference & 0x3ff, 10); bw.put(h.type, 3); bw.put(0xffff, 16);
        if (h.type == 2 || h.type == 3) { bw.put(0, 1); bw.put(7, 3); }
        if (h.type == 3) { bw.put(0, 1); bw.put(7, 3); }
        bw.put(0, 1);
        bw.start_code(0xB5);                        // picture_coding_extension
        bw.put(8, 4);
        for (int s = 0; s < 2; s++) for (int t = 0; t < 2; t++) bw.put(h.f_code[s][t], 4);
        bw.put(h.dc_prec, 2); bw.put(3, 2);         // frame picture
        const int fpfd = p.pct_field_dct > 0 ? 0 : 1;      // frame_pred_frame_dct = 0 only in interlaced frames (progressive_frame = 0)
        bw.put(fpfd ? 0 : 1, 1); bw.put(fpfd, 1); bw.put(0, 1);   // top_field_first, frame_pred_frame_dct, concealment_mv
        bw.put(h.q_scale_type, 1); bw.put(p.intra_vlc_table0 ? 0 : 1, 1); bw.put(h.alt_scan, 1);   // q_scale_type, intra_vlc_format, alternate_scan
        bw.put(0, 1); bw.put(p.chroma_format == 1 && fpfd ? 1 : 0, 1); bw.put(fpfd, 1); bw.put(0, 1);   // repeat_first_field, chroma_420_type, progressive_frame, composite
        if (p.matrices_once && gop_tx_valid) {      // no extension: the matrices of the GOP's first picture stay in force
            for (int k = 0; k < 4; k++) {
                memcpy(pic.tx[k], gop_tx[k], 64);
                pic.tx_loaded[k] = 0;
                build_scan_indexed_matrix(pic.tx[k], h.alt_scan, pic.params.W[k]);
            }
            return;
        }
        bw.start_code(0xB5);                        // quant_matrix_extension, in EVERY picture
        bw.put(3, 4);
        const int nload = p.chroma_format == 1 ? 2 : 4;
        for (int k = 0; k < 4; k++) {
            const int load = k < nload;
            pic.tx_loaded[k] = load;
            bw.put(load, 1);
            if (!load) continue;
            for (int i = 0; i < 64; i++) {
                int v;
                if (p.mode >= 1) v = (k & 1) ? 16 : kDefaultIntra[scan_tables().shuffle[0][i]];
                else v = rng.range(8, 80);
                pic.tx[k][i] = (uint8_t)v;
                bw.put(v, 8);
            }
            build_scan_indexed_matrix(pic.tx[k], h.alt_scan, pic.params.W[k]);
        }
        if (p.matrices_once) {
            // 4:2:0 loads two matrices: the chroma pair follows the luminance pair (6.3.11)
            for (int k = nload; k < 4; k++) { memcpy(pic.tx[k], pic.tx[k - 2], 64); build_scan_indexed_matrix(pic.tx[k], h.alt_scan, pic.params.W[k]); }
            memcpy(gop_tx, pic.tx, sizeof(gop_tx));
            gop_tx_valid = true;
        }
    }

    // ------------------------------------------------------------------ coefficients
    void put_run_level(bool table_one, int run, int level) {   // level != 0
        const int mag = level < 0 ? -level : level;
        const char* code = (run < 64 && mag <= 40) ? (table_one ? enc().b15[run][mag] : enc().b14[run][mag]) : nullptr;
        if (code) { bw.put(code); bw.put(level < 0, 1); }
        else { bw.put(kCoefEscape); bw.put(run, 6); bw.put((uint32_t)level & 0xfff, 12); }
    }

    int draw_level() {
        int mag;
        if (p.mode == 1) { mag = 1; while (mag < 30 && rng.pct(35)) mag++; }
        else {
            const int r = rng.below(100);
            if (r < p.pct_big_levels) mag = rng.range(1, 2047);
            else if (r < p.pct_big_levels + 22) mag = rng.range(1, 40);
            else mag = rng.range(1, 3);
        }
        return rng.pct(50) ? -mag : mag;
    }
    int draw_count(bool at_least_one) {
        int n;
        if (p.mode == 1) {
            const int m = p.natural_mean_coefs;                  // geometric with mean m (default: continue 72 % -> 2.6)
            const int cont = m > 0 ? 100 * m / (m + 1) : 72;
            n = 0;
            while (n < (m > 0 ? 40 : 20) && rng.pct(cont)) n++;
        }
        else { static const int k[9] = {0, 0, 1, 1, 2, 3, 5, 8, 12}; n = k[rng.below(9)]; }
        return (at_least_one && n == 0) ? 1 : n;
    }
    int draw_run() {
        if (p.mode == 1) { int r = 0; while (r < 6 && rng.pct(30)) r++; return r; }
        static const int k[9] = {0, 0, 0, 1, 1, 2, 3, 5, 9};
        return k[rng.below(9)];
    }

    void code_block(picture_t& pic, int b, bool intra, int dc_prec) {
        int i = 0;
        if (intra) {
            const int comp = b < 4 ? 0 : 1 + (b & 1);
            const int maxdc = (1 << (8 + dc_prec)) - 1;
            int dc;
            if (p.mode == 1) { dc = dc_pred[comp] + rng.range(-12, 12) * (1 << dc_prec); dc = dc < 0 ? 0 : dc > maxdc ? maxdc : dc; }
            else dc = rng.range(0, maxdc);
            const int diff = dc - dc_pred[comp];
            dc_pred[comp] = dc;
            int size = 0;
            for (int a = diff < 0 ? -diff : diff; a; a >>= 1) size++;
            bw.put(enc().dcsize[comp ? 1 : 0][size]);
            if (size) bw.put((uint32_t)(diff > 0 ? diff : diff + (1 << size) - 1), size);
            pic.coef.push_back(MP2V_COEF((int16_t)(uint16_t)((uint32_t)dc << (3 - dc_prec)), 0, b, MP2V_COEF_RAW) | MP2V_COEF_MB(cur_mbx));
            i = 1;
        }
        const int n = draw_count(!intra);
        for (int k = 0; k < n; k++) {
            const int run = draw_run();
            if (i + run > 63) break;
            const int level = draw_level();
            i += run;
            if (!intra && i == 0 && k == 0 && (level == 1 || level == -1)) {   // B.14 note 3: "1s"
                bw.put("1"); bw.put(level < 0, 1);
                pic.coef.push_back(MP2V_COEF(level, 0, b, MP2V_COEF_FIRST) | MP2V_COEF_MB(cur_mbx));
            } else {
                put_run_level(intra && !p.intra_vlc_table0, run, level);
                pic.coef.push_back(MP2V_COEF(level, i, b, 0) | MP2V_COEF_MB(cur_mbx));
            }
            i++;
        }
        if (!intra && i == 0) {   // every run overshot: a coded non-intra block still needs one coefficient
            bw.put("1"); bw.put(0, 1);
            pic.coef.push_back(MP2V_COEF(1, 0, b, MP2V_COEF_FIRST) | MP2V_COEF_MB(cur_mbx));
        }
        bw.put((intra && !p.intra_vlc_table0) ? kEobB15 : kEobB14);
    }

    // ------------------------------------------------------------------ mode 2: texture content
    // A translating procedural texture plus per-frame noise is ENCODED: forward DCT of the source (intra) or of
    // the motion-compensated difference against the SOURCE reference frames (open loop: the decoder drifts by the
    // quantisation error, which is irrelevant here), quantised at quantiser_scale 4..8 with the default matrices.
    // Everything is integer arithmetic, so a seed gives the same stream on every machine.
    struct frame_t { int w[3], h[3]; std::vector<uint8_t> px[3]; };
    struct wave_t { int fx, fy, phase, amp; };
    wave_t waves[3][4];
    int gvx = 0, gvy = 0;                      // global motion, luma half-pels per frame
    frame_t ref_src[2];                        // source frames of the two live references (older, newer)
    int ref_time[2] = {0, 0};
    frame_t cur_src;

    static int psin(int p) {                   // parabolic sine, period 1024, amplitude 1024
        const int q = p & 511;
        const int y = (q * (512 - q)) >> 6;
        return (p & 512) ? -y : y;
    }
    void texture_setup() {
        gvx = rng.range(-5, 5); gvy = rng.range(-3, 3);
        if (gvx == 0 && gvy == 0) gvx = 3;
        static const int period[4] = {220, 46, 11, 6};      // pixels
        static const int amp[3][4] = {{38, 20, 9, 5}, {22, 10, 4, 0}, {22, 10, 4, 0}};
        for (int pl = 0; pl < 3; pl++)
            for (int i = 0; i < 4; i++) {
                // phase advance per HALF-pel step so that one period spans period[i] pixels, split over x and y
                const int f = 1024 / (2 * period[i]) + 1;
                const int a = rng.range(0, f);
                waves[pl][i] = {rng.pct(50) ? a : -a, rng.pct(50) ? f - a : a - f, rng.below(1024), amp[pl][i]};
            }
    }
    static uint32_t hash32(uint64_t x) {
        x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
        return (uint32_t)x;
    }
    // source picture at display time t: plane sample (x, y) looks at texture position (sx * x + t * gv) in luma half-pels
    void synth_frame(int t, frame_t& f) {
        const int noise = p.texture_noise > 0 ? p.texture_noise : 3;
        for (int pl = 0; pl < 3; pl++) {
            const int w = pl == 0 ? p.width : (p.chroma_format == 3 ? p.width : p.width / 2);
            const int h = pl == 0 ? p.height : (p.chroma_format == 1 ? p.height / 2 : p.height);
            const int sx = 2 * p.width / w, sy = 2 * p.height / h;      // luma half-pels per sample of this plane
            f.w[pl] = w; f.h[pl] = h;
            f.px[pl].resize((size_t)w * h);
            uint8_t* const dstpx = f.px[pl].data();
            parallel_rows(h, [=](int y) {
                const int Y = sy * y + t * gvy;
                for (int x = 0; x < w; x++) {
                    const int X = sx * x + t * gvx;
                    int v = 0;
                    for (int i = 0; i < 4; i++) { const wave_t& q = waves[pl][i]; v += q.amp * psin(q.fx * X + q.fy * Y + q.phase); }
                    const uint32_t hsh = hash32(((uint64_t)p.seed << 40) ^ ((uint64_t)(t & 0xfff) << 28) ^ ((uint64_t)pl << 26) ^ ((uint64_t)y << 13) ^ (uint64_t)x);
                    v = 128 + (v >> 10) + (int)(hsh % (uint32_t)(2 * noise + 1)) - noise;
                    dstpx[(size_t)y * w + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
                }
            });
        }
    }
    // forward 8x8 DCT (ISO/IEC 13818-2 Annex A normalisation), 14-bit fixed-point basis, out[v * 8 + u]
    static void fdct8x8(const int in[64], int out[64]) {
        static const int c16[9] = {16384, 16069, 15137, 13623, 11585, 9102, 6270, 3196, 0};      // 2^14 cos(k pi / 16)
        auto cosi = [&](int i) { i &= 31; if (i > 16) i = 32 - i; return i <= 8 ? c16[i] : -c16[16 - i]; };
        long long tmp[64];
        for (int y = 0; y < 8; y++)
            for (int u = 0; u < 8; u++) {
                long long acc = 0;
                for (int x = 0; x < 8; x++) acc += (long long)in[y * 8 + x] * (u ? cosi((2 * x + 1) * u) : 11585);
                tmp[y * 8 + u] = acc;                                   // scale 2^14 (x C(u) with C(0) = 1/sqrt 2)
            }
        for (int u = 0; u < 8; u++)
            for (int v = 0; v < 8; v++) {
                long long acc = 0;
                for (int y = 0; y < 8; y++) acc += tmp[y * 8 + u] * (v ? cosi((2 * y + 1) * v) : 11585);
                out[v * 8 + u] = (int)((acc + (1ll << 29)) >> 30);      // / 2^28 for the two bases, / 4 for the 2/N factors
            }
    }
    // half-pel prediction of a w x h block from a source reference plane, the reference decoder's arithmetic (mc_c.hpp:3-17)
    static void predict_block(const frame_t& ref, int pl, int x0, int y0, int hx, int hy, int w, int h, int* out /* w*h */) {
        const int W = ref.w[pl];
        const uint8_t* px = ref.px[pl].data();
        auto avg = [](int a, int b) { return (a + b + 1) >> 1; };
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                const uint8_t* q = px + (size_t)(y0 + y) * W + x0 + x;
                int v = q[0];
                if (hx && !hy) v = avg(q[0], q[1]);
                else if (!hx && hy) v = avg(q[0], q[W]);
                else if (hx && hy) v = avg(avg(q[0], q[1]), avg(q[W], q[W + 1]));
                out[y * w + x] = v;
            }
    }
    // residual (or intra samples) of block b of macroblock (mbx, mby) -> quantised levels in SCAN order; returns true when any is non-zero.
    // dc_out: intra only, the quantised DC (dct_dc as transmitted at this precision)
    // field_dct: the block holds one field of its half of the macroblock (rows two apart; the lower block starts on row 1)
    bool quantise_block(const picture_t& pic, int b, int mbx, int mby, bool intra, bool field_dct, const int* pred_y, const int* pred_c[2], int alt, int dc_prec,
                        int qs, int levels[64], int* dc_out) {
        const int cf = p.chroma_format;
        int pl, bx, top, lower;                                            // plane, pixel column, macroblock's first row, upper / lower block (mb_decoder.cpp:177-195)
        const int cw = cf == 3 ? 16 : 8, chh = cf == 1 ? 8 : 16;
        if (b < 4) { pl = 0; bx = mbx * 16 + 8 * (b & 1); top = mby * 16; lower = b >> 1; }
        else {
            pl = 1 + (b & 1);
            const int k = (b - 4) >> 1;                                    // 0; 4:2:2 1 = lower; 4:4:4 2,3 = right column
            bx = mbx * cw + (cf == 3 ? 8 * (k >> 1) : 0);
            top = mby * chh; lower = k & 1;
        }
        const bool fields = field_dct && (pl == 0 || cf != 1);             // 4:2:0 chroma is always frame organised (:172)
        const int W = cur_src.w[pl];
        int blk[64], F[64];
        for (int y = 0; y < 8; y++) {
            const int row = fields ? lower + 2 * y : 8 * lower + y;        // inside the macroblock
            for (int x = 0; x < 8; x++) {
                int v = cur_src.px[pl][(size_t)(top + row) * W + bx + x];
                if (!intra) {
                    if (pl == 0) v -= pred_y[row * 16 + (bx - mbx * 16 + x)];
                    else v -= pred_c[pl - 1][row * cw + (bx - mbx * cw + x)];
                }
                blk[y * 8 + x] = v;
            }
        }
        fdct8x8(blk, F);
        const scan_tables_t& st = scan_tables();
        const uint8_t* Wq = pic.params.W[(b < 6 ? 0 : 2) + (intra ? 0 : 1)];   // the reference's matrix choice (blocks 4, 5 use the luminance pair)
        bool any = false;
        for (int i = 0; i < 64; i++) {
            const int f = F[st.shuffle[alt][i]];
            const int a = f < 0 ? -f : f;
            int q;
            if (intra && i == 0) {
                const int mult = 8 >> dc_prec;
                int dc = (F[0] + mult / 2) / mult;
                const int maxdc = (1 << (8 + dc_prec)) - 1;
                *dc_out = dc < 0 ? 0 : dc > maxdc ? maxdc : dc;
                levels[0] = 0;
                continue;
            }
            const int step16 = Wq[i] * qs;                                 // 16 x the reconstruction step
            if (intra) q = (16 * a + (step16 * 3) / 8) / step16;           // (level * W * qs) >> 4 with a small dead zone
            else q = (16 * a) / step16;                                    // ((2 level + 1) * W * qs) >> 5: centre of the cell
            if (q > 2047) q = 2047;
            levels[i] = f < 0 ? -q : q;
            any = any || q != 0;
        }
        return any || intra;
    }
    // VLC-code one block from its quantised levels (scan order) and append the ground-truth records
    void emit_block(picture_t& pic, int b, bool intra, int dc_prec, int dc, const int levels[64]) {
        int i = 0;
        if (intra) {
            const int comp = b < 4 ? 0 : 1 + (b & 1);
            const int diff = dc - dc_pred[comp];
            dc_pred[comp] = dc;
            int size = 0;
            for (int a = diff < 0 ? -diff : diff; a; a >>= 1) size++;
            bw.put(enc().dcsize[comp ? 1 : 0][size]);
            if (size) bw.put((uint32_t)(diff > 0 ? diff : diff + (1 << size) - 1), size);
            pic.coef.push_back(MP2V_COEF((int16_t)(uint16_t)((uint32_t)dc << (3 - dc_prec)), 0, b, MP2V_COEF_RAW) | MP2V_COEF_MB(cur_mbx));
            i = 1;
        }
        int run = 0;
        bool first = !intra;
        for (; i < 64; i++) {
            const int level = levels[i];
            if (!level) { run++; continue; }
            if (first && i == 0 && (level == 1 || level == -1)) {          // B.14 note 3: "1s"
                bw.put("1"); bw.put(level < 0, 1);
                pic.coef.push_back(MP2V_COEF(level, 0, b, MP2V_COEF_FIRST) | MP2V_COEF_MB(cur_mbx));
            } else {
                put_run_level(intra && !p.intra_vlc_table0, run, level);
                pic.coef.push_back(MP2V_COEF(level, i, b, 0) | MP2V_COEF_MB(cur_mbx));
            }
            first = false;
            run = 0;
        }
        bw.put((intra && !p.intra_vlc_table0) ? kEobB15 : kEobB14);
    }

    // rows of a picture are analysed on several threads (results do not depend on the thread count)
    template <class F>
    static void parallel_rows(int n, F fn) {
        unsigned hw = std::thread::hardware_concurrency();
        const int nt = (int)std::min<unsigned>(hw ? hw : 4u, 32u);
        if (nt <= 1 || n < 4) { for (int i = 0; i < n; i++) fn(i); return; }
        std::vector<std::thread> th;
        for (int k = 0; k < nt; k++) th.emplace_back([=] { for (int i = k; i < n; i += nt) fn(i); });
        for (auto& x : th) x.join();
    }

    struct mb_plan_t { bool intra, fwd, bwd, field_dct; int mv[2][2]; uint32_t cbp; int dcq[12]; int16_t lv[12][64]; };
    std::vector<mb_plan_t> plan;

    void code_picture_texture(const pic_hdr_t& h, picture_t& pic, int t) {
        const bool big = p.height > 2800;
        const int cf = p.chroma_format, cw = cf == 3 ? 16 : 8, ch = cf == 1 ? 8 : 16;
        // motion of this picture against its references: content moves gv per frame, so the matching block of a
        // reference shown at time tr lies (t - tr) * gv half-pels away
        int gmv[2][2] = {{0, 0}, {0, 0}};
        if (h.type == 2) { gmv[0][0] = (t - ref_time[1]) * gvx; gmv[0][1] = (t - ref_time[1]) * gvy; }
        if (h.type == 3) {
            gmv[0][0] = (t - ref_time[0]) * gvx; gmv[0][1] = (t - ref_time[0]) * gvy;
            gmv[1][0] = (t - ref_time[1]) * gvx; gmv[1][1] = (t - ref_time[1]) * gvy;
        }
        // ---- quantiser per slice: quantiser_scale 4 .. 8 with either mapping (decoder.cpp:140-145)
        std::vector<int> row_qcode((size_t)mbh);
        for (int mby = 0; mby < mbh; mby++) row_qcode[mby] = h.q_scale_type ? rng.range(4, 8) : rng.range(2, 4);
        // ---- analysis (parallel): prediction mode, prediction, transform, quantisation of every macroblock
        plan.resize((size_t)mbw * mbh);
        const size_t pic_no = pics.size();
        parallel_rows(mbh, [&, pic_no](int mby) {
            std::vector<int> py(256), pcb(16 * 16), pcr(16 * 16), tmp(256);
            const int qs = quantiser_scale_of(row_qcode[mby], h.q_scale_type);
            for (int mbx = 0; mbx < mbw; mbx++) {
                mb_plan_t& m = plan[(size_t)mby * mbw + mbx];
                // the global vector(s) where the window stays inside the frame, intra otherwise
                bool fwd = false, bwd = false;
                if (h.type == 2) fwd = window_ok(mbx, mby, gmv[0][0], gmv[0][1]);
                if (h.type == 3) {
                    const bool okf = window_ok(mbx, mby, gmv[0][0], gmv[0][1]), okb = window_ok(mbx, mby, gmv[1][0], gmv[1][1]);
                    const int d = (int)(hash32(((uint64_t)p.seed << 32) ^ ((uint64_t)pic_no << 20) ^ (uint64_t)(mby * mbw + mbx)) % 10u);
                    fwd = okf && d < 7;                                    // 30 % forward, 40 % both, 30 % backward
                    bwd = okb && d >= 3;
                    if (!fwd && !bwd) { fwd = okf; bwd = !okf && okb; }    // whichever window stays inside the frame
                }
                const bool intra = h.type == 1 || (!fwd && !bwd) ||
                                   (int)(hash32(((uint64_t)p.seed << 33) ^ ((uint64_t)pic_no << 21) ^ (uint64_t)(mby * mbw + mbx) ^ 0x5bd1e995u) % 100u) < p.pct_intra_in_pb;
                if (intra) fwd = bwd = false;
                m.intra = intra; m.fwd = fwd; m.bwd = bwd;
                m.field_dct = p.pct_field_dct > 0 &&
                              (int)(hash32(((uint64_t)p.seed << 31) ^ ((uint64_t)pic_no << 22) ^ (uint64_t)(mby * mbw + mbx) ^ 0x9e3779b9u) % 100u) < p.pct_field_dct;
                memset(m.mv, 0, sizeof(m.mv));
                const int* pc[2] = {pcb.data(), pcr.data()};
                if (!intra) {
                    bool have = false;
                    for (int s2 = 0; s2 < 2; s2++) {
                        if (!(s2 ? bwd : fwd)) continue;
                        m.mv[s2][0] = gmv[s2][0]; m.mv[s2][1] = gmv[s2][1];
                        const frame_t& ref = h.type == 2 ? ref_src[1] : ref_src[s2];
                        const int cx = cf < 3 ? m.mv[s2][0] >> 1 : m.mv[s2][0], cy = cf < 2 ? m.mv[s2][1] >> 1 : m.mv[s2][1];   // floor (mb_decoder.cpp:198-206)
                        for (int pl = 0; pl < 3; pl++) {
                            const int w = pl ? cw : 16, hh = pl ? ch : 16, vx = pl ? cx : m.mv[s2][0], vy = pl ? cy : m.mv[s2][1];
                            int* dst = pl == 0 ? py.data() : pl == 1 ? pcb.data() : pcr.data();
                            predict_block(ref, pl, mbx * w + (vx >> 1), mby * hh + (vy >> 1), vx & 1, vy & 1, w, hh, have ? tmp.data() : dst);
                            if (have) for (int k = 0; k < w * hh; k++) dst[k] = (tmp[k] + dst[k] + 1) >> 1;    // avg(backward, forward), mb_decoder.cpp:240-249
                        }
                        have = true;
                    }
                }
                m.cbp = 0;
                int lv[64];
                for (int b = 0; b < nblk; b++) {
                    if (quantise_block(pic, b, mbx, mby, intra, m.field_dct, py.data(), pc, h.alt_scan, h.dc_prec, qs, lv, &m.dcq[b])) m.cbp |= 1u << b;
                    for (int i = 0; i < 64; i++) m.lv[b][i] = (int16_t)lv[i];
                }
            }
        });
        // ---- syntax (serial): skipped macroblocks, VLC codes, ground-truth records
        int lvi[64];
        for (int mby = 0; mby < mbh; mby++) {
            bw.start_code(big ? (mby & 127) + 1 : mby + 1);
            if (big) bw.put(mby >> 7, 3);
            qcode = row_qcode[mby];
            qscale = quantiser_scale_of(qcode, h.q_scale_type);
            bw.put(qcode, 5); bw.put(0, 1);
            memset(pmv, 0, sizeof(pmv));
            for (int c = 0; c < 3; c++) dc_pred[c] = 1 << (h.dc_prec + 7);
            prev_flags = 0;
            int pending_skips = 0;
            for (int mbx = 0; mbx < mbw; mbx++) {
                const mb_plan_t& m = plan[(size_t)mby * mbw + mbx];
                cur_mbx = mbx;
                mp2v_mb_info_t rec{};
                rec.coef_off = (uint32_t)pic.coef.size();
                const bool edge = mbx == 0 || mbx == mbw - 1;
                const bool intra = m.intra;
                bool fwd = m.fwd, bwd = m.bwd;
                uint32_t cbp = m.cbp;
                const int (*mv)[2] = m.mv;
                // ---- skipped?  P: zero vector and nothing coded; B: same prediction as the previous macroblock and nothing coded
                if (!intra && cbp == 0 && !edge) {
                    bool skip = false;
                    if (h.type == 2) skip = mv[0][0] == 0 && mv[0][1] == 0;
                    else {
                        const uint32_t fl = (fwd ? MP2V_MB_FWD : 0u) | (bwd ? MP2V_MB_BWD : 0u);
                        skip = (prev_flags & (MP2V_MB_FWD | MP2V_MB_BWD)) == fl && !(prev_flags & MP2V_MB_INTRA) && prev_flags != 0;
                        for (int s2 = 0; s2 < 2 && skip; s2++) if ((s2 ? bwd : fwd) && (pmv[s2][0] != mv[s2][0] || pmv[s2][1] != mv[s2][1])) skip = false;
                    }
                    if (skip) {
                        const uint32_t fl = h.type == 2 ? MP2V_MB_FWD : (prev_flags & (MP2V_MB_FWD | MP2V_MB_BWD));
                        if (h.type == 2) memset(pmv, 0, sizeof(pmv));      // mb_decoder.cpp:542-543
                        rec.bits = MP2V_MB_BITS(0, qscale, 0, fl);
                        for (int s2 = 0; s2 < 2; s2++) for (int k = 0; k < 2; k++) rec.mv[s2][k] = (int16_t)((fl & (s2 ? MP2V_MB_BWD : MP2V_MB_FWD)) ? mv[s2][k] : 0);
                        pic.mb.push_back(rec);
                        pending_skips++;
                        continue;
                    }
                }
                // ---- coded macroblock
                int inc = pending_skips + 1;
                const bool had_skips = pending_skips > 0;
                pending_skips = 0;
                while (inc > 33) { bw.put(enc().mba_escape); inc -= 33; }
                bw.put(enc().mba[inc]);
                const bool pattern = !intra && cbp != 0;
                if (h.type == 2 && !intra && !pattern) fwd = true;         // "MC, not coded"
                uint32_t type = intra ? 0x02 : (fwd ? 0x10 : 0) | (bwd ? 0x08 : 0) | (pattern ? 0x04 : 0);
                bw.put(enc().mbtype[h.type][type]);
                if (p.pct_field_dct > 0) {                                 // macroblock_modes with frame_pred_frame_dct = 0
                    if (fwd || bwd) bw.put(2, 2);                          // frame_motion_type: frame-based
                    if (intra || pattern) { bw.put(m.field_dct ? 1 : 0, 1); if (m.field_dct) rec.coef_off |= MP2V_MB_FIELD_DCT; }
                }
                for (int s2 = 0; s2 < 2; s2++) {
                    if (!(s2 ? bwd : fwd)) continue;
                    put_mv_component(mv[s2][0], pmv[s2][0], h.f_code[s2][0]);
                    put_mv_component(mv[s2][1], pmv[s2][1], h.f_code[s2][1]);
                }
                if (intra) memset(pmv, 0, sizeof(pmv));                    // mb_decoder.cpp:599-603
                if (had_skips || !intra) for (int c = 0; c < 3; c++) dc_pred[c] = 1 << (h.dc_prec + 7);   // mb_decoder.cpp:623-626
                if (intra) cbp = (1u << nblk) - 1;
                else if (pattern) {
                    uint32_t c420 = 0;
                    for (int i = 0; i < 6; i++) if (cbp & (1u << i)) c420 |= 1u << (5 - i);
                    bw.put(enc().cbp[c420]);
                    if (cf == 2) bw.put((cbp >> 6 & 1) << 1 | (cbp >> 7 & 1), 2);
                    if (cf == 3) for (int i = 6; i < 12; i++) bw.put((cbp >> i) & 1, 1);
                }
                for (int b = 0; b < nblk; b++) if (cbp & (1u << b)) {
                    for (int i = 0; i < 64; i++) lvi[i] = m.lv[b][i];
                    emit_block(pic, b, intra, h.dc_prec, m.dcq[b], lvi);
                }
                uint32_t fl = intra ? MP2V_MB_INTRA : (fwd ? MP2V_MB_FWD : 0u) | (bwd ? MP2V_MB_BWD : 0u);
                if (!intra) for (int s2 = 0; s2 < 2; s2++) for (int k = 0; k < 2; k++) rec.mv[s2][k] = (int16_t)mv[s2][k];
                rec.bits = MP2V_MB_BITS(pic.coef.size() - MP2V_MB_COEF_OFF(rec.coef_off), qscale, cbp, fl);
                pic.mb.push_back(rec);
                prev_flags = fl;
            }
        }
        pic.params.n_coef = (uint32_t)pic.coef.size();
        if (getenv("MP2V_GEN_DEBUG"))
            fprintf(stderr, "[gen] picture %zu type %d t %d gv (%d,%d) fwd mv (%d,%d) bwd mv (%d,%d) ref times %d %d: %u coefficients\n", pics.size() - 1, h.type, t, gvx, gvy,
                    gmv[0][0], gmv[0][1], gmv[1][0], gmv[1][1], ref_time[0], ref_time[1], pic.params.n_coef);
    }

    // ------------------------------------------------------------------ motion vectors
    bool window_ok(int mbx, int mby, int mvx, int mvy) const {
        const int x0 = mbx * 16 + (mvx >> 1), y0 = mby * 16 + (mvy >> 1);
        return x0 >= 0 && y0 >= 0 && x0 + 16 + (mvx & 1) <= p.width && y0 + 16 + (mvy & 1) <= p.height;
    }
    int draw_mv_component(int pos, int extent, int f_code) {
        const int f = 1 << (f_code - 1);
        int half = rng.pct(50);
        int lo = -p.mv_range, hi = p.mv_range;
        if (p.unclamped_mv) return rng.range(lo, hi) * 2 + half;
        if (lo < -pos) lo = -pos;
        if (hi > extent - 16 - pos - half) hi = extent - 16 - pos - half;
        if (hi < lo) { half = 0; hi = extent - 16 - pos; if (hi > p.mv_range) hi = p.mv_range; if (hi < lo) hi = lo; }
        (void)f;   // generate(): f_code is chosen so that +-(2*mv_range+1) half-pels always fit [-16f, 16f-1]
        return rng.range(lo, hi) * 2 + half;
    }
    void put_mv_component(int mv, int& pred, int f_code) {   // mb_decoder.cpp:447-503 inverted
        const int r_size = f_code - 1, f = 1 << r_size;
        int delta = mv - pred;
        if (delta < -16 * f) delta += 32 * f;
        if (delta > 16 * f - 1) delta -= 32 * f;
        if (delta == 0) bw.put(enc().motion[0]);
        else {
            const int m = (delta < 0 ? -delta : delta) - 1;
            bw.put(enc().motion[m / f + 1]); bw.put(delta < 0, 1);
            if (r_size) bw.put((uint32_t)(m % f), r_size);
        }
        pred = mv;
    }

    // ------------------------------------------------------------------ slices / macroblocks
    void code_picture(const pic_hdr_t& h, picture_t& pic) {
        const bool big = p.height > 2800;
        for (int mby = 0; mby < mbh; mby++) {
            bw.start_code(big ? (mby & 127) + 1 : mby + 1);
            if (big) bw.put(mby >> 7, 3);
            qcode = rng.range(1, p.qscale_code_max);
            qscale = quantiser_scale_of(qcode, h.q_scale_type);
            bw.put(qcode, 5); bw.put(0, 1);
            memset(pmv, 0, sizeof(pmv));
            for (int c = 0; c < 3; c++) dc_pred[c] = 1 << (h.dc_prec + 7);
            prev_flags = 0;
            int pending_skips = 0;
            for (int mbx = 0; mbx < mbw; mbx++) {
                cur_mbx = mbx;
                mp2v_mb_info_t rec{};
                rec.coef_off = (uint32_t)pic.coef.size();
                const bool edge = mbx == 0 || mbx == mbw - 1;
                // ---- skipped?
                if (h.type != 1 && !edge && rng.pct(p.pct_skipped)) {
                    bool ok = true;
                    uint32_t fl = MP2V_MB_FWD;
                    int mv[2][2] = {{0, 0}, {0, 0}};
                    if (h.type == 3) {
                        fl = prev_flags & (MP2V_MB_FWD | MP2V_MB_BWD);
                        ok = fl != 0;                                   // not after an intra macroblock
                        memcpy(mv, pmv, sizeof(mv));
                        if (ok && (fl & MP2V_MB_FWD)) ok = window_ok(mbx, mby, mv[0][0], mv[0][1]);
                        if (ok && (fl & MP2V_MB_BWD)) ok = window_ok(mbx, mby, mv[1][0], mv[1][1]);
                    }
                    if (ok) {
                        if (h.type == 2) memset(pmv, 0, sizeof(pmv));   // mb_decoder.cpp:542-543
                        rec.bits = MP2V_MB_BITS(0, qscale, 0, fl);
                        for (int s = 0; s < 2; s++) for (int t = 0; t < 2; t++)
                            rec.mv[s][t] = (int16_t)(((fl & (s ? MP2V_MB_BWD : MP2V_MB_FWD)) ? mv[s][t] : 0));
                        pic.mb.push_back(rec);
                        pending_skips++;
                        continue;
                    }
                }
                // ---- coded macroblock: address increment
                int inc = pending_skips + 1;
                const bool had_skips = pending_skips > 0;
                pending_skips = 0;
                while (inc > 33) { bw.put(enc().mba_escape); inc -= 33; }
                bw.put(enc().mba[inc]);
                // ---- modes
                bool intra = h.type == 1 || rng.pct(p.pct_intra_in_pb);
                uint32_t type = 0;      // macroblock_type flag byte
                bool fwd = false, bwd = false, pattern = false;
                if (intra) type = 0x02;
                else if (h.type == 2) {
                    pattern = rng.pct(p.pct_coded);
                    fwd = pattern ? rng.pct(75) : true;
                    type = (fwd ? 0x10 : 0) | (pattern ? 0x04 : 0);
                } else {
                    const int d = rng.below(3);
                    fwd = d != 1; bwd = d != 0;
                    pattern = rng.pct(p.pct_coded);
                    type = (fwd ? 0x10 : 0) | (bwd ? 0x08 : 0) | (pattern ? 0x04 : 0);
                }
                const bool quant = (intra || pattern) && rng.pct(p.pct_mb_quant);
                if (quant) type |= 0x20;
                bw.put(enc().mbtype[h.type][type]);
                if (p.pct_field_dct > 0) {                              // macroblock_modes with frame_pred_frame_dct = 0
                    if (fwd || bwd) bw.put(2, 2);                       // frame_motion_type: frame-based
                    if (intra || pattern) { const bool fd = rng.pct(p.pct_field_dct); bw.put(fd ? 1 : 0, 1); if (fd) rec.coef_off |= MP2V_MB_FIELD_DCT; }
                }
                if (quant) {
                    qcode = rng.range(1, p.qscale_code_max);
                    qscale = quantiser_scale_of(qcode, h.q_scale_type);
                    bw.put(qcode, 5);
                }
                // ---- motion vectors
                int mv[2][2] = {{0, 0}, {0, 0}};
                for (int s = 0; s < 2; s++) {
                    if (!(s ? bwd : fwd)) continue;
                    mv[s][0] = draw_mv_component(mbx * 16, p.width, h.f_code[s][0]);
                    mv[s][1] = draw_mv_component(mby * 16, p.height, h.f_code[s][1]);
                    put_mv_component(mv[s][0], pmv[s][0], h.f_code[s][0]);
                    put_mv_component(mv[s][1], pmv[s][1], h.f_code[s][1]);
                }
                if (intra || (h.type == 2 && !fwd)) memset(pmv, 0, sizeof(pmv));   // mb_decoder.cpp:599-603
                // ---- DC predictor reset (mb_decoder.cpp:623-626)
                if (had_skips || !intra) for (int c = 0; c < 3; c++) dc_pred[c] = 1 << (h.dc_prec + 7);
                // ---- coded block pattern
                uint32_t cbp = 0;
                if (intra) cbp = (1u << nblk) - 1;
                else if (pattern) {
                    if (p.all_blocks_coded) cbp = (1u << nblk) - 1;
                    else do {
                        cbp = (uint32_t)rng.next() & ((1u << nblk) - 1);
                    } while (cbp == 0 || (p.chroma_format == 1 && (cbp & 63) == 0));
                    uint32_t c420 = 0;
                    for (int i = 0; i < 6; i++) if (cbp & (1u << i)) c420 |= 1u << (5 - i);
                    bw.put(enc().cbp[c420]);
                    if (p.chroma_format == 2) bw.put((cbp >> 6 & 1) << 1 | (cbp >> 7 & 1), 2);
                    if (p.chroma_format == 3) for (int i = 6; i < 12; i++) bw.put((cbp >> i) & 1, 1);
                }
                for (int b = 0; b < nblk; b++) if (cbp & (1u << b)) code_block(pic, b, intra, h.dc_prec);
                // ---- ground-truth record
                uint32_t fl = intra ? MP2V_MB_INTRA : 0;
                if (!intra) {
                    if (fwd) fl |= MP2V_MB_FWD;
                    if (bwd) fl |= MP2V_MB_BWD;
                    if (!fwd && !bwd) fl |= MP2V_MB_FWD;      // P "no MC": forward, zero vector (mb_decoder.cpp:329-338)
                    for (int s = 0; s < 2; s++) for (int t = 0; t < 2; t++) rec.mv[s][t] = (int16_t)mv[s][t];
                }
                rec.bits = MP2V_MB_BITS(pic.coef.size() - MP2V_MB_COEF_OFF(rec.coef_off), qscale, cbp, fl);
                pic.mb.push_back(rec);
                prev_flags = fl;
            }
        }
        pic.params.n_coef = (uint32_t)pic.coef.size();
    }

    bool generate() {
        if (p.width <= 0 || p.height <= 0 || (p.width & 15) || (p.height & 15) || p.chroma_format < 1 || p.chroma_format > 3 ||
            p.n_gops < 1 || p.gop_n < 1 || p.gop_m < 1 || p.qscale_code_max < 1 || p.qscale_code_max > 31 || p.mv_range < 0) {
            err = "bad generator parameters";
            return false;
        }
        mbw = p.width / 16; mbh = p.height / 16;
        nblk = p.chroma_format == 1 ? 6 : p.chroma_format == 2 ? 8 : 12;
        if (p.mode == 2) texture_setup();
        int display_base = 0;
        for (int g = 0; g < p.n_gops; g++) {
            bw.align();
            gop_off.push_back(bw.bytes.size());
            sequence_header();
            gop_header(1);
            gop_tx_valid = false;                   // a sequence header resets the matrices
            // display positions of the references of this closed GOP: 0, m, 2m, ... and the last picture
            std::vector<int> refs;
            if (p.intra_only || p.gop_m == 1) for (int d = 0; d < p.gop_n; d++) refs.push_back(d);
            else { for (int d = 0; d < p.gop_n; d += p.gop_m) refs.push_back(d); if (refs.back() != p.gop_n - 1) refs.push_back(p.gop_n - 1); }
            int prev_ref_coded = -1, prev_prev_ref_coded = -1;
            for (size_t r = 0; r < refs.size(); r++) {
                // the reference picture itself, then the B pictures displayed before it
                const int first_b = r ? refs[r - 1] + 1 : refs[r];
                for (int k = -1; k < refs[r] - first_b; k++) {
                    const int disp = k < 0 ? refs[r] : first_b + k;
                    const bool is_ref = k < 0;
                    if (!is_ref && r == 0) break;
                    pic_hdr_t h{};
                    h.type = is_ref ? ((r == 0 || p.intra_only) ? 1 : 2) : 3;
                    h.temporal_reference = disp;
                    h.alt_scan = p.alternate_scan < 0 ? rng.below(2) : p.alternate_scan;
                    h.q_scale_type = p.q_scale_type < 0 ? rng.below(2) : p.q_scale_type;
                    h.dc_prec = p.intra_dc_precision < 0 ? rng.below(4) : p.intra_dc_precision;
                    int need = 1;                           // smallest f_code covering +-(2*mv_range+1) half-pels
                    while ((16 << (need - 1)) <= 2 * p.mv_range + 1) need++;
                    if (p.mode == 2) {                      // ... or the global motion over the longest reference distance
                        const int far = (gvx < 0 ? -gvx : gvx) > (gvy < 0 ? -gvy : gvy) ? (gvx < 0 ? -gvx : gvx) : (gvy < 0 ? -gvy : gvy);
                        need = 1;
                        while ((16 << (need - 1)) <= far * (p.gop_m > p.gop_n ? p.gop_n : p.gop_m) + 1) need++;
                    }
                    for (int s = 0; s < 2; s++) for (int t = 0; t < 2; t++) {
                        const bool used = (h.type == 2 && s == 0) || h.type == 3;
                        int fc = need + rng.below(3);
                        h.f_code[s][t] = used ? (fc > 9 ? 9 : fc) : 15;
                    }
                    pics.emplace_back();
                    picture_t& pic = pics.back();
                    const int coded = (int)pics.size() - 1;
                    pic.gop = g; pic.display_index = display_base + disp;
                    pic.q_scale_type = h.q_scale_type; pic.intra_dc_precision = h.dc_prec;
                    pic.params.picture_coding_type = h.type;
                    pic.params.alternate_scan = h.alt_scan;
                    pic.params.dst_frame = coded;
                    pic.params.l0_frame = h.type == 2 ? prev_ref_coded : h.type == 3 ? prev_prev_ref_coded : -1;
                    pic.params.l1_frame = h.type == 3 ? prev_ref_coded : -1;
                    picture_headers(h, pic);
                    if (p.mode == 2) {
                        const int t = display_base + disp;
                        synth_frame(t, cur_src);
                        code_picture_texture(h, pic, t);
                        if (is_ref) { std::swap(ref_src[0], ref_src[1]); std::swap(ref_src[1], cur_src); ref_time[0] = ref_time[1]; ref_time[1] = t; }
                    } else {
                        code_picture(h, pic);
                    }
                    if (is_ref) { prev_prev_ref_coded = prev_ref_coded; prev_ref_coded = coded; }
                }
            }
            display_base += p.gop_n;
        }
        bw.align();
        gop_off.push_back(bw.bytes.size());
        bw.start_code(0xB7);                               // sequence_end_code
        // the reference's readers run past the end (start_codes_search.hpp:11-16, bitstream.h:28-34)
        bw.bytes.insert(bw.bytes.end(), 256, 0);
        return check_no_emulation();
    }

    // Start-code emulation (a byte-aligned 00 00 01 that is not a real start code) or 23 zero bits
    // inside a slice (the reference's end-of-slice test, decoder.cpp:150) would derail decoding.
    // The syntax written above cannot produce either; verify anyway so a generator bug is loud.
    bool check_no_emulation() {
        const std::vector<uint8_t>& b = bw.bytes;
        const size_t end = b.size() - 256;
        size_t k = 0;
        for (size_t i = 0; i + 2 < end; i++) {
            if (b[i] || b[i + 1] || b[i + 2] != 1) continue;
            while (k < bw.start_codes.size() && bw.start_codes[k] < i) k++;
            if (k >= bw.start_codes.size() || bw.start_codes[k] != i) { err = "start code emulation at byte " + std::to_string(i); return false; }
        }
        for (size_t c = 0; c + 1 < bw.start_codes.size(); c++) {
            const size_t at = bw.start_codes[c], next = bw.start_codes[c + 1];
            if (b[at + 3] < 1 || b[at + 3] > 0xaf) continue;                 // slices only
            size_t last_one = 0;                                              // bit index of the last 1 bit of the payload
            for (size_t bit = (at + 4) * 8; bit < next * 8; bit++) if ((b[bit >> 3] >> (7 - (bit & 7))) & 1) last_one = bit;
            int zeros = 0;
            for (size_t bit = (at + 4) * 8; bit <= last_one; bit++) {
                if ((b[bit >> 3] >> (7 - (bit & 7))) & 1) zeros = 0;
                else if (++zeros >= 23) { err = "23 zero bits inside slice at byte " + std::to_string(bit >> 3); return false; }
            }
        }
        return true;
    }
};

extern "C" {

MP2V_API void mp2v_gen_default_params(mp2v_gen_params_t* p, int width, int height, int chroma_format) {
    memset(p, 0, sizeof(*p));
    p->width = width; p->height = height; p->chroma_format = chroma_format;
    p->n_gops = 1; p->gop_n = 15; p->gop_m = 3; p->seed = 1;
    p->mode = 0; p->mv_range = 24; p->qscale_code_max = 12;
    p->alternate_scan = -1; p->q_scale_type = -1; p->intra_dc_precision = -1;
    p->pct_skipped = 15; p->pct_intra_in_pb = 8; p->pct_coded = 60; p->pct_mb_quant = 20; p->pct_big_levels = 3;
}

MP2V_API mp2v_gen_t* mp2v_gen_create(const mp2v_gen_params_t* p) {
    if (!p) return nullptr;
    mp2v_gen* g = new mp2v_gen(*p);
    if (!g->generate() && g->err.empty()) g->err = "generation failed";
    return g;
}
MP2V_API void mp2v_gen_destroy(mp2v_gen_t* g) { delete g; }
MP2V_API const char* mp2v_gen_error(mp2v_gen_t* g) { return (!g) ? "null generator" : g->err.empty() ? nullptr : g->err.c_str(); }

MP2V_API size_t mp2v_gen_stream(mp2v_gen_t* g, const uint8_t** data) {
    if (!g || g->bw.bytes.size() < 256) return 0;
    if (data) *data = g->bw.bytes.data();
    return g->bw.bytes.size() - 256;
}
MP2V_API size_t mp2v_gen_gop_offset(mp2v_gen_t* g, int gop) {
    if (!g || gop < 0 || gop >= (int)g->gop_off.size()) return 0;
    return g->gop_off[gop];
}
MP2V_API int mp2v_gen_num_pictures(mp2v_gen_t* g) { return g ? (int)g->pics.size() : 0; }
MP2V_API int mp2v_gen_picture(mp2v_gen_t* g, int i, mp2v_gen_picture_t* out) {
    if (!g || !out || i < 0 || i >= (int)g->pics.size()) return MP2V_ERR_ARG;
    const picture_t& s = g->pics[i];
    out->params = s.params;
    out->mb = s.mb.data(); out->coef = s.coef.data();
    out->mb_count = (uint32_t)s.mb.size(); out->n_coef = (uint32_t)s.coef.size();
    out->display_index = s.display_index; out->gop = s.gop;
    out->q_scale_type = s.q_scale_type; out->intra_dc_precision = s.intra_dc_precision;
    memcpy(out->tx, s.tx, sizeof(s.tx));
    memcpy(out->tx_loaded, s.tx_loaded, sizeof(s.tx_loaded));
    return MP2V_OK;
}

}  // extern "C"
