// Prefix-indexed decode tables for the Annex B codes, built once from vlc_tables.h.
//
// The reference decodes with leading-zero-count tables generated offline (mp2v_luts.hpp,
// mp2v_vlc_dec.hpp); here every table is a flat "peek N bits -> {length, value}" array generated at
// start-up from the same bit-strings the stream generator encodes with, so encoder and decoder cannot
// drift apart.  Run/level codes use a two-level table (8-bit root, 9-bit leaves) to stay L1-resident.
#pragma once
#include <cstdint>
#include <cstring>

#include "vlc_tables.h"

// The lookup side of these tables is shared by the host slice parser and the device-side one
// (csrc/vlc_kernel.cu): everything is pointer-free so that the whole vlc_decode_tables_t object can be
// copied to the GPU byte for byte.
#ifdef __CUDACC__
#define MP2V_HD __host__ __device__
#else
#define MP2V_HD
#endif

namespace mp2v {

// (entries are aligned to their size so that one look-up is ONE load, also on the device)
struct alignas(4) vlc_entry_t { int8_t len; int8_t pad; int16_t val; };          // len 0 = invalid code

struct alignas(8) coef_entry_t {     // run/level tables
    uint8_t len;          // code length WITHOUT the sign bit; 0 = invalid; root entries with sub != 0 chain to a leaf table
    uint8_t run;
    int16_t level;        // magnitude; kEob / kEscape markers below
    uint16_t sub;         // root only: 1-based leaf table index
    uint16_t pad;
};
constexpr int16_t kCoefEob = -1;
constexpr int16_t kCoefEsc = -2;

// One table look-up = ONE load.  The host compiler does that by itself; the device compiler splits a
// small struct copy into one narrow load per field (and predicates some on others), which doubles the
// latency of the look-up a slice thread spends its life waiting for -- so fetch the entry as a word.
template <class T>
MP2V_HD inline T fetch_entry(const T* p) {
#ifdef __CUDA_ARCH__
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "table entries are one or two words");
    T r;
    if (sizeof(T) == 4) { const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p)); memcpy(&r, &w, 4); }
    else { const uint2 w = __ldg(reinterpret_cast<const uint2*>(p)); memcpy(&r, &w, 8); }
    return r;
#else
    return *p;
#endif
}

template <int BITS>
struct flat_vlc_t {
    vlc_entry_t e[1 << BITS];
    flat_vlc_t() { memset(e, 0, sizeof(e)); }
    void add(const char* bits, int val) {
        const int len = (int)strlen(bits);
        uint32_t code = 0;
        for (int i = 0; i < len; i++) code = (code << 1) | (uint32_t)(bits[i] == '1');
        const uint32_t lo = code << (BITS - len), n = 1u << (BITS - len);
        for (uint32_t k = 0; k < n; k++) { e[lo + k].len = (int8_t)len; e[lo + k].val = (int16_t)val; }
    }
    MP2V_HD inline vlc_entry_t look(uint32_t peek_bits) const { return fetch_entry(&e[peek_bits]); }
};

// Fast path of the run/level decoder: one lookup on the next 11 bits resolves code AND sign for every symbol whose
// code + sign bit fit (all the frequent ones) into a word that is added to the running record (slice_core.h);
// anything else is an escape or one of the long codes below.
constexpr int kFastBits = 11;

struct coef_vlc_t {
    static constexpr int ROOT = 8, LEAF = 9;                      // longest code is 16 bits (+ sign)
    // A fast symbol's coefficient record is `q + (entry & 0x7fffff)` where q carries block and position:
    //   [15:0] level (two's complement)  [22:16] run + 1  [27:24] bits consumed (incl. sign)
    //   [31] not a fast symbol -> [30] end of block (then [27:24] = its length), else an escape or a long code
    uint32_t gpu_fast[1 << kFastBits];
    // ... and the codes the fast table does not hold: besides the escape (prefix 000001) they all start with seven
    // zeros and are 12..16 bits long, so the ten bits behind the zeros resolve code and sign:
    //   [15:0] level  [22:16] run + 1  [28:24] bits consumed (incl. the zeros and the sign)  [31] no such code
    static constexpr int kLongZeros = 7, kLongBits = 10;
    uint32_t gpu_long[1 << kLongBits];
    void build_fast() {
        for (uint32_t i = 0; i < (1u << kLongBits); i++) {
            const coef_entry_t e = look(i);                       // a 17-bit window whose top seven bits are zero
            gpu_long[i] = 0x80000000u;
            if (e.len > kLongZeros && e.len <= 16 && e.level > 0) {
                const int neg = (int)(i >> (16 - e.len)) & 1;
                gpu_long[i] = (uint32_t)(uint16_t)(neg ? -e.level : e.level) | ((uint32_t)(e.run + 1) << 16) | ((uint32_t)(e.len + 1) << 24);
            }
        }
        for (uint32_t i = 0; i < (1u << kFastBits); i++) {
            const coef_entry_t e = look(i << (17 - kFastBits));
            gpu_fast[i] = 0x80000000u;
            if (e.len && e.level > 0 && e.len + 1 <= kFastBits) {
                const int neg = (int)(i >> (kFastBits - 1 - e.len)) & 1;
                gpu_fast[i] = (uint32_t)(uint16_t)(neg ? -e.level : e.level) | ((uint32_t)(e.run + 1) << 16) | ((uint32_t)(e.len + 1) << 24);
            } else if (e.len && e.level == kCoefEob && e.len <= kFastBits) {
                gpu_fast[i] = 0xc0000000u | ((uint32_t)e.len << 24);
            }
        }
    }
    static constexpr int MAX_LEAVES = 8;                          // distinct 8-bit prefixes with longer codes (B.14/B.15 need 6)
    coef_entry_t root[1 << ROOT];
    coef_entry_t leaves[MAX_LEAVES << LEAF];                      // (1 << LEAF) entries per leaf table
    int n_leaves = 0;
    coef_vlc_t() { memset(root, 0, sizeof(root)); memset(leaves, 0, sizeof(leaves)); }
    void add(const char* bits, int run, int level) {
        const int len = (int)strlen(bits);
        uint32_t code = 0;
        for (int i = 0; i < len; i++) code = (code << 1) | (uint32_t)(bits[i] == '1');
        if (len <= ROOT) {
            const uint32_t lo = code << (ROOT - len), n = 1u << (ROOT - len);
            for (uint32_t k = 0; k < n; k++) { root[lo + k].len = (uint8_t)len; root[lo + k].run = (uint8_t)run; root[lo + k].level = (int16_t)level; root[lo + k].sub = 0; }
        } else {
            const uint32_t prefix = code >> (len - ROOT);
            if (!root[prefix].sub) {
                if (n_leaves >= MAX_LEAVES) return;               // cannot happen with the Annex B tables (checked in tests)
                root[prefix].sub = (uint16_t)(++n_leaves);
                root[prefix].len = 0;
            }
            coef_entry_t* leaf = &leaves[(size_t)(root[prefix].sub - 1) << LEAF];
            const int rem = len - ROOT;
            const uint32_t low = code & ((1u << rem) - 1u);
            const uint32_t lo = low << (LEAF - rem), n = 1u << (LEAF - rem);
            for (uint32_t k = 0; k < n; k++) { leaf[lo + k].len = (uint8_t)len; leaf[lo + k].run = (uint8_t)run; leaf[lo + k].level = (int16_t)level; }
        }
    }
    // peek17 = next 17 bits of the stream
    MP2V_HD inline coef_entry_t look(uint32_t peek17) const {
        const coef_entry_t r = fetch_entry(&root[peek17 >> (17 - ROOT)]);
        if (!r.sub) return r;
        return fetch_entry(&leaves[((size_t)(r.sub - 1) << LEAF) + (peek17 & ((1u << LEAF) - 1u))]);
    }
};

// dct_dc_size + dct_dc_differential in one lookup on the next 12 bits (covers sizes whose code and
// differential bits fit together -- every small differential); len == 0 -> take the two-step path
struct alignas(4) dc_fast_t { int16_t diff; uint8_t len; uint8_t pad; };
constexpr int kDcFastBits = 12;

struct vlc_decode_tables_t {
    dc_fast_t dc_fast[2][1 << kDcFastBits];
    flat_vlc_t<11> mba;            // val = increment, 0 = escape (adds 33)
    flat_vlc_t<6> mbtype[4];       // per picture_coding_type; val = flag byte
    flat_vlc_t<9> cbp;
    flat_vlc_t<10> motion;         // magnitude 0..16 (sign bit follows when != 0)
    flat_vlc_t<10> dcsize[2];      // 0 luminance, 1 chrominance
    coef_vlc_t b14, b15;

    vlc_decode_tables_t() {
        for (int i = 0; i < MP2V_COUNT(kTabMbAddrInc); i++) mba.add(kTabMbAddrInc[i].bits, kTabMbAddrInc[i].a);
        for (int i = 0; i < MP2V_COUNT(kTabMbType); i++) mbtype[kTabMbType[i].b].add(kTabMbType[i].bits, kTabMbType[i].a);
        for (int i = 0; i < MP2V_COUNT(kTabCbp); i++) cbp.add(kTabCbp[i].bits, kTabCbp[i].a);
        for (int i = 0; i < MP2V_COUNT(kTabMotionCode); i++) motion.add(kTabMotionCode[i].bits, kTabMotionCode[i].a);
        for (int i = 0; i < MP2V_COUNT(kTabDcSize); i++) dcsize[kTabDcSize[i].b].add(kTabDcSize[i].bits, kTabDcSize[i].a);
        for (int i = 0; i < MP2V_COUNT(kTabCoefB14); i++) b14.add(kTabCoefB14[i].bits, kTabCoefB14[i].a, kTabCoefB14[i].b);
        for (int i = 0; i < MP2V_COUNT(kTabCoefB15); i++) b15.add(kTabCoefB15[i].bits, kTabCoefB15[i].a, kTabCoefB15[i].b);
        b14.add(kEobB14, 0, kCoefEob); b15.add(kEobB15, 0, kCoefEob);
        b14.add(kCoefEscape, 0, kCoefEsc); b15.add(kCoefEscape, 0, kCoefEsc);
        b14.build_fast(); b15.build_fast();
        for (int lc = 0; lc < 2; lc++)
            for (uint32_t i = 0; i < (1u << kDcFastBits); i++) {
                const vlc_entry_t e = dcsize[lc].look(i >> (kDcFastBits - 10));
                dc_fast_t f{0, 0, 0};
                if (e.len && e.len + e.val <= kDcFastBits) {
                    int diff = 0;
                    if (e.val) {                                   // mb_decoder.cpp:59-68
                        const int v = (int)(i >> (kDcFastBits - e.len - e.val)) & ((1 << e.val) - 1);
                        const int half = 1 << (e.val - 1);
                        diff = v >= half ? v : v + 1 - 2 * half;
                    }
                    f.diff = (int16_t)diff; f.len = (uint8_t)(e.len + e.val);
                }
                dc_fast[lc][i] = f;
            }
    }
};

inline const vlc_decode_tables_t& vlc_decode_tables() {
    static const vlc_decode_tables_t t;
    return t;
}

}  // namespace mp2v
