// sm_100a reconstruction kernel: inverse quantisation + mismatch control + the reference's exact
// 16-bit saturating IDCT + half-pel forward / backward / bidirectional motion compensation +
// residual add + clip, fused per macroblock so that a prediction is never written to HBM and read
// back (the CPU reference writes it with mc_pred*/mc_bidir* and re-reads it in
// inverse_dct_template<true>, idct_sse2.hpp:110-114).
//
// Work decomposition (v2, "warp-autonomous"): a warp owns a run of consecutive macroblocks of one
// picture and walks it in batches of <= 32 coded blocks.  There is ONE __syncthreads() per CTA (after
// the picture tables are staged); everything else is __syncwarp(), so warps of a CTA drift apart and
// the memory latency of one overlaps the integer work of the others.  Per batch:
//   1  lanes load the macroblock records (one 16-byte record per lane); a warp scan of popc(cbp)
//      assigns every coded block one of 32 residual-tile slots (144 B each) -- and one IDCT lane
//   2  the first macroblock's reference windows start flowing into shared memory (cp.async)
//   3  dequantise: the warp streams each macroblock's coefficient records with coalesced 32-bit
//      loads, applies parse_block's arithmetic (mb_decoder.cpp:139-146) and scatters int16 values to
//      tile[slot][g_scan_trans[pos]]; mismatch parity of all (<= 12) blocks of a macroblock is one
//      REDUX.XOR per 32 records; a weighted L1 norm per block feeds the saturation bound below
//   4  IDCT: four lanes per coded block, 8 blocks per round, values as packed int16 pairs in the
//      tile (the 8x8 transpose between the passes, idct_sse2.hpp:67-94, is the conflict-free shared
//      memory round trip).  Three bit-exact variants of the lane arithmetic, chosen per round of 8
//      blocks from rigorous range bounds (tools/dev/idct_bounds.py):
//        pass 1  dequantised inputs (|F| <= 2048, first coefficient <= 3036) cannot saturate or wrap
//                anywhere except in the last 8 additions -> plain int32 ops + 8 saturating adds
//        pass 2  if sum |F[k][c]| * Omax[k] * G[c] of every block in the warp stays below the
//                threshold no intermediate can leave int16 -> plain int32 ops; otherwise the exact form
//        exact   every SSE2 lane op reproduced: adds/subs = VIADDMNMX+VIMNMX, mulhi = IMAD+SHF,
//                slli = shift pair with 16-bit wrap
//   5  prediction: per macroblock the (w+1)x(h+1) reference windows of all planes / directions are
//      staged as aligned 16-byte chunks (cp.async, double buffered: the next macroblock's windows load
//      while this one is computed), then every lane produces one output row: funnel-shift
//      realignment, __vavgu4 rounding averages in the reference's order (mc_c.hpp:3-17), residual
//      add + unsigned saturation as VIADDMNMX.S16x2.RELU, one 128-bit (or 64-bit) store.
//
// The roofline that bounds this kernel and the byte counts are in DESIGN.md.
#include "recon_kernels.cuh"

#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "host/scan_tables.h"

namespace mp2v {

__constant__ uint8_t c_scan_trans[2][64];

// Saturation bound weights, 16 * Omax[k] * G[c] rounded up (tools/dev/idct_bounds.py):
// G[c]   = largest |coefficient| input c has in ANY intermediate of one idct_1d_sse2 lane,
// Omax[k] = largest |coefficient| input k has in any OUTPUT of a lane.
// For the second pass: max |intermediate| <= sum_c G[c] * |y_c| + E and |y_c| <= sum_k Omax[k] * |F[k][c]| + E.
__constant__ uint16_t c_bound_w[64] = {
    129, 329, 238, 279, 129, 187, 168, 179,  179, 456, 329, 387, 179, 259, 233, 247,
    168, 430, 310, 364, 168, 244, 220, 233,  179, 456, 329, 387, 179, 259, 233, 247,
    129, 329, 238, 279, 129, 187, 168, 179,  179, 456, 329, 387, 179, 259, 233, 247,
    168, 430, 310, 364, 168, 244, 220, 233,  179, 456, 329, 387, 179, 259, 233, 247 };
// 16 * (32767 - sum(G) * E - E), E = 56.3 (max accumulated mulhi rounding error), sum(G) = 36.0; kept below
constexpr int kBoundLimit = 16 * 30000;
constexpr int kBoundWild = 1 << 28;       // added when a block must take the fully exact path
constexpr int kMaxFirstCoef = 3036;       // (3 * 255 * 127) >> 5: largest unclamped pass-1 input the analysis covers

// Window row pitch: 2 data chunks + 1 chunk of padding.  With 48 bytes the 128-bit row reads of 8
// consecutive rows (a quarter warp) fall into 8 disjoint 4-bank groups (12*r mod 32), i.e. they are
// bank-conflict free; the natural 32-byte pitch measured 4.1 wavefronts per shared load.
constexpr int kWinPitch = 48;

template <int CF>
struct fmt_t {
    static constexpr int NBLK = CF == 1 ? 6 : CF == 2 ? 8 : 12;
    static constexpr int CW = CF == 3 ? 16 : 8;     // chroma macroblock width
    static constexpr int CH = CF == 1 ? 8 : 16;     // chroma macroblock height
    static constexpr int WIN_LUMA = 17 * kWinPitch; // 17 rows x 2 aligned 16-byte chunks (+ 16 B of pitch padding)
    static constexpr int WIN_CHROMA = (CH + 1) * kWinPitch;
    static constexpr int WIN_DIR = WIN_LUMA + 2 * WIN_CHROMA;
    static constexpr int N_ITEMS = 2 * 17 + 2 * 2 * (CH + 1);   // 16-byte chunks per direction
    static constexpr int N_UNITS = 16 + 2 * CH;                 // output rows per macroblock
};

constexpr int kTilePitch = 72;   // int16 per slot: 64 + 8 pad -> 144 B, conflict-free 128-bit row access
#ifndef MP2V_SLOTS
#define MP2V_SLOTS 24      // measured best: 24 slots (4:2:0 4 MBs, 4:2:2 3, 4:4:4 2 all-coded macroblocks fill it exactly)
#endif
#ifndef MP2V_WINBUF
#define MP2V_WINBUF 1      // one window buffer: 29 KB / CTA -> 8 CTAs per SM; two buffers measured 3-6 % slower
#endif
#ifndef MP2V_MINCTAS
#define MP2V_MINCTAS 6
#endif
constexpr int kSlots = MP2V_SLOTS;   // coded blocks per batch (multiple of 8: the IDCT runs 8 blocks per round)
constexpr int kWinBuf = MP2V_WINBUF; // 2: the next macroblock's windows load while this one is computed
constexpr int kWarps = kCtaThreads / 32;

template <int CF>
struct warp_smem_t {
    alignas(16) int16_t tile[kSlots][kTilePitch];
    alignas(16) uint8_t win[kWinBuf][2][fmt_t<CF>::WIN_DIR];   // [buffer][direction]
    int bound[kSlots];
    // per-batch context of the macroblocks that carry records, for the flat dequantisation loop:
    // {first record index in the batch, coef_off, bits, first tile slot}
    alignas(16) uint4 mb_ctx[32];
};

template <int CF>
struct smem_t {
    warp_smem_t<CF> w[kWarps];
    alignas(16) uint8_t W[4][64];
    alignas(16) uint8_t scan[64];
    alignas(16) uint16_t bw[64];
};

// ------------------------------------------------------------------------------------------------
// Exact 16-bit lane arithmetic on sign-extended int32 (idct_sse2.hpp:7-21 helpers).
// Inline PTX on purpose: given C++ min/max on values the compiler knows to be sign-extended int16,
// LLVM canonicalises the clamp into sadd.sat.i16 and the NVPTX back end legalises that into ~10
// instructions.  ptxas turns the PTX below into VIADDMNMX + VIMNMX, IMAD + SHF and SHF/LEA pairs.
__device__ __forceinline__ int adds16(int a, int b) {   // _mm_adds_epi16
    int r;
    asm("{\n\t.reg .s32 t;\n\tadd.s32 t, %1, %2;\n\tmin.s32 t, t, 32767;\n\tmax.s32 %0, t, -32768;\n\t}" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ int subs16(int a, int b) {   // _mm_subs_epi16
    int r;
    asm("{\n\t.reg .s32 t;\n\tsub.s32 t, %1, %2;\n\tmin.s32 t, t, 32767;\n\tmax.s32 %0, t, -32768;\n\t}" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
template <int C> __device__ __forceinline__ int mulhi16(int a) {   // _mm_mulhi_epi16 by a constant
    int r;
    asm("{\n\t.reg .s32 t;\n\tmul.lo.s32 t, %1, %2;\n\tshr.s32 %0, t, 16;\n\t}" : "=r"(r) : "r"(a), "n"(C));
    return r;
}
template <int N> __device__ __forceinline__ int slli16(int a) {   // _mm_slli_epi16: wraps at 16 bits
    int r;
    asm("{\n\t.reg .s32 t;\n\tshl.b32 t, %1, %2;\n\tshr.s32 %0, t, 16;\n\t}" : "=r"(r) : "r"(a), "n"(16 + N));
    return r;
}

// One lane of idct_1d_sse2 (idct_sse2.hpp:23-65), op for op.
//   MODE 0  exact: every add saturates, every shift wraps
//   MODE 1  steps 0-2 in plain int32 (proven in range), the 8 output additions saturate   (pass 1)
//   MODE 2  plain int32 throughout (proven in range by the per-block bound)                (pass 2)
template <int MODE>
__device__ __forceinline__ void idct_lane(int& x0, int& x1, int& x2, int& x3, int& x4, int& x5, int& x6, int& x7) {
    if (MODE == 0) {
        const int v15 = adds16(slli16<1>(mulhi16<27145>(x0)), slli16<1>(x0));
        const int v26 = adds16(mulhi16<-5037>(x1), slli16<2>(x1));
        const int v21 = adds16(mulhi16<-19954>(x2), slli16<2>(x2));
        const int v28 = adds16(slli16<1>(mulhi16<-22089>(x3)), slli16<2>(x3));
        const int v16 = adds16(slli16<1>(mulhi16<27145>(x4)), slli16<1>(x4));
        const int v25 = adds16(mulhi16<14567>(x5), slli16<1>(x5));
        const int v22 = adds16(slli16<1>(mulhi16<17391>(x6)), x6);
        const int v27 = slli16<1>(mulhi16<25570>(x7));
        const int v19 = subs16(v25, v28), v20 = subs16(v26, v27), v23 = adds16(v26, v27), v24 = adds16(v25, v28);
        const int v7 = adds16(v23, v24), v11 = adds16(v21, v22), v13 = subs16(v23, v24), v17 = subs16(v21, v22);
        const int v8 = adds16(v15, v16), v9 = subs16(v15, v16);
        const int v18 = mulhi16<25079>(subs16(v19, v20));
        const int v12 = subs16(v18, adds16(v19, mulhi16<20090>(v19)));
        const int v14 = subs16(subs16(v20, mulhi16<30068>(v20)), v18);
        const int v6 = subs16(slli16<1>(v14), v7);
        const int v5 = subs16(adds16(v13, mulhi16<27145>(v13)), v6);
        const int v4 = adds16(v5, slli16<1>(v12));
        const int v10 = subs16(adds16(v17, mulhi16<27145>(v17)), v11);
        const int v0 = adds16(v8, v11), v1 = adds16(v9, v10), v2 = subs16(v9, v10), v3 = subs16(v8, v11);
        x0 = adds16(v0, v7); x1 = adds16(v1, v6); x2 = adds16(v2, v5); x3 = subs16(v3, v4);
        x4 = adds16(v3, v4); x5 = subs16(v2, v5); x6 = subs16(v1, v6); x7 = subs16(v0, v7);
    } else {
#define MH(a, c) __mulhi((a), (c) * 65536)      // (a*c)>>16 as ONE IMAD.HI on the FMA pipe (frees an ALU-pipe shift)
        const int v15 = (MH(x0, 27145) << 1) + (x0 << 1);
        const int v26 = MH(x1, -5037) + (x1 << 2);
        const int v21 = MH(x2, -19954) + (x2 << 2);
        const int v28 = (MH(x3, -22089) << 1) + (x3 << 2);
        const int v16 = (MH(x4, 27145) << 1) + (x4 << 1);
        const int v25 = MH(x5, 14567) + (x5 << 1);
        const int v22 = (MH(x6, 17391) << 1) + x6;
        const int v27 = MH(x7, 25570) << 1;
        const int v19 = v25 - v28, v20 = v26 - v27, v23 = v26 + v27, v24 = v25 + v28;
        const int v7 = v23 + v24, v11 = v21 + v22, v13 = v23 - v24, v17 = v21 - v22;
        const int v8 = v15 + v16, v9 = v15 - v16;
        const int v18 = MH(v19 - v20, 25079);
        const int v12 = v18 - (v19 + MH(v19, 20090));
        const int v14 = (v20 - MH(v20, 30068)) - v18;
        const int v6 = (v14 << 1) - v7;
        const int v5 = (v13 + MH(v13, 27145)) - v6;
        const int v4 = v5 + (v12 << 1);
        const int v10 = (v17 + MH(v17, 27145)) - v11;
        const int v0 = v8 + v11, v1 = v9 + v10, v2 = v9 - v10, v3 = v8 - v11;
#undef MH
        if (MODE == 1) {
            x0 = adds16(v0, v7); x1 = adds16(v1, v6); x2 = adds16(v2, v5); x3 = subs16(v3, v4);
            x4 = adds16(v3, v4); x5 = subs16(v2, v5); x6 = subs16(v1, v6); x7 = subs16(v0, v7);
        } else {
            x0 = v0 + v7; x1 = v1 + v6; x2 = v2 + v5; x3 = v3 - v4;
            x4 = v3 + v4; x5 = v2 - v5; x6 = v1 - v6; x7 = v0 - v7;
        }
    }
}

// inverse_dct_template up to the >>6 (idct_sse2.hpp:96-107) for 8 tile slots at a time, in place:
// in  F[k*8+c] (the transposed-raster layout parse_block writes), out res[r*8+c].
// FOUR lanes share a block; lane j owns columns 2j,2j+1 in pass 1 and rows 2j,2j+1 in pass 2, always
// as packed int16 pairs (one 32-bit word), so a lane holds 16 values, the 8x8 transpose between the
// passes (transpose_8x8_sse2, idct_sse2.hpp:67-94) is the shared-memory round trip, and with the
// 36-word slot pitch every access pattern below is bank-conflict free:
//   word k*4+j of 8 slots x 4 lanes -> banks 4*slot + 4*k + j, all distinct;
//   128-bit rows 2j, 2j+1           -> quarter-warps cover 8 disjoint 4-bank groups.
// p1_exact / p2_exact are uniform over the warp.
__device__ __forceinline__ void idct_round(int16_t* slot_base, int j, bool active, bool p1_exact, bool p2_exact) {
    uint32_t* t = reinterpret_cast<uint32_t*>(slot_base);
    int a[8], b[8];
    // ---- pass 1: the transform runs across the vector index k for columns 2j and 2j+1
    if (active) {
#pragma unroll
        for (int k = 0; k < 8; k++) { const uint32_t w = t[k * 4 + j]; a[k] = (int)(short)(w & 0xffffu); b[k] = (int)w >> 16; }
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = b[k] = 0;
    }
    if (p1_exact) {
        idct_lane<0>(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
        idct_lane<0>(b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7]);
    } else if (p2_exact) {
        idct_lane<1>(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
        idct_lane<1>(b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7]);
    } else {
        // the bound that clears pass 2 also bounds every pass-1 OUTPUT: |y_c| <= bound / (16 G[c]) + E with
        // G[c] >= 1, i.e. below 30 100 -- the 8 output additions cannot saturate either
        idct_lane<2>(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
        idct_lane<2>(b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7]);
    }
    if (active) {
#pragma unroll
        for (int k = 0; k < 8; k++) t[k * 4 + j] = __byte_perm(a[k], b[k], 0x5410);
    }
    __syncwarp();
    // ---- pass 2: rows 2j and 2j+1 of the first-pass result are lanes 2j, 2j+1 of the transposed block
    if (active) {
        const uint4 r0 = reinterpret_cast<const uint4*>(t)[2 * j], r1 = reinterpret_cast<const uint4*>(t)[2 * j + 1];
        a[0] = (int)(short)(r0.x & 0xffffu); a[1] = (int)r0.x >> 16; a[2] = (int)(short)(r0.y & 0xffffu); a[3] = (int)r0.y >> 16;
        a[4] = (int)(short)(r0.z & 0xffffu); a[5] = (int)r0.z >> 16; a[6] = (int)(short)(r0.w & 0xffffu); a[7] = (int)r0.w >> 16;
        b[0] = (int)(short)(r1.x & 0xffffu); b[1] = (int)r1.x >> 16; b[2] = (int)(short)(r1.y & 0xffffu); b[3] = (int)r1.y >> 16;
        b[4] = (int)(short)(r1.z & 0xffffu); b[5] = (int)r1.z >> 16; b[6] = (int)(short)(r1.w & 0xffffu); b[7] = (int)r1.w >> 16;
    }
    __syncwarp();      // every lane has its rows in registers before anyone overwrites the slot
    if (p2_exact) {
        idct_lane<0>(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
        idct_lane<0>(b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7]);
    } else {
        idct_lane<2>(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
        idct_lane<2>(b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7]);
    }
    // output r of row k is res[r][k]: rows 2j, 2j+1 give the adjacent columns 2j, 2j+1 of every result row
    if (active) {
#pragma unroll
        for (int r = 0; r < 8; r++) t[r * 4 + j] = __byte_perm(a[r] >> 6, b[r] >> 6, 0x5410);   // _mm_srai_epi16(.,6)
    }
}

// ------------------------------------------------------------------------------------------------
// prediction row: NW words (4 pixels each) of one row of a staged window, realigned from byte offset
// o, with the reference's half-pel averaging order (mc_c.hpp:3-17).  The 32 data bytes of a row come
// in as two conflict-free 128-bit loads; the word offset (o >> 2) is resolved by a two-level
// register mux, the byte offset by funnel shifts.
// _mm_avg_epu8 on 4 packed bytes: (a + b + 1) >> 1 = (a | b) - (((a ^ b) & 0xfe..) >> 1), 4 ops (LOP3 fuses xor+and)
__device__ __forceinline__ uint32_t avg4(uint32_t a, uint32_t b) {
    uint32_t t;
    asm("lop3.b32 %0, %1, %2, 0xfefefefe, 0x28;" : "=r"(t) : "r"(a), "r"(b));    // (a ^ b) & c
    return (a | b) - (t >> 1);
}

#ifndef MP2V_WINREAD128
#define MP2V_WINREAD128 0
#endif
template <int NW>
__device__ __forceinline__ void window_words(const uint8_t* win_row, int k, uint32_t (&u)[NW + 1]) {
    if (!MP2V_WINREAD128) {     // NW+1 32-bit loads at the dynamic word offset (at most 2-way conflicts with the 48-byte pitch)
        const uint32_t* p = reinterpret_cast<const uint32_t*>(win_row) + k;
#pragma unroll
        for (int i = 0; i <= NW; i++) u[i] = p[i];
        return;
    }
    const uint4 a = *reinterpret_cast<const uint4*>(win_row), b = *reinterpret_cast<const uint4*>(win_row + 16);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t t[NW + 2];
#pragma unroll
    for (int i = 0; i < NW + 2; i++) t[i] = (k & 2) ? w[i + 2] : w[i];
#pragma unroll
    for (int i = 0; i <= NW; i++) u[i] = (k & 1) ? t[i + 1] : t[i];
}

template <int NW>
__device__ __forceinline__ void pred_row(const uint8_t* win_row, int o, int hx, int hy, uint32_t (&out)[NW]) {
    const int sh = (o & 3) * 8;
    uint32_t w[NW + 1];
    window_words<NW>(win_row, o >> 2, w);
#pragma unroll
    for (int j = 0; j < NW; j++) out[j] = __funnelshift_rc(w[j], w[j + 1], sh);
    if (hx) {
#pragma unroll
        for (int j = 0; j < NW; j++) out[j] = avg4(out[j], __funnelshift_rc(w[j], w[j + 1], sh + 8));
    }
    if (hy) {
        uint32_t b[NW];
        window_words<NW>(win_row + kWinPitch, o >> 2, w);   // next window row
#pragma unroll
        for (int j = 0; j < NW; j++) b[j] = __funnelshift_rc(w[j], w[j + 1], sh);
        if (hx) {
#pragma unroll
            for (int j = 0; j < NW; j++) b[j] = avg4(b[j], __funnelshift_rc(w[j], w[j + 1], sh + 8));
        }
#pragma unroll
        for (int j = 0; j < NW; j++) out[j] = avg4(out[j], b[j]);
    }
}

// pred (4 pixels) + residual (two int16x2 words), unsigned-saturated: packus(adds_epi16(zext(dst), res))
__device__ __forceinline__ uint32_t add_clip4(uint32_t pred, uint32_t r01, uint32_t r23) {
    const uint32_t lo = __vimin_s16x2_relu(__vadd2(__byte_perm(pred, 0, 0x4140), r01), 0x00ff00ffu);
    const uint32_t hi = __vimin_s16x2_relu(__vadd2(__byte_perm(pred, 0, 0x4342), r23), 0x00ff00ffu);
    return __byte_perm(lo, hi, 0x6420);
}

// intra (add=false, idct_sse2.hpp:108-109): packus of the residual alone
__device__ __forceinline__ uint32_t clip4(uint32_t r01, uint32_t r23) {
    return __byte_perm(__vimin_s16x2_relu(r01, 0x00ff00ffu), __vimin_s16x2_relu(r23, 0x00ff00ffu), 0x6420);
}

__device__ __forceinline__ int chroma_mv(int mv, bool halve) { return halve ? (mv >> 1) : mv; }   // floor, mb_decoder.cpp:198-206

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");   // L2 only: window chunks are not re-read through L1
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// Stage the reference windows of one macroblock: aligned 16-byte chunks, 2 per row, (h+1) rows per
// plane.  Lane l always copies chunk (l & 1) of row (l >> 1) of some plane, so its source offset inside
// a plane (lane_off = row * stride + 16 * chunk) is a per-kernel constant and a trip is just
// "plane base + lane_off -> window + 16 * lane".  Trips: luma rows 0-15; chroma rows 0-15 (4:2:0: Cb
// rows 0-7 on lanes 0-15, Cr on lanes 16-31); and, only for vertical half-pel vectors, the extra
// bottom rows.
template <int CF>
__device__ __forceinline__ void stage_windows(const pic_desc_t& pd, const batch_desc_t& batch, uint4 m, int mbx, int mby,
                                              uint8_t* win /* [2][WIN_DIR] */, int lane, int lane_off_y, int lane_off_c) {
    const int lane_woff = (lane >> 1) * kWinPitch + (lane & 1) * 16;      // window position of this lane's chunk
    using F = fmt_t<CF>;
    if (m.y & MP2V_MB_INTRA) return;
#pragma unroll
    for (int d = 0; d < 2; d++) {
        if (!(m.y & (d ? MP2V_MB_BWD : MP2V_MB_FWD))) continue;
        const uint32_t mvw = d ? m.w : m.z;
        const int mvx = (int)(short)(mvw & 0xffffu), mvy = (int)mvw >> 16;
        const int cx = chroma_mv(mvx, CF < 3), cy = chroma_mv(mvy, CF < 2);
        const uint8_t* const* ref = d ? pd.l1 : pd.l0;
        uint8_t* w = win + d * F::WIN_DIR;
        const int x0 = mbx * 16 + (mvx >> 1), y0 = mby * 16 + (mvy >> 1);
        const int cx0 = mbx * F::CW + (cx >> 1), cy0 = mby * F::CH + (cy >> 1);
        const uint8_t* by = ref[0] + (size_t)y0 * batch.stride[0] + (x0 & ~15);
        const size_t coff = (size_t)cy0 * batch.stride[1] + (cx0 & ~15);
        const uint8_t* bcb = ref[1] + coff;
        const uint8_t* bcr = ref[2] + coff;
        cp_async16(w + lane_woff, by + lane_off_y);                                   // luma rows 0..15
        if (CF == 1) {
            const int l = lane & 15;                                                   // Cb rows 0..7 | Cr rows 0..7
            cp_async16(w + F::WIN_LUMA + (lane >> 4) * F::WIN_CHROMA + (l >> 1) * kWinPitch + (l & 1) * 16, (lane < 16 ? bcb : bcr) + (l >> 1) * batch.stride[1] + (l & 1) * 16);
        } else {
            cp_async16(w + F::WIN_LUMA + lane_woff, bcb + lane_off_c);                 // Cb rows 0..15
            cp_async16(w + F::WIN_LUMA + F::WIN_CHROMA + lane_woff, bcr + lane_off_c); // Cr rows 0..15
        }
        if (((mvy | cy) & 1) && lane < 6) {                                            // bottom rows for vertical half-pel
            const int pl = lane >> 1, ch = lane & 1;
            const uint8_t* src = pl == 0 ? by + (size_t)16 * batch.stride[0] : (pl == 1 ? bcb : bcr) + (size_t)F::CH * batch.stride[1];
            uint8_t* dst = w + (pl == 0 ? 16 * kWinPitch : F::WIN_LUMA + (pl - 1) * F::WIN_CHROMA + F::CH * kWinPitch);
            cp_async16(dst + 16 * ch, src + 16 * ch);
        }
    }
}

// prediction + residual + clip + store of one macroblock; windows already staged in `win`.
// Whole-row variant (4:2:0): 16 luma rows of 16 pixels + 2 x 8 chroma rows of 8 pixels = exactly 32 lanes,
// one trip; the luma and chroma half warps run different word counts one after the other.
template <int CF>
__device__ __forceinline__ void reconstruct_mb_rows(const pic_desc_t& pd, const batch_desc_t& batch, uint4 m, int mbx, int mby, int base,
                                               const uint8_t* win, const int16_t (*tile)[kTilePitch], int lane) {
    using F = fmt_t<CF>;
    const uint32_t cbp = MP2V_MB_CBP(m.y);
    const bool fwd = (m.y & MP2V_MB_FWD) != 0, bwd = (m.y & MP2V_MB_BWD) != 0, intra = (m.y & MP2V_MB_INTRA) != 0;
    for (int u = lane; u < F::N_UNITS; u += 32) {
        int p, r;
        if (u < 16) { p = 0; r = u; }
        else if (u < 16 + F::CH) { p = 1; r = u - 16; }
        else { p = 2; r = u - 16 - F::CH; }
        const bool wide = (p == 0) || (CF == 3);
        const int pw = p ? F::CW : 16, ph = p ? F::CH : 16;
        uint32_t pred[4] = {0, 0, 0, 0};
        bool have = false;
        if (!intra) {
#pragma unroll
            for (int d = 0; d < 2; d++) {
                if (!(d ? bwd : fwd)) continue;
                const uint32_t mvw = d ? m.w : m.z;
                const int mvx = (int)(short)(mvw & 0xffffu), mvy = (int)mvw >> 16;
                const int cx = p ? chroma_mv(mvx, CF < 3) : mvx, cy = p ? chroma_mv(mvy, CF < 2) : mvy;
                const int o = (mbx * pw + (cx >> 1)) & 15;
                const uint8_t* row = win + d * F::WIN_DIR + (p == 0 ? 0 : F::WIN_LUMA + (p - 1) * F::WIN_CHROMA) + r * kWinPitch;
                // one four-word path for all 32 lanes: the 8-pixel chroma lanes compute (and drop) two words of
                // padding instead of making the warp run a second, two-word path after the luma one
                uint32_t q[4];
                pred_row<4>(row, o, cx & 1, cy & 1, q);
                if (have) {
#pragma unroll
                    for (int j = 0; j < 4; j++) pred[j] = avg4(q[j], pred[j]);   // bidirectional rounding average
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) pred[j] = q[j];
                }
                have = true;
            }
        }
        // residual blocks covering this row (block geometry: mb_decoder.cpp:177-195)
        int bl, br = -1;
        if (p == 0) { bl = (r >> 3) * 2; br = bl + 1; }
        else if (CF == 1) bl = 3 + p;
        else if (CF == 2) bl = 3 + p + ((r >> 3) << 1);
        else { bl = 3 + p + ((r >> 3) << 1); br = bl + 4; }
        const int rr = r & 7;
        uint32_t out[4] = {pred[0], pred[1], pred[2], pred[3]};
        if (cbp >> bl & 1) {
            const uint4 res = *reinterpret_cast<const uint4*>(&tile[base + __popc(cbp & ((1u << bl) - 1u))][rr * 8]);
            if (intra) { out[0] = clip4(res.x, res.y); out[1] = clip4(res.z, res.w); }      // add=false: packus(res)
            else { out[0] = add_clip4(pred[0], res.x, res.y); out[1] = add_clip4(pred[1], res.z, res.w); }
        }
        uint8_t* drow = pd.dst[p] + (size_t)(mby * ph + r) * batch.stride[p] + mbx * pw;
        if (wide) {
            if (cbp >> br & 1) {
                const uint4 res = *reinterpret_cast<const uint4*>(&tile[base + __popc(cbp & ((1u << br) - 1u))][rr * 8]);
                if (intra) { out[2] = clip4(res.x, res.y); out[3] = clip4(res.z, res.w); }
                else { out[2] = add_clip4(pred[2], res.x, res.y); out[3] = add_clip4(pred[3], res.z, res.w); }
            }
            *reinterpret_cast<uint4*>(drow) = make_uint4(out[0], out[1], out[2], out[3]);
        } else {
            *reinterpret_cast<uint2*>(drow) = make_uint2(out[0], out[1]);
        }
    }
}

// Half-row variant (4:2:2, 4:4:4), same contract.
// The unit of work is an 8-pixel half row = one row of ONE 8x8 block, for luma and chroma alike, so all
// lanes run the same two-word code path (a 16-pixel / 8-pixel split made the luma and chroma half
// warps execute two different paths one after the other).  Units: 32 luma (16 rows x 2 halves), then
// 2 planes x CH rows x (CW/8) halves of chroma.
template <int CF>
__device__ __forceinline__ void reconstruct_mb_halfrows(const pic_desc_t& pd, const batch_desc_t& batch, uint4 m, int mbx, int mby, int base,
                                                        const uint8_t* win, const int16_t (*tile)[kTilePitch], int lane) {
    using F = fmt_t<CF>;
    constexpr int CHALF = F::CW / 8;                       // 8-pixel halves per chroma row
    constexpr int N_UNITS = 32 + 2 * F::CH * CHALF;
    const uint32_t cbp = MP2V_MB_CBP(m.y);
    const bool fwd = (m.y & MP2V_MB_FWD) != 0, bwd = (m.y & MP2V_MB_BWD) != 0, intra = (m.y & MP2V_MB_INTRA) != 0;
#pragma unroll 1
    for (int u = lane; u < N_UNITS; u += 32) {
        int p, r, half;
        if (u < 32) { p = 0; r = u >> 1; half = u & 1; }
        else {
            const int v = u - 32;
            p = 1 + v / (F::CH * CHALF);
            const int w = v - (p - 1) * (F::CH * CHALF);
            r = CHALF == 2 ? w >> 1 : w; half = CHALF == 2 ? w & 1 : 0;
        }
        const int pw = p ? F::CW : 16, ph = p ? F::CH : 16;
        uint32_t pred[2] = {0, 0};
        if (!intra) {
            bool have = false;
#pragma unroll
            for (int d = 0; d < 2; d++) {
                if (!(d ? bwd : fwd)) continue;
                const uint32_t mvw = d ? m.w : m.z;
                const int mvx = (int)(short)(mvw & 0xffffu), mvy = (int)mvw >> 16;
                const int cx = p ? chroma_mv(mvx, CF < 3) : mvx, cy = p ? chroma_mv(mvy, CF < 2) : mvy;
                const int o = ((mbx * pw + (cx >> 1)) & 15) + 8 * half;                 // byte offset of this half row in the window row
                const uint8_t* row = win + d * F::WIN_DIR + (p == 0 ? 0 : F::WIN_LUMA + (p - 1) * F::WIN_CHROMA) + r * kWinPitch;
                uint32_t q[2];
                pred_row<2>(row, o, cx & 1, cy & 1, q);
                if (have) { pred[0] = avg4(q[0], pred[0]); pred[1] = avg4(q[1], pred[1]); }   // bidirectional rounding average
                else { pred[0] = q[0]; pred[1] = q[1]; }
                have = true;
            }
        }
        // the 8x8 block this half row belongs to (block geometry: mb_decoder.cpp:177-195)
        int blk;
        if (p == 0) blk = (r >> 3) * 2 + half;
        else if (CF == 1) blk = 3 + p;
        else if (CF == 2) blk = 3 + p + ((r >> 3) << 1);
        else blk = 3 + p + ((r >> 3) << 1) + 4 * half;
        uint32_t out0 = pred[0], out1 = pred[1];
        if (cbp >> blk & 1) {
            const uint4 res = *reinterpret_cast<const uint4*>(&tile[base + __popc(cbp & ((1u << blk) - 1u))][(r & 7) * 8]);
            if (intra) { out0 = clip4(res.x, res.y); out1 = clip4(res.z, res.w); }          // add=false: packus(res)
            else { out0 = add_clip4(pred[0], res.x, res.y); out1 = add_clip4(pred[1], res.z, res.w); }
        }
        *reinterpret_cast<uint2*>(pd.dst[p] + (size_t)(mby * ph + r) * batch.stride[p] + mbx * pw + 8 * half) = make_uint2(out0, out1);
    }
}

// 4:2:0 fills one trip of whole rows exactly (measured 9 % faster than 1.5 trips of half rows); 4:2:2 and
// 4:4:4 fill 2 / 3 trips of half rows exactly (measured 17 % / 24 % faster than whole rows)
template <int CF>
__device__ __forceinline__ void reconstruct_mb(const pic_desc_t& pd, const batch_desc_t& batch, uint4 m, int mbx, int mby, int base,
                                               const uint8_t* win, const int16_t (*tile)[kTilePitch], int lane) {
    if (CF == 1) reconstruct_mb_rows<CF>(pd, batch, m, mbx, mby, base, win, tile, lane);
    else reconstruct_mb_halfrows<CF>(pd, batch, m, mbx, mby, base, win, tile, lane);
}

template <int CF>
__global__ void __launch_bounds__(kCtaThreads, MP2V_MINCTAS) recon_kernel(const __grid_constant__ batch_desc_t batch) {
    using F = fmt_t<CF>;
    __shared__ smem_t<CF> s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pi = blockIdx.x / batch.ctas_per_pic;
    const int grp = blockIdx.x - pi * batch.ctas_per_pic;
    const pic_desc_t& pd = batch.pic[pi];

    // ---- picture tables (the only CTA-wide barrier)
    if (tid < 64) {
        reinterpret_cast<uint32_t*>(&s.W[0][0])[tid] = reinterpret_cast<const uint32_t*>(&pd.params->W[0][0])[tid];
    } else if (tid < 80) {
        const int alt = pd.params->alternate_scan ? 1 : 0;
        reinterpret_cast<uint32_t*>(s.scan)[tid - 64] = reinterpret_cast<const uint32_t*>(c_scan_trans[alt])[tid - 64];
    } else if (tid < 112) {
        reinterpret_cast<uint32_t*>(s.bw)[tid - 80] = reinterpret_cast<const uint32_t*>(c_bound_w)[tid - 80];
    }
    __syncthreads();

    warp_smem_t<CF>& ws = s.w[warp];
    const int mbw = batch.mbw;
    const int lane_off_y = (lane >> 1) * batch.stride[0] + (lane & 1) * 16;     // window staging: row lane/2, chunk lane&1
    const int lane_off_c = (lane >> 1) * batch.stride[1] + (lane & 1) * 16;
    const int run = batch.mbs_per_warp;
    const int mb_begin = (grp * kWarps + warp) * run;
    const int mb_end = min(mb_begin + run, batch.mb_count);

    uint4 rec_next = (mb_begin + lane < mb_end) ? __ldg(reinterpret_cast<const uint4*>(pd.mb) + mb_begin + lane) : make_uint4(0, 0, 0, 0);
    int mby0 = mb_begin / mbw, mbx0 = mb_begin - mby0 * mbw;      // the only division of the warp; batches advance it
    for (int first = mb_begin; first < mb_end;) {
        // ---- 1. macroblock records of the batch: as many as fit the tile's coded-block slots
        // (lane i holds macroblock first+i; the records were requested during the previous batch)
        const int idx = first + lane;
        const bool have = idx < mb_end;
        uint4 rec = rec_next;
        const int cnt = have ? __popc(MP2V_MB_CBP(rec.y)) : 0;
        const int ncoef_all = have ? (int)MP2V_MB_NCOEF(rec.y) : 0;
        // one warp scan for both prefixes: coded blocks (<= 12 each) in the low half, records (<= 768 each) in the high half
        int scan2 = cnt | (ncoef_all << 16);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, scan2, d);
            if (lane >= d) scan2 += t;
        }
        const int incl = scan2 & 0xffff;
        const int nb = max(__popc(__ballot_sync(0xffffffffu, have && incl <= kSlots)), 1);
        const int base = incl - cnt;
        const int nslots = __shfl_sync(0xffffffffu, incl, nb - 1);
        // request the next batch's records now; they arrive while this batch is processed
        rec_next = (idx + nb < mb_end) ? __ldg(reinterpret_cast<const uint4*>(pd.mb) + idx + nb) : make_uint4(0, 0, 0, 0);

        // ---- 2. first macroblock's windows start loading now; they land while we dequantise and transform
        {
            const uint4 m0 = make_uint4(__shfl_sync(0xffffffffu, rec.x, 0), __shfl_sync(0xffffffffu, rec.y, 0),
                                        __shfl_sync(0xffffffffu, rec.z, 0), __shfl_sync(0xffffffffu, rec.w, 0));
            stage_windows<CF>(pd, batch, m0, mbx0, mby0, &ws.win[0][0][0], lane, lane_off_y, lane_off_c);
            cp_async_commit();
        }

        // ---- 3. zero the used slots (QFS[64] = {0}, mb_decoder.cpp:159), then dequantise + saturate + mismatch.
        // All records of the batch are walked as ONE flat index space (lane = record), so sparse P/B
        // macroblocks do not cost a loop trip each.
        for (int i = lane; i < nslots * 8; i += 32)
            reinterpret_cast<uint4*>(&ws.tile[i >> 3][0])[i & 7] = make_uint4(0, 0, 0, 0);
        if (lane < kSlots) ws.bound[lane] = 0;
        const int ncoef = lane < nb ? ncoef_all : 0;
        const int total = __shfl_sync(0xffffffffu, scan2, nb - 1) >> 16;     // records of the nb macroblocks taken
        const int start = (scan2 >> 16) - ncoef_all;                         // this lane's macroblock: first record index in the batch
        const bool ne = ncoef > 0;
        {   // compact list of the macroblocks that have records
            const uint32_t ne_mask = __ballot_sync(0xffffffffu, ne);
            if (ne) ws.mb_ctx[__popc(ne_mask & ((1u << lane) - 1u))] = make_uint4((uint32_t)start, rec.x, rec.y, (uint32_t)base);
        }
        __syncwarp();
        uint32_t parity = 0;                 // bit s = parity of the coefficient sum of tile slot s
        // Record f of the flat space belongs to the last macroblock whose first record index is <= f.  Per
        // trip of 32 records the owners are found with two warp votes instead of a search per lane: the
        // lanes that HOLD macroblocks mark where theirs starts inside the trip (REDUX.OR) and count the
        // ones that started before it (ballot); a record lane then counts the marks up to itself.
        // Software pipelined: the record of the NEXT trip is requested before the current one is
        // processed, so its global-memory latency is off the critical path.
        auto fetch = [&](int f0, uint32_t& c, uint4& ctx) {
            const int rel = start - f0;
            const uint32_t starts = __reduce_or_sync(0xffffffffu, (ne && (unsigned)rel < 32u) ? 1u << rel : 0u);
            const int before = __popc(__ballot_sync(0xffffffffu, ne && rel < 0));
            c = 0; ctx = make_uint4(0, 0, 0, 0);
            if (f0 + lane < total) {
                ctx = ws.mb_ctx[before + __popc(starts & (0xffffffffu >> (31 - lane))) - 1];
                c = __ldg(pd.coef + ctx.y + (uint32_t)(f0 + lane - (int)ctx.x));
            }
        };
        uint32_t c_nx;
        uint4 ctx_nx;
        fetch(0, c_nx, ctx_nx);
        for (int f0 = 0; f0 < total; f0 += 32) {
            const int f = f0 + lane;
            const uint32_t c = c_nx, m_bits = ctx_nx.z, slot0 = ctx_nx.w;
            fetch(f0 + 32, c_nx, ctx_nx);
            uint32_t pbit = 0;
            if (f < total) {
                const uint32_t cbp = MP2V_MB_CBP(m_bits);
                const int blk = (c >> 22) & 15;
                if (cbp >> blk & 1) {        // a record naming an uncoded block is ignored (memory safety)
                    const bool intra = (m_bits & MP2V_MB_INTRA) != 0;
                    const int qs = MP2V_MB_QSCALE(m_bits);
                    const int level = (int)(short)(c & 0xffffu);
                    const int pos = (c >> 16) & 63;
                    const int slot = (int)slot0 + __popc(cbp & ((1u << blk) - 1u));
                    const bool raw = (c & MP2V_COEF_RAW) != 0, first = (c & MP2V_COEF_FIRST) != 0;
                    const int w = s.W[(blk < 6 ? 0 : 2) + (intra ? 0 : 1)][pos];            // luma matrices for blocks 4,5 (:184-185)
                    const int mag = abs(level);
                    // intra (level*W*qs)>>4, non-intra ((2*level+1)*W*qs)>>5 (:142-143); "1s" is the latter with level 1 (:84)
                    int val = ((intra ? mag : 2 * mag + 1) * w * qs) >> (intra ? 4 : 5);
                    val = level < 0 ? -val : val;                                           // :144
                    const int clamped = max(min((int)(short)val, 2047), -2048);             // int16 wrap, then clamp (:146)
                    val = raw ? level : first ? val : clamped;                              // DC as is (:160); "1s" unclamped (:84)
                    const int idx2 = s.scan[pos];                                           // pos 0 -> 0 for DC / "1s"
                    const int av = abs(val);
                    const int wsum = raw ? (av <= kMaxFirstCoef ? av * (int)s.bw[0] : kBoundWild)
                                         : (av + 1) * (int)s.bw[idx2];   // +1: the mismatch toggle may change |F[63]| by one
                    pbit = raw ? 0u : (uint32_t)(val & 1) << slot;                          // DC is not part of the sum (:160)
                    ws.tile[slot][idx2] = (int16_t)val;
                    atomicAdd(&ws.bound[slot], wsum);
                }
            }
            parity ^= __reduce_xor_sync(0xffffffffu, pbit);
        }
        __syncwarp();
        // qfs[63] ^= (sum & 1) ^ 1 for every coded block (:150-152)
        if (lane < nslots) ws.tile[lane][63] ^= (int16_t)(((parity >> lane) & 1u) ^ 1u);
        __syncwarp();

        // ---- 4. IDCT, four lanes per coded block, 8 blocks per round; arithmetic variant chosen per round
        for (int r0 = 0; r0 < nslots; r0 += 8) {
            const int slot = r0 + (lane >> 2);
            const bool active = slot < nslots;
            const int bnd = active ? ws.bound[slot] + (int)s.bw[63] : 0;    // a toggled-in F[63] = 1 counts too
            const bool p1_exact = __any_sync(0xffffffffu, bnd >= kBoundWild);
            const bool p2_exact = __any_sync(0xffffffffu, bnd > kBoundLimit);
            idct_round(&ws.tile[active ? slot : 0][0], lane & 3, active, p1_exact, p2_exact);
        }
        __syncwarp();

        // ---- 5. prediction + residual + clip + store; with two window buffers the next macroblock's
        // windows load while this one is computed
        int mbx = mbx0, mby = mby0;
        for (int mi = 0; mi < nb; mi++) {
            const uint4 m = make_uint4(__shfl_sync(0xffffffffu, rec.x, mi), __shfl_sync(0xffffffffu, rec.y, mi),
                                       __shfl_sync(0xffffffffu, rec.z, mi), __shfl_sync(0xffffffffu, rec.w, mi));
            const int mbase = __shfl_sync(0xffffffffu, base, mi);
            if (kWinBuf == 2) {
                if (mi + 1 < nb) {
                    const uint4 mn = make_uint4(__shfl_sync(0xffffffffu, rec.x, mi + 1), __shfl_sync(0xffffffffu, rec.y, mi + 1),
                                                __shfl_sync(0xffffffffu, rec.z, mi + 1), __shfl_sync(0xffffffffu, rec.w, mi + 1));
                    const int nbx = mbx + 1 == mbw ? 0 : mbx + 1, nby = mbx + 1 == mbw ? mby + 1 : mby;
                    stage_windows<CF>(pd, batch, mn, nbx, nby, &ws.win[(mi + 1) & (kWinBuf - 1)][0][0], lane, lane_off_y, lane_off_c);
                }
                cp_async_commit();
                cp_async_wait<1>();      // everything but the group just committed has landed: this macroblock's windows
            } else {
                if (mi > 0) { stage_windows<CF>(pd, batch, m, mbx, mby, &ws.win[0][0][0], lane, lane_off_y, lane_off_c); cp_async_commit(); }
                cp_async_wait<0>();
            }
            __syncwarp();
            reconstruct_mb<CF>(pd, batch, m, mbx, mby, mbase, &ws.win[mi & (kWinBuf - 1)][0][0], ws.tile, lane);
            __syncwarp();
            if (++mbx == mbw) { mbx = 0; mby++; }
        }
        cp_async_wait<0>();
        first += nb;
        mbx0 = mbx; mby0 = mby;
    }
}

}  // namespace mp2v

namespace mp2v {
#include "recon_kernel3.cuh"
}

namespace mp2v {

// Which cut of the kernel launches: recon_kernel3 unless MP2V_RECON_KERNEL=2 (development A/B switch).
static int kernel_generation() {
    static const int gen = [] { const char* v = getenv("MP2V_RECON_KERNEL"); return (v && atoi(v) == 2) ? 2 : 3; }();
    return gen;
}

template <int CF>
static cudaError_t prepare_kernel3() {
    return cudaFuncSetAttribute(v3::recon_kernel3<CF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(v3::cta_smem_t<CF>) + 128));
}

// __constant__ symbols and function attributes are per device: initialise each device once (any thread)
static cudaError_t ensure_tables() {
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lk(mu);
    if (!done[dev]) {
        const scan_tables_t& t = scan_tables();
        e = cudaMemcpyToSymbol(c_scan_trans, t.scan_trans, sizeof(t.scan_trans));
        if (e != cudaSuccess) return e;
        if ((e = prepare_kernel3<1>()) != cudaSuccess || (e = prepare_kernel3<2>()) != cudaSuccess || (e = prepare_kernel3<3>()) != cudaSuccess) return e;
        done[dev] = true;
    }
    return cudaSuccess;
}

cudaError_t make_frame_tmaps(int chroma_format, uint8_t* frames, const mp2v_frame_layout_t& lay, size_t frame_alloc, int n_frames, recon_tmaps_t* out) {
    // the driver entry point is resolved at run time: the library links against the CUDA runtime only
    static PFN_cuTensorMapEncodeTiled_v12000 encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) fn = nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }();
    if (!encode) return cudaErrorNotSupported;
    if (chroma_format < 1 || chroma_format > 3 || !out || n_frames < 1) return cudaErrorInvalidValue;
    const cuuint32_t cbox_h = (chroma_format == 1 ? 8 : 16) + 1;      // v3::geo_t: every box is kBoxW = 32 bytes wide
    for (int p = 0; p < 3; p++) {
        const cuuint64_t gdim[3] = {(cuuint64_t)lay.width[p], (cuuint64_t)lay.height[p], (cuuint64_t)n_frames};
        const cuuint64_t gstride[2] = {(cuuint64_t)lay.stride[p], (cuuint64_t)frame_alloc};      // bytes, dimensions 1 and 2
        const cuuint32_t box[3] = {(cuuint32_t)v3::kBoxW, p == 0 ? 17u : cbox_h, 1u};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = encode(&out->plane[p], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, frames + lay.plane_offset[p], gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
    return cudaSuccess;
}

cudaError_t launch_recon(int chroma_format, const batch_desc_t& batch, const recon_tmaps_t& tm, cudaStream_t stream) {
    cudaError_t e = ensure_tables();
    if (e != cudaSuccess) return e;
    if (batch.n_pics < 1 || batch.n_pics > kMaxBatch) return cudaErrorInvalidValue;
    const dim3 grid((unsigned)(batch.n_pics * batch.ctas_per_pic)), block(kCtaThreads);
    if (kernel_generation() == 2) {
        switch (chroma_format) {
            case 1: recon_kernel<1><<<grid, block, 0, stream>>>(batch); break;
            case 2: recon_kernel<2><<<grid, block, 0, stream>>>(batch); break;
            case 3: recon_kernel<3><<<grid, block, 0, stream>>>(batch); break;
            default: return cudaErrorInvalidValue;
        }
    } else {
        switch (chroma_format) {
            case 1: v3::recon_kernel3<1><<<grid, block, sizeof(v3::cta_smem_t<1>) + 128, stream>>>(batch, tm); break;
            case 2: v3::recon_kernel3<2><<<grid, block, sizeof(v3::cta_smem_t<2>) + 128, stream>>>(batch, tm); break;
            case 3: v3::recon_kernel3<3><<<grid, block, sizeof(v3::cta_smem_t<3>) + 128, stream>>>(batch, tm); break;
            default: return cudaErrorInvalidValue;
        }
    }
    return cudaGetLastError();
}

cudaError_t recon_kernel_attributes(int chroma_format, cudaFuncAttributes* out) {
    if (kernel_generation() == 2) {
        switch (chroma_format) {
            case 1: return cudaFuncGetAttributes(out, recon_kernel<1>);
            case 2: return cudaFuncGetAttributes(out, recon_kernel<2>);
            case 3: return cudaFuncGetAttributes(out, recon_kernel<3>);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (chroma_format) {
        case 1: return cudaFuncGetAttributes(out, v3::recon_kernel3<1>);
        case 2: return cudaFuncGetAttributes(out, v3::recon_kernel3<2>);
        case 3: return cudaFuncGetAttributes(out, v3::recon_kernel3<3>);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace mp2v
