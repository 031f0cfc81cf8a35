// sm_100a reconstruction kernel: inverse quantisation + mismatch control + the reference's exact
// 16-bit saturating IDCT + half-pel forward / backward / bidirectional motion compensation +
// residual add + clip, fused per macroblock so that a prediction is never written to HBM and read
// back (the CPU reference writes it with mc_pred*/mc_bidir* and re-reads it in
// inverse_dct_template<true>, idct_sse2.hpp:110-114).
//
// Work decomposition ("warp-autonomous"): a warp owns a run of consecutive macroblocks of one picture
// and walks it in batches of <= 24 coded blocks / <= 16 macroblocks of one macroblock row.  There is
// ONE __syncthreads() per CTA (after the picture tables are staged); everything else is __syncwarp()
// and per-warp mbarriers, so the warps of a CTA drift apart and the memory latency of one overlaps the
// integer work of the others.  Per batch:
//   1  lanes load the macroblock records (one 16-byte record per lane, requested during the previous
//      batch); a warp scan of popc(cbp) assigns every coded block a 144-byte slot of the warp's
//      residual tile; lane 0 issues the first macroblock's reference boxes
//   2  dequantise: the coefficient records of the batch are ONE contiguous list (lane = record,
//      coalesced 32-bit loads, the next trip requested before this one is processed).  A record names
//      its macroblock by the column tag in bits 31:28 (include/mp2v_recon.h), so its context
//      {cbp, first slot, quantiser scale, shift} is one 8-byte shared-memory load.  parse_block's
//      arithmetic (mb_decoder.cpp:139-146), scatter to tile[slot][g_scan_trans[pos]]; mismatch parity
//      of all slots is one REDUX.XOR per 32 records; a weighted L1 norm per block feeds the
//      saturation bound below.  A mismatch toggle that only turns F[63] from 0 into 1 is dropped:
//      column 7 of pass 1 maps x7 = 1 to (1 * 25570) >> 16 = 0, an all-zero column
//      (idct_sse2.hpp:32,41) -- the block is bit-identical without it
//   3  IDCT: one lane per column in pass 1 and one lane per row in pass 2 (8 lanes per block, 4
//      blocks per trip); 16-bit shared-memory loads sign-extend, 16-bit stores pack, pass 2 reads its
//      row with one conflict-free 128-bit load -- the 8x8 transpose between the passes
//      (idct_sse2.hpp:67-94) is that shared-memory round trip.  Three bit-exact variants of the lane
//      arithmetic, chosen per batch from rigorous range bounds (tools/dev/idct_bounds.py):
//        pass 1  dequantised inputs (|F| <= 2048, first coefficient <= 3036) cannot saturate or wrap
//                anywhere except in the last 8 additions -> plain int32 ops + 8 saturating adds
//        pass 2  if sum |F[k][c]| * Omax[k] * G[c] of every block in the batch stays below the
//                threshold no intermediate can leave int16 -> plain int32 ops; otherwise the exact form
//        exact   every SSE2 lane op reproduced: adds/subs = VIADDMNMX+VIMNMX, mulhi = IMAD+SHF,
//                slli = shift pair with 16-bit wrap
//   4  prediction: per macroblock, ONE TMA box load per plane and direction
//      (cp.async.bulk.tensor.3d over the frame pool viewed as [frame][row][pixel], completion on a
//      per-warp mbarrier).  A box must start on a 16-byte boundary of the innermost dimension
//      (measured: any other x coordinate raises an illegal-instruction fault), so a box is 32 bytes
//      wide from (x & ~15) and its rows are read at the byte offset x & 15.  Boxes that leave the
//      plane are zero-filled by the TMA unit: a bad vector cannot fault.  Every lane then produces one
//      output row (4:2:0; half rows for the wider formats): 32-bit loads at the dynamic word offset,
//      funnel-shift realignment, packed rounding averages in the reference's order (mc_c.hpp:3-17),
//      residual add + unsigned saturation as VIADDMNMX.S16x2.RELU, one 128-bit (or 64-bit) store.
//      Every lane owns the same (plane, row) of every macroblock: block indices, box offsets and the
//      destination pointer are per-lane constants / running pointers.
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

constexpr int kTilePitch = 72;   // int16 per slot: 64 + 8 pad -> 144 B, conflict-free 128-bit row access

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

// _mm_avg_epu8 on 4 packed bytes: (a + b + 1) >> 1 = (a | b) - (((a ^ b) & 0xfe..) >> 1), 4 ops (LOP3 fuses xor+and)
__device__ __forceinline__ uint32_t avg4(uint32_t a, uint32_t b) {
    uint32_t t;
    asm("lop3.b32 %0, %1, %2, 0xfefefefe, 0x28;" : "=r"(t) : "r"(a), "r"(b));    // (a ^ b) & c
    return (a | b) - (t >> 1);
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



#ifndef MP2V_V3_WINBUF
#define MP2V_V3_WINBUF 1
#endif
#ifndef MP2V_V3_MINCTAS
#define MP2V_V3_MINCTAS 8
#endif
constexpr int kSlots = 24;                 // coded blocks per batch
constexpr int kTileRows = kSlots + 1;      // + one spare row: a record naming an uncoded block lands there at worst
constexpr int kWinBuf = MP2V_V3_WINBUF;    // window buffers per warp (2: the next macroblock's boxes load during this one)
constexpr int kWarps = kCtaThreads / 32;
constexpr int kBoxW = 32;                  // bytes per box row: 16-byte aligned start + up to 15 bytes of offset + 17 pixels

constexpr int align128(int x) { return (x + 127) & ~127; }

template <int CF>
struct geo_t {
    static constexpr int NBLK = CF == 1 ? 6 : CF == 2 ? 8 : 12;
    static constexpr int CW = CF == 3 ? 16 : 8;      // chroma macroblock width
    static constexpr int CH = CF == 1 ? 8 : 16;      // chroma macroblock height
    static constexpr int Y_BYTES = kBoxW * 17, C_BYTES = kBoxW * (CH + 1);
    static constexpr int CB_OFF = align128(Y_BYTES), CR_OFF = CB_OFF + align128(C_BYTES);
    static constexpr int DIR_BYTES = CR_OFF + align128(C_BYTES);      // every box 128-byte aligned
    static constexpr uint32_t TX_BYTES = Y_BYTES + 2 * C_BYTES;      // bytes one direction's three boxes deliver
};

template <int CF>
struct alignas(128) warp_smem_t {
    alignas(128) uint8_t win[kWinBuf][2][geo_t<CF>::DIR_BYTES];   // [buffer][direction]
    alignas(16) int16_t tile[kTileRows][kTilePitch];
    // the (up to 16) macroblocks of the batch, by position: {cbp | W row offset << 16, first slot | qscale << 8 | shift << 16}
    alignas(16) uint2 mb_ctx[16];
    alignas(16) int bound[kTileRows + 3];  // saturation bound per slot
    alignas(8) uint64_t mbar[kWinBuf];
};

template <int CF>
struct cta_smem_t {
    warp_smem_t<CF> w[kWarps];
    alignas(16) uint8_t W[4][64];          // quantiser matrices by scan position
    alignas(16) uint8_t scan[64];          // scan position -> tile index (g_scan_trans)
    alignas(16) uint16_t bwp[64];          // bound weight by scan position
};

// ---- mbarrier / TMA wrappers (PTX ISA 8.6, sm_100a)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded: a box that never arrives traps instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spins = 0; spins < (1 << 22); spins++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
// one box of a [frame][row][pixel] plane tensor; coordinates in elements, innermost first; x must be a multiple of 16
__device__ __forceinline__ void tma_box_3d(void* dst, const CUtensorMap* tm, int x, int y, int z, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

// prediction of one unit (NW words = 4 * NW pixels) from a staged box row at byte offset o; half-pel
// averaging in the reference's order (mc_c.hpp:3-17): H = avg(p[x], p[x+1]), V = avg(p[x], p[x+stride]),
// HV = avg(H(row), H(row + 1)).  32-bit loads at the dynamic word offset, funnel shifts for the byte part.
template <int NW>
__device__ __forceinline__ void pred_unit(const uint8_t* row, int o, int hx, int hy, uint32_t (&out)[NW]) {
    const int sh = (o & 3) * 8;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(row + (o & ~3));
    uint32_t w[NW + 1];
#pragma unroll
    for (int i = 0; i <= NW; i++) w[i] = p[i];
#pragma unroll
    for (int j = 0; j < NW; j++) out[j] = __funnelshift_rc(w[j], w[j + 1], sh);
    if (hx) {
#pragma unroll
        for (int j = 0; j < NW; j++) out[j] = avg4(out[j], __funnelshift_rc(w[j], w[j + 1], sh + 8));
    }
    if (hy) {
        uint32_t b[NW];
#pragma unroll
        for (int i = 0; i <= NW; i++) w[i] = p[i + kBoxW / 4];      // next box row
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

// residual of block `blk` (row rr) on top of two words of prediction, or alone for intra macroblocks
__device__ __forceinline__ void add_residual(const int16_t (*tile)[kTilePitch], int base, uint32_t cbp, uint32_t below, int blk, int rr8, bool intra,
                                             uint32_t& o0, uint32_t& o1) {
    if (cbp >> blk & 1) {
        const uint4 res = *reinterpret_cast<const uint4*>(&tile[base + __popc(cbp & below)][rr8]);
        if (intra) { o0 = clip4(res.x, res.y); o1 = clip4(res.z, res.w); }      // add=false: packus(res), idct_sse2.hpp:108-109
        else { o0 = add_clip4(o0, res.x, res.y); o1 = add_clip4(o1, res.z, res.w); }
    }
}

template <int CF>
__global__ void __launch_bounds__(kCtaThreads, MP2V_V3_MINCTAS)
recon_kernel3(const __grid_constant__ batch_desc_t batch, const __grid_constant__ recon_tmaps_t tm) {
    using G = geo_t<CF>;
    extern __shared__ uint8_t smem_raw[];
    cta_smem_t<CF>& s = *reinterpret_cast<cta_smem_t<CF>*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pi = blockIdx.x / batch.ctas_per_pic;
    const int grp = blockIdx.x - pi * batch.ctas_per_pic;
    const pic_desc_t& pd = batch.pic[pi];
    warp_smem_t<CF>& ws = s.w[warp];

    // ---- picture tables + this warp's barriers (the only CTA-wide barrier)
    if (tid < 64) {
        reinterpret_cast<uint32_t*>(&s.W[0][0])[tid] = reinterpret_cast<const uint32_t*>(&pd.params->W[0][0])[tid];
        const int alt = pd.params->alternate_scan ? 1 : 0;
        const int t = c_scan_trans[alt][tid];
        s.scan[tid] = (uint8_t)t;
        s.bwp[tid] = c_bound_w[t];
    }
    if (lane == 0) {
#pragma unroll
        for (int b = 0; b < kWinBuf; b++) mbar_init(&ws.mbar[b], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int mbw = batch.mbw;
    const int run = batch.mbs_per_warp;
    const int mb_begin = (grp * kWarps + warp) * run;
    const int mb_end = min(mb_begin + run, batch.mb_count);
    if (mb_begin >= mb_end) return;

    // ---- per-lane constants of the output stage.  4:2:0: lane = one whole row (16 luma rows, 8 Cb, 8 Cr: one
    // trip of four-word units, the chroma lanes drop two words).  4:2:2 / 4:4:4: trips of 8-pixel half rows.
    constexpr int NT = CF == 1 ? 1 : CF == 2 ? 2 : 3;      // trips per macroblock
    constexpr int NW = CF == 1 ? 4 : 2;                    // words per unit
    int u_woff[NT], u_xoff[NT], u_blk[NT], u_rr8[NT], u_adv[NT], u_wrap[NT], u_field[NT];
    uint32_t u_below[NT];
    uint8_t* u_dst[NT];
    bool u_chroma[NT];
    int mby0 = mb_begin / mbw, mbx0 = mb_begin - mby0 * mbw;      // the only division of the warp
#pragma unroll
    for (int t = 0; t < NT; t++) {
        int p, r, half;
        if (CF == 1) { p = lane < 16 ? 0 : lane < 24 ? 1 : 2; r = lane < 16 ? lane : (lane & 7); half = 0; }
        else if (CF == 2) { p = t == 0 ? 0 : 1 + (lane >> 4); r = t == 0 ? lane >> 1 : lane & 15; half = t == 0 ? lane & 1 : 0; }
        else { p = t; r = lane >> 1; half = lane & 1; }
        u_chroma[t] = p != 0;
        u_woff[t] = (p == 0 ? 0 : p == 1 ? G::CB_OFF : G::CR_OFF) + r * kBoxW;
        u_xoff[t] = 8 * half;
        // the 8x8 block this unit's (left) half belongs to (block geometry: mb_decoder.cpp:177-195)
        int blk;
        if (p == 0) blk = (r >> 3) * 2 + half;
        else if (CF == 1) blk = 3 + p;
        else if (CF == 2) blk = 3 + p + ((r >> 3) << 1);
        else blk = 3 + p + ((r >> 3) << 1) + 4 * half;
        u_blk[t] = blk;
        u_below[t] = (1u << blk) - 1u;
        u_rr8[t] = (r & 7) * 8;
        // dct_type = 1 (mb_decoder.cpp:172-195): the two blocks above one another hold the two fields of their 16 rows:
        // frame row r is row r >> 1 of the upper (even r) or lower (odd r) block.  4:2:0 chroma is frame organised.
        int fblk = blk;
        if (p == 0) fblk = (r & 1) * 2 + half;
        else if (CF == 2) fblk = 3 + p + ((r & 1) << 1);
        else if (CF == 3) fblk = 3 + p + ((r & 1) << 1) + 4 * half;
        u_field[t] = (p == 0 || CF != 1) ? (fblk | ((r >> 1) * 8) << 8) : (blk | u_rr8[t] << 8);
        const int pw = p ? G::CW : 16, ph = p ? G::CH : 16;
        u_dst[t] = pd.dst[p] + (size_t)(mby0 * ph + r) * batch.stride[p] + mbx0 * pw + 8 * half;
        u_adv[t] = pw;                                                     // destination step to the next macroblock of the row ...
        u_wrap[t] = ph * batch.stride[p] - (mbw - 1) * pw;                 // ... and from the last one to the first of the next row
    }

    uint32_t wphase = 0;                                   // bit b: parity the next wait on buffer b expects
    uint32_t path_counts = 0;                              // batches | exact pass 2 << 8 | exact pass 1 << 16
    uint4 rec_next = (mb_begin + lane < mb_end) ? __ldg(reinterpret_cast<const uint4*>(pd.mb) + mb_begin + lane) : make_uint4(0, 0, 0, 0);
    // the 32 coefficient records behind the previous batch's last one: the next batch's first trip when its list continues there
    uint32_t pref_idx = 0xffffffffu, pref_c = 0;

    // boxes of one macroblock (its record broadcast in m_*), issued by one lane into window buffer `buf`
    auto issue_windows = [&](uint32_t m_y, uint32_t m_z, uint32_t m_w, int ix, int iy, int buf) {
        if ((m_y & MP2V_MB_INTRA) || lane != 0) return;
        const uint32_t ndir = ((m_y & MP2V_MB_FWD) ? 1u : 0u) + ((m_y & MP2V_MB_BWD) ? 1u : 0u);
        mbar_expect_tx(&ws.mbar[buf], ndir * G::TX_BYTES);
#pragma unroll
        for (int d = 0; d < 2; d++) {
            if (!(m_y & (d ? MP2V_MB_BWD : MP2V_MB_FWD))) continue;
            const uint32_t mvw = d ? m_w : m_z;
            const int mvx = (int)(short)(mvw & 0xffffu), mvy = (int)mvw >> 16;
            const int cx = CF < 3 ? mvx >> 1 : mvx, cy = CF < 2 ? mvy >> 1 : mvy;     // chroma vector: floor (mb_decoder.cpp:198-206)
            const int z = d ? pd.l1_id : pd.l0_id;
            const int cxa = (ix * G::CW + (cx >> 1)) & ~15, cya = iy * G::CH + (cy >> 1);
            uint8_t* w = &ws.win[buf][d][0];
            tma_box_3d(w, &tm.plane[0], (ix * 16 + (mvx >> 1)) & ~15, iy * 16 + (mvy >> 1), z, &ws.mbar[buf]);
            tma_box_3d(w + G::CB_OFF, &tm.plane[1], cxa, cya, z, &ws.mbar[buf]);
            tma_box_3d(w + G::CR_OFF, &tm.plane[2], cxa, cya, z, &ws.mbar[buf]);
        }
    };

    int16_t* const tile0 = &ws.tile[0][0];
    for (int first = mb_begin; first < mb_end;) {
        // ---- 1. macroblock records of the batch: as many as fit the tile's coded-block slots, at most 16, inside one
        // macroblock row (a coefficient record names its macroblock by its column modulo 16), with their coefficient
        // records contiguous in the arena (they are inside a slice; anything else only shortens the batch)
        const int idx = first + lane;
        const bool have = idx < mb_end;
        uint4 rec = rec_next;
        const uint32_t field_mbs = __ballot_sync(0xffffffffu, (rec.x & MP2V_MB_FIELD_DCT) != 0);      // bit i: macroblock i of the batch is field-DCT coded
        rec.x = MP2V_MB_COEF_OFF(rec.x);
        const int cnt = have ? __popc(MP2V_MB_CBP(rec.y)) : 0;
        const int ncoef_all = have ? (int)MP2V_MB_NCOEF(rec.y) : 0;
        int scan2 = cnt | (ncoef_all << 16);       // both prefixes in one scan: coded blocks (low half), records (high half)
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, scan2, d);
            if (lane >= d) scan2 += t;
        }
        const int incl = scan2 & 0xffff;
        const int start = (scan2 >> 16) - ncoef_all;                     // this macroblock's first record in the batch's flat list
        int nb = max(__popc(__ballot_sync(0xffffffffu, have && incl <= kSlots && lane < 16 && lane < mbw - mbx0)), 1);
        uint32_t base_off = 0;                                           // arena index of flat record 0
        bool any_records;
        {
            const uint32_t ne_mask = __ballot_sync(0xffffffffu, lane < nb && ncoef_all > 0);
            any_records = ne_mask != 0;
            const int f_ne = ne_mask ? __ffs(ne_mask) - 1 : 0;
            base_off = __shfl_sync(0xffffffffu, rec.x - (uint32_t)start, f_ne);
            const uint32_t gap = __ballot_sync(0xffffffffu, lane < nb && ncoef_all > 0 && rec.x - (uint32_t)start != base_off);
            if (gap) nb = __ffs(gap) - 1;                                // (>= 1: the first macroblock with records defines base_off)
        }
        const int base = incl - cnt;
        const int nslots = __shfl_sync(0xffffffffu, incl, nb - 1);
        const int total = __shfl_sync(0xffffffffu, scan2, nb - 1) >> 16;
        rec_next = (idx + nb < mb_end) ? __ldg(reinterpret_cast<const uint4*>(pd.mb) + idx + nb) : make_uint4(0, 0, 0, 0);

        // ---- 2. the first trip's records on their way (usually already requested during the previous batch), and the
        // per-macroblock context of the dequantisation loop: {cbp | W row offset, first slot | qscale << 8 | shift << 16}
        uint32_t c_nx = 0;
        if (lane < total) c_nx = (base_off + (uint32_t)lane == pref_idx) ? pref_c : __ldg(pd.coef + base_off + lane);
        if (lane < 16) {
            const uint32_t ni = (rec.y & MP2V_MB_INTRA) ? 0u : 1u;
            ws.mb_ctx[lane] = lane < nb ? make_uint2(MP2V_MB_CBP(rec.y) | (ni << 22), (uint32_t)base | (MP2V_MB_QSCALE(rec.y) << 8) | ((4u + ni) << 16))
                                        : make_uint2(0u, 0u);            // (a record with a wrong column tag lands in slot 0 of the batch: garbage in, no fault)
        }

        // ---- 3. the first macroblocks' boxes start loading now; they land while we dequantise and transform
        {
            int ix = mbx0, iy = mby0;
#pragma unroll
            for (int b = 0; b < kWinBuf; b++) {
                if (b < nb) issue_windows(__shfl_sync(0xffffffffu, rec.y, b), __shfl_sync(0xffffffffu, rec.z, b), __shfl_sync(0xffffffffu, rec.w, b), ix, iy, b);
                if (++ix == mbw) { ix = 0; iy++; }
            }
        }

        // ---- 4. clear the used slots (QFS[64] = {0}, mb_decoder.cpp:159) and bounds; dequantise.  All records of the
        // batch are ONE flat list (lane = record), so sparse P/B macroblocks do not cost a loop trip each.
        for (int i = lane; i < nslots * 8; i += 32) reinterpret_cast<uint4*>(tile0)[(i >> 3) * (kTilePitch / 8) + (i & 7)] = make_uint4(0, 0, 0, 0);
        if (lane < 7) reinterpret_cast<uint4*>(ws.bound)[lane] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        uint32_t parity = 0;                 // bit s = parity of the coefficient sum of tile slot s
        uint32_t col7 = 0;                   // bit s = column 7 of slot s holds a coefficient
        const uint32_t mb0 = (uint32_t)mbx0 & 15u;
        for (int f0 = 0; f0 < total; f0 += 32) {
            const uint32_t c = c_nx;
            const bool live = f0 + lane < total;
            c_nx = (f0 + 32 + lane < total) ? __ldg(pd.coef + base_off + (uint32_t)(f0 + 32 + lane)) : 0u;      // next trip's record
            uint32_t pbit = 0, c7bit = 0;
            if (live) {
                const uint2 ctx = ws.mb_ctx[((c >> 28) - mb0) & 15u];
                const uint32_t cz = ctx.x, pk = ctx.y;
                const int blk = (c >> 22) & 15;
                const int slot = (int)(pk & 0xffu) + __popc(cz & 0xfffu & ((1u << blk) - 1u));      // <= kSlots even for a record naming an uncoded block
                const int qs = (pk >> 8) & 0xff, sh = pk >> 16;
                const int level = (int)(short)(c & 0xffffu);
                const int pos = (c >> 16) & 63;
                const bool raw = (c & MP2V_COEF_RAW) != 0, first_coef = (c & MP2V_COEF_FIRST) != 0;
                const int w = (&s.W[0][0])[((CF > 1 && blk >= 6) ? 128 : 0) + ((cz >> 16) & 0x40) + pos];     // luma matrices for blocks 4,5 (:184-185)
                const int mag = abs(level);
                // intra (level*W*qs)>>4, non-intra ((2*level+1)*W*qs)>>5 (:142-143); "1s" is the latter with level 1 (:84)
                int val = (((sh == 5 ? 2 * mag + 1 : mag) * w) * qs) >> sh;
                val = level < 0 ? -val : val;                                              // :144
                const int clamped = max(min((int)(short)val, 2047), -2048);                // int16 wrap, then clamp (:146)
                val = raw ? level : first_coef ? val : clamped;                            // DC as is (:160); "1s" unclamped (:84)
                const int idx2 = s.scan[pos];
                const int av = abs(val), bwv = s.bwp[pos];
                // weighted L1 norm for the saturation bound; +1: the mismatch toggle may change |F[63]| by one
                const int wsum = raw ? (av <= kMaxFirstCoef ? av * bwv : kBoundWild) : (av + 1) * bwv;
                pbit = raw ? 0u : (uint32_t)(val & 1) << slot;                             // DC is not part of the sum (:160)
                c7bit = (idx2 & 7) == 7 ? 1u << slot : 0u;
                tile0[slot * kTilePitch + idx2] = (int16_t)val;
                atomicAdd(&ws.bound[slot], wsum);
            }
            parity ^= __reduce_xor_sync(0xffffffffu, pbit);
            col7 |= __reduce_or_sync(0xffffffffu, c7bit);
        }
        // the next batch's first records (its list continues where this one ends, inside a slice): requested now, used
        // after this batch's transform and output.  (Arenas are padded: 32 records past the last one stay inside the allocation.)
        pref_idx = any_records ? base_off + (uint32_t)total + (uint32_t)lane : 0xffffffffu;
        if (any_records) pref_c = __ldg(pd.coef + pref_idx);
        __syncwarp();
        // qfs[63] ^= (sum & 1) ^ 1 (:150-152) -- only where column 7 already holds something (see the header)
        if (lane < nslots && (col7 >> lane & 1u)) tile0[lane * kTilePitch + 63] ^= (int16_t)(((parity >> lane) & 1u) ^ 1u);

        // ---- 5. inverse transform; arithmetic variant for the whole batch from the per-block bounds (a toggled-in F[63] = 1 counts too)
        const int bnd = lane < nslots ? ws.bound[lane] + (int)s.bwp[63] : 0;      // scan position 63 is tile index 63 in both scans
        const bool p1_exact = __any_sync(0xffffffffu, bnd >= kBoundWild);
        const bool p2_exact = __any_sync(0xffffffffu, bnd > kBoundLimit);
        path_counts += 1u + (p2_exact ? 1u << 8 : 0u) + (p1_exact ? 1u << 16 : 0u);      // (a warp's run is at most 60 batches)
        __syncwarp();
        // pass 1: one lane per column, in place; the transform runs across the vector index k (idct_sse2.hpp:98)
        // (no predication: the lanes of a last, partly filled trip transform whatever the tile rows behind the used
        // slots hold -- every trip stays inside the tile's kTileRows -- and nobody reads those rows)
        for (int i0 = 0; i0 < nslots * 8; i0 += 32) {
            const int i = i0 + lane;
            int16_t* p = tile0 + (i >> 3) * kTilePitch + (i & 7);
            int x[8];
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = (int)p[k * 8];
            if (p1_exact) idct_lane<0>(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
            else if (p2_exact) idct_lane<1>(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
            else idct_lane<2>(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);     // the bound that clears pass 2 also bounds every pass-1 output
#pragma unroll
            for (int k = 0; k < 8; k++) p[k * 8] = (int16_t)x[k];
        }
        __syncwarp();
        // pass 2: one lane per row of the transposed block (transpose_8x8_sse2 is the addressing); output r of
        // row k is res[r][k], shifted down by 6 (idct_sse2.hpp:100-107)
        for (int i0 = 0; i0 < nslots * 8; i0 += 32) {
            const int i = i0 + lane;
            int16_t* t = tile0 + (i >> 3) * kTilePitch;
            const int k = i & 7;
            const uint4 q = *reinterpret_cast<const uint4*>(t + k * 8);      // a quarter warp reads the 128 contiguous bytes of one block
            int x[8] = {(int)(short)(q.x & 0xffffu), (int)q.x >> 16, (int)(short)(q.y & 0xffffu), (int)q.y >> 16,
                        (int)(short)(q.z & 0xffffu), (int)q.z >> 16, (int)(short)(q.w & 0xffffu), (int)q.w >> 16};
            __syncwarp();      // the eight rows of a block sit in one trip: all are in registers before any is overwritten
            if (p2_exact) idct_lane<0>(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
            else idct_lane<2>(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
#pragma unroll
            for (int r = 0; r < 8; r++) t[r * 8 + k] = (int16_t)(x[r] >> 6);      // _mm_srai_epi16(., 6)
        }
        __syncwarp();

        // ---- 6. prediction + residual + clip + store, macroblock by macroblock
        int mbx = mbx0, mby = mby0;
        for (int mi = 0; mi < nb; mi++) {
            const uint32_t m_y = __shfl_sync(0xffffffffu, rec.y, mi), m_z = __shfl_sync(0xffffffffu, rec.z, mi), m_w = __shfl_sync(0xffffffffu, rec.w, mi);
            const int mbase = __shfl_sync(0xffffffffu, base, mi);
            const uint32_t cbp = MP2V_MB_CBP(m_y);
            const bool intra = (m_y & MP2V_MB_INTRA) != 0;
            const int buf = mi & (kWinBuf - 1);
            uint32_t pred[NT][NW];
#pragma unroll
            for (int t = 0; t < NT; t++) {
#pragma unroll
                for (int j = 0; j < NW; j++) pred[t][j] = 0;
            }
            if (!intra) {
                mbar_wait(&ws.mbar[buf], (wphase >> buf) & 1u);
                wphase ^= 1u << buf;
                const int odd8 = (mbx & 1) << 3;      // (mbx * 8) & 15 for the 8-pixel-wide chroma of 4:2:0 / 4:2:2
                bool have_pred = false;
#pragma unroll
                for (int d = 0; d < 2; d++) {
                    if (!(m_y & (d ? MP2V_MB_BWD : MP2V_MB_FWD))) continue;      // warp-uniform
                    const uint32_t mvw = d ? m_w : m_z;
                    const int mvx = (int)(short)(mvw & 0xffffu), mvy = (int)mvw >> 16;
#pragma unroll
                    for (int t = 0; t < NT; t++) {
                        // this unit's vector: luma as coded, chroma floor-halved where the format subsamples (mb_decoder.cpp:198-206)
                        const int cx = (u_chroma[t] && CF < 3) ? mvx >> 1 : mvx, cy = (u_chroma[t] && CF < 2) ? mvy >> 1 : mvy;
                        const int o = ((((u_chroma[t] && CF < 3) ? odd8 : 0) + (cx >> 1)) & 15) + u_xoff[t];
                        uint32_t q[NW];
                        pred_unit<NW>(&ws.win[buf][d][0] + u_woff[t], o, cx & 1, cy & 1, q);
                        if (have_pred) {
#pragma unroll
                            for (int j = 0; j < NW; j++) pred[t][j] = avg4(q[j], pred[t][j]);      // avg(backward, forward) (mb_decoder.cpp:240-249)
                        } else {
#pragma unroll
                            for (int j = 0; j < NW; j++) pred[t][j] = q[j];
                        }
                    }
                    have_pred = true;
                }
            }
            __syncwarp();      // every lane has read this buffer: the boxes of a later macroblock may overwrite it
            if (mi + kWinBuf < nb) {
                int ix = mbx, iy = mby;
#pragma unroll
                for (int a = 0; a < kWinBuf; a++) { if (++ix == mbw) { ix = 0; iy++; } }
                const int nx = mi + kWinBuf;
                issue_windows(__shfl_sync(0xffffffffu, rec.y, nx), __shfl_sync(0xffffffffu, rec.z, nx), __shfl_sync(0xffffffffu, rec.w, nx), ix, iy, buf);
            }
            const bool row_end = mbx + 1 == mbw;
#pragma unroll
            for (int t = 0; t < NT; t++) {
                int blk = u_blk[t], rr8 = u_rr8[t];
                uint32_t below = u_below[t];
                if (field_mbs >> mi & 1u) {                                // warp-uniform
                    blk = u_field[t] & 0xff; rr8 = u_field[t] >> 8;
                    below = (1u << blk) - 1u;
                }
                add_residual(ws.tile, mbase, cbp, below, blk, rr8, intra, pred[t][0], pred[t][1]);
                if (CF == 1) {
                    // whole rows: the right-hand block too for luma rows (the chroma lanes drop their upper two words)
                    if (lane < 16) {
                        add_residual(ws.tile, mbase, cbp, below * 2u + 1u, blk + 1, rr8, intra, pred[t][2], pred[t][3]);
                        *reinterpret_cast<uint4*>(u_dst[t]) = make_uint4(pred[t][0], pred[t][1], pred[t][2], pred[t][3]);
                    } else {
                        *reinterpret_cast<uint2*>(u_dst[t]) = make_uint2(pred[t][0], pred[t][1]);
                    }
                } else {
                    *reinterpret_cast<uint2*>(u_dst[t]) = make_uint2(pred[t][0], pred[t][1]);
                }
                u_dst[t] += row_end ? u_wrap[t] : u_adv[t];      // next macroblock: one to the right, or the first of the next macroblock row
            }
            if (row_end) { mbx = 0; mby++; } else mbx++;
            __syncwarp();
        }
        first += nb;
        mbx0 = mbx; mby0 = mby;
    }
    if (lane == 0 && batch.counters) {
        atomicAdd(&batch.counters[0], (unsigned long long)(path_counts & 0xffu));
        if (path_counts >> 8) {
            atomicAdd(&batch.counters[1], (unsigned long long)((path_counts >> 8) & 0xffu));
            atomicAdd(&batch.counters[2], (unsigned long long)(path_counts >> 16));
        }
    }
}

template <int CF>
static cudaError_t prepare_kernel3() {
    return cudaFuncSetAttribute(recon_kernel3<CF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(cta_smem_t<CF>) + 128));
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
    const cuuint32_t cbox_h = (chroma_format == 1 ? 8 : 16) + 1;      // geo_t: every box is kBoxW = 32 bytes wide
    for (int p = 0; p < 3; p++) {
        const cuuint64_t gdim[3] = {(cuuint64_t)lay.width[p], (cuuint64_t)lay.height[p], (cuuint64_t)n_frames};
        const cuuint64_t gstride[2] = {(cuuint64_t)lay.stride[p], (cuuint64_t)frame_alloc};      // bytes, dimensions 1 and 2
        const cuuint32_t box[3] = {(cuuint32_t)kBoxW, p == 0 ? 17u : cbox_h, 1u};
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
    switch (chroma_format) {
        case 1: recon_kernel3<1><<<grid, block, sizeof(cta_smem_t<1>) + 128, stream>>>(batch, tm); break;
        case 2: recon_kernel3<2><<<grid, block, sizeof(cta_smem_t<2>) + 128, stream>>>(batch, tm); break;
        case 3: recon_kernel3<3><<<grid, block, sizeof(cta_smem_t<3>) + 128, stream>>>(batch, tm); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t recon_kernel_attributes(int chroma_format, cudaFuncAttributes* out) {
    switch (chroma_format) {
        case 1: return cudaFuncGetAttributes(out, recon_kernel3<1>);
        case 2: return cudaFuncGetAttributes(out, recon_kernel3<2>);
        case 3: return cudaFuncGetAttributes(out, recon_kernel3<3>);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace mp2v
