// sm_100a reconstruction kernel: inverse quantisation + mismatch control + the reference's exact
// 16-bit saturating IDCT + half-pel forward / backward / bidirectional motion compensation +
// residual add + clip, fused per macroblock so that a prediction is never written to HBM and read
// back (the CPU reference writes it with mc_pred*/mc_bidir* and re-reads it in
// inverse_dct_template<true>, idct_sse2.hpp:110-114).
//
// One CTA (128 threads) owns a group of consecutive macroblocks of one picture and walks four phases
// separated by __syncthreads():
//   0  stage W[4][64], the scan table and the group's macroblock records in shared memory;
//      exclusive scan of popc(cbp) assigns every coded block a 144-byte slot of the residual tile
//   1  zero the used slots
//   2  dequantise: a warp streams one macroblock's coefficient records with coalesced 32-bit loads,
//      applies parse_block's arithmetic (mb_decoder.cpp:139-146) and scatters int16 values to
//      tile[slot][g_scan_trans[pos]]; the mismatch parity of all (<= 12) blocks of the macroblock is
//      one REDUX.XOR per 32 records
//   3  IDCT: one thread per coded block, all 64 values in registers, so the 8x8 transpose between
//      the two passes (transpose_8x8_sse2, idct_sse2.hpp:67-94) is register renaming; every SSE2
//      lane operation is reproduced on sign-extended int32: adds/subs = VIADDMNMX+VIMNMX,
//      mulhi = one IMAD.HI against (c << 16), slli = SHF + sign-extending PRMT
//   4  prediction: a warp takes one macroblock at a time, copies the (w+1)x(h+1) reference windows
//      of all planes / directions into shared memory as aligned 16-byte chunks (one L1 wavefront
//      per window row), then every lane produces one output row: funnel-shift realignment,
//      __vavgu4 rounding averages in the reference's order (mc_c.hpp:3-17), residual add and
//      unsigned saturation as VIADDMNMX.S16x2.RELU, one 128-bit (or 64-bit) store.
//
// The roofline that bounds this kernel and the byte counts are in DESIGN.md.
#include "recon_kernels.cuh"

#include "host/scan_tables.h"

namespace mp2v {

__constant__ uint8_t c_scan_trans[2][64];

template <int CF>
struct fmt_t {
    static constexpr int NBLK = CF == 1 ? 6 : CF == 2 ? 8 : 12;
    static constexpr int MBG = mbs_per_cta(CF);
    static constexpr int NSLOT = MBG * NBLK;
    static constexpr int CW = CF == 3 ? 16 : 8;     // chroma macroblock width
    static constexpr int CH = CF == 1 ? 8 : 16;     // chroma macroblock height
    static constexpr int WIN_LUMA = 17 * 32;        // 17 rows x 2 aligned 16-byte chunks
    static constexpr int WIN_CHROMA = (CH + 1) * 32;
    static constexpr int WIN_DIR = WIN_LUMA + 2 * WIN_CHROMA;
    static constexpr int N_ITEMS = 2 * 17 + 2 * 2 * (CH + 1);   // 16-byte chunks per direction
    static constexpr int N_UNITS = 16 + 2 * CH;                 // output rows per macroblock
};

constexpr int kTilePitch = 72;   // int16 per slot: 64 + 8 pad -> 144 B, conflict-free 128-bit row access

template <int CF>
struct smem_t {
    alignas(16) int16_t tile[fmt_t<CF>::NSLOT][kTilePitch];
    alignas(16) uint8_t win[kCtaThreads / 32][2][fmt_t<CF>::WIN_DIR];
    alignas(16) uint4 mb[fmt_t<CF>::MBG];
    alignas(16) uint8_t W[4][64];
    alignas(16) uint8_t scan[64];
    int prefix[fmt_t<CF>::MBG + 1];
};

// ------------------------------------------------------------------------------------------------
// The reference's 16-bit lane arithmetic on sign-extended int32 (idct_sse2.hpp:7-21 helpers).
// Written as inline PTX on purpose: given C++ min/max on values the compiler knows to be
// sign-extended int16, LLVM canonicalises the clamp into sadd.sat.i16 and the NVPTX back end then
// legalises that into ~10 instructions (PRMT sign extensions + ISETP/LOP3 overflow logic).  ptxas
// turns the PTX below into VIADDMNMX + VIMNMX (2 ALU ops), IMAD + SHF, and SHF/LEA pairs.
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

// one lane of idct_1d_sse2 (idct_sse2.hpp:23-65), op for op
__device__ __forceinline__ void idct_lane(int& x0, int& x1, int& x2, int& x3, int& x4, int& x5, int& x6, int& x7) {
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
}

// inverse_dct_template up to the >>6 (idct_sse2.hpp:96-107), in place on one tile slot:
// in  F[k*8+c] (the transposed-raster layout parse_block writes), out res[r*8+c]
__device__ __forceinline__ void idct_block(int16_t* slot) {
    int v[64];
    uint4* q = reinterpret_cast<uint4*>(slot);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint4 w = q[k];
        v[k * 8 + 0] = (int)(short)(w.x & 0xffff); v[k * 8 + 1] = (int)w.x >> 16;
        v[k * 8 + 2] = (int)(short)(w.y & 0xffff); v[k * 8 + 3] = (int)w.y >> 16;
        v[k * 8 + 4] = (int)(short)(w.z & 0xffff); v[k * 8 + 5] = (int)w.z >> 16;
        v[k * 8 + 6] = (int)(short)(w.w & 0xffff); v[k * 8 + 7] = (int)w.w >> 16;
    }
    // pass 1: the transform runs across the vector index k for every lane c
#pragma unroll
    for (int c = 0; c < 8; c++)
        idct_lane(v[c], v[8 + c], v[16 + c], v[24 + c], v[32 + c], v[40 + c], v[48 + c], v[56 + c]);
    // transpose + pass 2: lane c of the transposed block is row c; results land in v[c*8 + r] = res[r][c]
#pragma unroll
    for (int c = 0; c < 8; c++)
        idct_lane(v[c * 8 + 0], v[c * 8 + 1], v[c * 8 + 2], v[c * 8 + 3], v[c * 8 + 4], v[c * 8 + 5], v[c * 8 + 6], v[c * 8 + 7]);
#pragma unroll
    for (int r = 0; r < 8; r++) {
        uint4 w;
        w.x = __byte_perm(v[0 * 8 + r] >> 6, v[1 * 8 + r] >> 6, 0x5410);
        w.y = __byte_perm(v[2 * 8 + r] >> 6, v[3 * 8 + r] >> 6, 0x5410);
        w.z = __byte_perm(v[4 * 8 + r] >> 6, v[5 * 8 + r] >> 6, 0x5410);
        w.w = __byte_perm(v[6 * 8 + r] >> 6, v[7 * 8 + r] >> 6, 0x5410);
        q[r] = w;
    }
}

// ------------------------------------------------------------------------------------------------
// prediction row: NW words (4 pixels each) of plane row `row` of a staged window, realigned from
// byte offset o, with the reference's half-pel averaging order (mc_c.hpp:3-17)
template <int NW>
__device__ __forceinline__ void pred_row(const uint8_t* win_row, int o, int hx, int hy, uint32_t (&out)[NW]) {
    const int sh = (o & 3) * 8;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(win_row) + (o >> 2);
    uint32_t w[NW + 1];
#pragma unroll
    for (int j = 0; j <= NW; j++) w[j] = p[j];
#pragma unroll
    for (int j = 0; j < NW; j++) out[j] = __funnelshift_rc(w[j], w[j + 1], sh);
    if (hx) {
#pragma unroll
        for (int j = 0; j < NW; j++) out[j] = __vavgu4(out[j], __funnelshift_rc(w[j], w[j + 1], sh + 8));
    }
    if (hy) {
        uint32_t b[NW];
#pragma unroll
        for (int j = 0; j <= NW; j++) w[j] = p[j + 8];   // next window row (32-byte pitch)
#pragma unroll
        for (int j = 0; j < NW; j++) b[j] = __funnelshift_rc(w[j], w[j + 1], sh);
        if (hx) {
#pragma unroll
            for (int j = 0; j < NW; j++) b[j] = __vavgu4(b[j], __funnelshift_rc(w[j], w[j + 1], sh + 8));
        }
#pragma unroll
        for (int j = 0; j < NW; j++) out[j] = __vavgu4(out[j], b[j]);
    }
}

// pred (4 pixels) + residual (two int16x2 words), unsigned-saturated: packus(adds_epi16(zext(dst), res))
__device__ __forceinline__ uint32_t add_clip4(uint32_t pred, uint32_t r01, uint32_t r23) {
    const uint32_t lo = __vimin_s16x2_relu(__vadd2(__byte_perm(pred, 0, 0x4140), r01), 0x00ff00ffu);
    const uint32_t hi = __vimin_s16x2_relu(__vadd2(__byte_perm(pred, 0, 0x4342), r23), 0x00ff00ffu);
    return __byte_perm(lo, hi, 0x6420);
}

__device__ __forceinline__ int chroma_mv(int mv, bool halve) { return halve ? (mv >> 1) : mv; }   // floor, mb_decoder.cpp:198-206

template <int CF>
__global__ void __launch_bounds__(kCtaThreads, 4) recon_kernel(const __grid_constant__ batch_desc_t batch) {
    using F = fmt_t<CF>;
    __shared__ smem_t<CF> s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pi = blockIdx.x / batch.ctas_per_pic;
    const int grp = blockIdx.x - pi * batch.ctas_per_pic;
    const pic_desc_t& pd = batch.pic[pi];
    const int mb0 = grp * F::MBG;
    const int nmb = min(F::MBG, batch.mb_count - mb0);

    // ---- phase 0: stage tables and macroblock records
    if (tid < 64) {
        reinterpret_cast<uint32_t*>(&s.W[0][0])[tid] = reinterpret_cast<const uint32_t*>(&pd.params->W[0][0])[tid];
    } else if (tid < 80) {
        const int alt = pd.params->alternate_scan ? 1 : 0;
        reinterpret_cast<uint32_t*>(s.scan)[tid - 64] = reinterpret_cast<const uint32_t*>(c_scan_trans[alt])[tid - 64];
    } else if (tid >= 96 && tid < 96 + F::MBG) {
        const int i = tid - 96;
        s.mb[i] = i < nmb ? reinterpret_cast<const uint4*>(pd.mb)[mb0 + i] : make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    if (warp == 0) {
        int c = lane < F::MBG ? __popc(MP2V_MB_CBP(s.mb[lane < F::MBG ? lane : 0].y)) : 0;
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane < F::MBG) s.prefix[lane + 1] = incl;
        if (lane == 0) s.prefix[0] = 0;
    }
    __syncthreads();
    const int nslots = s.prefix[nmb];

    // ---- phase 1: zero the used slots (QFS[64] = {0}, mb_decoder.cpp:159)
    for (int i = tid; i < nslots * 8; i += kCtaThreads)
        reinterpret_cast<uint4*>(&s.tile[i >> 3][0])[i & 7] = make_uint4(0, 0, 0, 0);
    __syncthreads();

    // ---- phase 2: dequantise + saturate + mismatch (mb_decoder.cpp:74-155)
    for (int i = warp; i < nmb; i += kCtaThreads / 32) {
        const uint4 m = s.mb[i];
        const int n = MP2V_MB_NCOEF(m.y);
        const int qs = MP2V_MB_QSCALE(m.y);
        const uint32_t cbp = MP2V_MB_CBP(m.y);
        const bool intra = (m.y & MP2V_MB_INTRA) != 0;
        const int base = s.prefix[i];
        const mp2v_coef_t* cp = pd.coef + m.x;
        uint32_t parity = 0;
        for (int k0 = 0; k0 < n; k0 += 32) {
            const int k = k0 + lane;
            uint32_t pbit = 0;
            const uint32_t c = k < n ? cp[k] : 0u;
            if (k < n && (cbp >> ((c >> 22) & 15) & 1)) {   // a record naming an uncoded block is ignored (memory safety)
                const int level = (int)(short)(c & 0xffffu);
                const int pos = (c >> 16) & 63, blk = (c >> 22) & 15;
                const int slot = base + __popc(cbp & ((1u << blk) - 1u));
                int val, idx;
                if (c & MP2V_COEF_RAW) { val = level; idx = 0; }                        // intra DC, not summed (:160)
                else {
                    const int w = s.W[(blk < 6 ? 0 : 2) + (intra ? 0 : 1)][pos];        // luma matrices for blocks 4,5 (:184-185)
                    const int mag = abs(level);
                    if (c & MP2V_COEF_FIRST) { val = (3 * w * qs) >> 5; idx = 0; }      // first coefficient "1s": no clamp (:84)
                    else {
                        val = intra ? (mag * w * qs) >> 4 : ((2 * mag + 1) * w * qs) >> 5;   // :142-143
                        idx = s.scan[pos];
                    }
                    if (level < 0) val = -val;                                          // :144
                    if (!(c & MP2V_COEF_FIRST)) val = max(min((int)(short)val, 2047), -2048);   // int16 wrap, then clamp (:146)
                    pbit = (uint32_t)(val & 1) << blk;
                }
                s.tile[slot][idx] = (int16_t)val;
            }
            parity ^= __reduce_xor_sync(0xffffffffu, pbit);
        }
        __syncwarp();
        // qfs[63] ^= (sum & 1) ^ 1 for every coded block (:150-152)
        if (lane < F::NBLK && (cbp >> lane & 1)) {
            const int slot = base + __popc(cbp & ((1u << lane) - 1u));
            s.tile[slot][63] ^= (int16_t)(((parity >> lane) & 1u) ^ 1u);
        }
    }
    __syncthreads();

    // ---- phase 3: IDCT, one thread per coded block
    for (int slot = tid; slot < nslots; slot += kCtaThreads) idct_block(&s.tile[slot][0]);
    __syncthreads();

    // ---- phase 4: prediction + residual + clip + store, one macroblock per warp at a time
    const int mbw = batch.mbw;
    for (int i = warp; i < nmb; i += kCtaThreads / 32) {
        const uint4 m = s.mb[i];
        const int mbi = mb0 + i;
        const int mby = mbi / mbw, mbx = mbi - mby * mbw;
        const uint32_t cbp = MP2V_MB_CBP(m.y);
        const bool fwd = (m.y & MP2V_MB_FWD) != 0, bwd = (m.y & MP2V_MB_BWD) != 0;
        const int base = s.prefix[i];
        uint8_t* win = &s.win[warp][0][0];

        // stage the reference windows: aligned 16-byte chunks, 2 per row, (h+1) rows per plane
#pragma unroll
        for (int d = 0; d < 2; d++) {
            if (!(d ? bwd : fwd)) continue;
            const uint32_t mvw = d ? m.w : m.z;
            const int mvx = (int)(short)(mvw & 0xffffu), mvy = (int)mvw >> 16;
            const uint8_t* const* ref = d ? pd.l1 : pd.l0;
            for (int j = lane; j < F::N_ITEMS; j += 32) {
                int p, jj;
                if (j < 34) { p = 0; jj = j; }
                else if (j < 34 + 2 * (F::CH + 1)) { p = 1; jj = j - 34; }
                else { p = 2; jj = j - 34 - 2 * (F::CH + 1); }
                const int r = jj >> 1, ch = jj & 1;
                const int cx = p ? chroma_mv(mvx, CF < 3) : mvx, cy = p ? chroma_mv(mvy, CF < 2) : mvy;
                const int pw = p ? F::CW : 16, ph = p ? F::CH : 16;
                const int x0 = mbx * pw + (cx >> 1), y0 = mby * ph + (cy >> 1);
                const uint8_t* src = ref[p] + (size_t)(y0 + r) * batch.stride[p] + (x0 & ~15) + 16 * ch;
                const int woff = d * F::WIN_DIR + (p == 0 ? 0 : F::WIN_LUMA + (p - 1) * F::WIN_CHROMA) + r * 32 + 16 * ch;
                *reinterpret_cast<uint4*>(win + woff) = __ldg(reinterpret_cast<const uint4*>(src));
            }
        }
        __syncwarp();

        for (int u = lane; u < F::N_UNITS; u += 32) {
            int p, r;
            if (u < 16) { p = 0; r = u; }
            else if (u < 16 + F::CH) { p = 1; r = u - 16; }
            else { p = 2; r = u - 16 - F::CH; }
            const bool wide = (p == 0) || (CF == 3);
            const int pw = p ? F::CW : 16, ph = p ? F::CH : 16;
            uint32_t pred[4] = {0, 0, 0, 0};
            bool have = false;
#pragma unroll
            for (int d = 0; d < 2; d++) {
                if (!(d ? bwd : fwd)) continue;
                const uint32_t mvw = d ? m.w : m.z;
                const int mvx = (int)(short)(mvw & 0xffffu), mvy = (int)mvw >> 16;
                const int cx = p ? chroma_mv(mvx, CF < 3) : mvx, cy = p ? chroma_mv(mvy, CF < 2) : mvy;
                const int o = (mbx * pw + (cx >> 1)) & 15;
                const uint8_t* row = win + d * F::WIN_DIR + (p == 0 ? 0 : F::WIN_LUMA + (p - 1) * F::WIN_CHROMA) + r * 32;
                uint32_t q[4] = {0, 0, 0, 0};
                if (wide) pred_row<4>(row, o, cx & 1, cy & 1, q);
                else { uint32_t q2[2]; pred_row<2>(row, o, cx & 1, cy & 1, q2); q[0] = q2[0]; q[1] = q2[1]; }
                if (have) {
#pragma unroll
                    for (int j = 0; j < 4; j++) pred[j] = __vavgu4(q[j], pred[j]);   // bidirectional rounding average
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) pred[j] = q[j];
                }
                have = true;
            }
            // residual blocks covering this row (block geometry: mb_decoder.cpp:177-195)
            int bl, br = -1;
            if (p == 0) { bl = (r >> 3) * 2; br = bl + 1; }
            else if (CF == 1) bl = 3 + p;
            else if (CF == 2) bl = 3 + p + ((r >> 3) << 1);
            else { bl = 3 + p + ((r >> 3) << 1); br = bl + 4; }
            const int rr = r & 7;
            uint32_t out[4];
            {
                uint4 res = make_uint4(0, 0, 0, 0);
                if (cbp >> bl & 1) res = *reinterpret_cast<const uint4*>(&s.tile[base + __popc(cbp & ((1u << bl) - 1u))][rr * 8]);
                out[0] = add_clip4(pred[0], res.x, res.y);
                out[1] = add_clip4(pred[1], res.z, res.w);
            }
            uint8_t* drow = pd.dst[p] + (size_t)(mby * ph + r) * batch.stride[p] + mbx * pw;
            if (wide) {
                uint4 res = make_uint4(0, 0, 0, 0);
                if (cbp >> br & 1) res = *reinterpret_cast<const uint4*>(&s.tile[base + __popc(cbp & ((1u << br) - 1u))][rr * 8]);
                out[2] = add_clip4(pred[2], res.x, res.y);
                out[3] = add_clip4(pred[3], res.z, res.w);
                *reinterpret_cast<uint4*>(drow) = make_uint4(out[0], out[1], out[2], out[3]);
            } else {
                *reinterpret_cast<uint2*>(drow) = make_uint2(out[0], out[1]);
            }
        }
        __syncwarp();
    }
}

// __constant__ symbols are per device: remember which devices have been initialised
static cudaError_t ensure_tables() {
    static bool done[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!done[dev]) {
        const scan_tables_t& t = scan_tables();
        e = cudaMemcpyToSymbol(c_scan_trans, t.scan_trans, sizeof(t.scan_trans));
        if (e != cudaSuccess) return e;
        done[dev] = true;
    }
    return cudaSuccess;
}

cudaError_t launch_recon(int chroma_format, const batch_desc_t& batch, cudaStream_t stream) {
    cudaError_t e = ensure_tables();
    if (e != cudaSuccess) return e;
    if (batch.n_pics < 1 || batch.n_pics > kMaxBatch) return cudaErrorInvalidValue;
    const dim3 grid((unsigned)(batch.n_pics * batch.ctas_per_pic)), block(kCtaThreads);
    switch (chroma_format) {
        case 1: recon_kernel<1><<<grid, block, 0, stream>>>(batch); break;
        case 2: recon_kernel<2><<<grid, block, 0, stream>>>(batch); break;
        case 3: recon_kernel<3><<<grid, block, 0, stream>>>(batch); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t recon_kernel_attributes(int chroma_format, cudaFuncAttributes* out) {
    switch (chroma_format) {
        case 1: return cudaFuncGetAttributes(out, recon_kernel<1>);
        case 2: return cudaFuncGetAttributes(out, recon_kernel<2>);
        case 3: return cudaFuncGetAttributes(out, recon_kernel<3>);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace mp2v
