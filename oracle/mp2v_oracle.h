/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's reconstruction path
 * (fxslava/tiny_mp2v_dec, x86-64 / SSE2 variant) -- the checker for the CUDA kernels.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the product
 * (libmp2v_b200.so) never links or calls it.
 *
 * Parity pin: tests/test_oracle_vs_reference.py decodes in-repo generated streams with the real
 * reference (oracle/_ref, built from /root/reference by oracle/Makefile) and requires this
 * restatement, fed the generator's ground-truth macroblock records, to produce the same YUV
 * bit for bit; tests/golden/ holds hashes of those reference outputs for boxes without the reference.
 */
#ifndef MP2V_ORACLE_H
#define MP2V_ORACLE_H
#include <stdint.h>
#include "mp2v_recon.h"

#ifdef __cplusplus
extern "C" {
#endif

/* scan tables: scan_trans[alt][i] = transposed raster index of scan position i (scan_c.cpp:4-21) */
void orc_scan_tables(uint8_t scan_trans[2][64], uint8_t shuffle[2][64], uint8_t scan[2][64]);

/* quantiser_matrices of mp2v_picture_c::init (decoder.cpp:154-192): tx[k] = matrix k as transmitted
 * (zig-zag order), W[k][i] = tx[k][ scan0[ shuffle[alt][i] ] ] */
void orc_build_W(const uint8_t tx[4][64], int alternate_scan, uint8_t W[4][64]);

/* parse_block's dequantisation + mismatch (mb_decoder.cpp:74-163) for ONE block from its records */
void orc_dequant_block(const mp2v_coef_t* coef, int n, int intra, const uint8_t W[64], int qscale,
                       int alternate_scan, int16_t F[64]);

/* inverse_dct_template (idct_sse2.hpp:23-120) up to and including the >>6: res[r*8+c] */
void orc_idct_sse2(const int16_t F[64], int16_t res[64]);

/* one prediction fetch, mc_c.hpp:3-17: w x h pixels from ref at (x0,y0) with half-pel flags */
void orc_mc_fetch(const uint8_t* ref, int stride, int x0, int y0, int hx, int hy, int w, int h, uint8_t* out);

/* whole picture: params + macroblock records + coefficient records -> dst planes.
 * planes use mp2v_frame_layout (frame_c rule, decoder.cpp:44-66); l0/l1 may be NULL. */
int orc_recon_picture(const mp2v_pic_params_t* params, const mp2v_mb_info_t* mb, const mp2v_coef_t* coef,
                      int width, int height, int chroma_format,
                      uint8_t* const dst[3], const uint8_t* const l0[3], const uint8_t* const l1[3]);

#ifdef __cplusplus
}
#endif
#endif
