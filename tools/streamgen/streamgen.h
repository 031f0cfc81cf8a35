/*
 * Synthetic MPEG-2 video elementary stream generator (test + benchmark infrastructure).
 *
 * There is no network and the reference ships no media, so every stream this
 * repository decodes is generated here, seeded and deterministic.  The generator writes syntax the
 * reference decoder accepts (SURVEY.md 8(c) "reference envelope": progressive frame pictures,
 * frame_pred_frame_dct=1, intra_vlc_format=1, a quant_matrix_extension in every picture, one slice
 * per macroblock row, closed GOPs, motion vectors inside the frame) and, next to the bitstream,
 * the GROUND-TRUTH macroblock / coefficient records (include/mp2v_recon.h) it encoded -- which lets
 * tests check the host slice parser and the CPU oracle independently of each other.
 */
#ifndef MP2V_STREAMGEN_H
#define MP2V_STREAMGEN_H
#include <stddef.h>
#include <stdint.h>
#include "mp2v_recon.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mp2v_gen_params {
    int32_t  width, height;        /* coded size, multiples of 16                                   */
    int32_t  chroma_format;        /* 1, 2, 3                                                        */
    int32_t  n_gops;               /* closed GOPs; sequence header + extension repeated before each */
    int32_t  gop_n;                /* pictures per GOP (display)                                     */
    int32_t  gop_m;                /* distance between references; 1 = no B pictures               */
    int32_t  intra_only;           /* every picture is an I picture                                  */
    uint64_t seed;
    int32_t  mode;                 /* 0 fuzz (random syntax), 1 statistical (decaying run/level spectra), 2 texture: a translating
                                      procedural texture + noise, really encoded (forward DCT, quantiser_scale 4..8, global motion) */
    int32_t  mv_range;             /* max |integer luma displacement| in pixels                    */
    int32_t  qscale_code_max;      /* quantiser_scale_code drawn from 1..this (<= 31)              */
    int32_t  alternate_scan;       /* 0 / 1 fixed, -1 random per picture                            */
    int32_t  q_scale_type;         /* 0 / 1 fixed, -1 random per picture                            */
    int32_t  intra_dc_precision;   /* 0..3 fixed, -1 random per picture                             */
    int32_t  pct_skipped;          /* % of P/B macroblocks skipped (when legal)                    */
    int32_t  pct_intra_in_pb;      /* % intra macroblocks inside P/B pictures                      */
    int32_t  pct_coded;            /* % of non-intra macroblocks that carry a coded_block_pattern  */
    int32_t  pct_mb_quant;         /* % of coded macroblocks that change quantiser_scale           */
    int32_t  pct_big_levels;       /* % of coefficients drawn from the full +-2047 escape range    */
    int32_t  all_blocks_coded;     /* cbp = all ones whenever pattern is present                   */
    int32_t  natural_mean_coefs;   /* mode 1: mean number of AC coefficients per coded block (0 = 2.6) */
    int32_t  unclamped_mv;         /* 1: vectors may leave the frame (invalid streams for error-path tests) */
    int32_t  user_data_bytes;      /* > 0: a user_data() of this many bytes follows every sequence_extension (the reference collects them, decoder.cpp:194-200) */
    int32_t  texture_noise;        /* mode 2: per-frame noise amplitude (+-n grey levels, 0 = 3): sets the bit rate */
    int32_t  matrices_once;        /* 1: quant_matrix_extension only in the first picture of every GOP; later pictures keep those matrices
                                      (ISO/IEC 13818-2 6.3.11).  OUTSIDE the reference's envelope: it needs the extension in every picture */
    int32_t  pct_field_dct;        /* > 0: pictures are coded with frame_pred_frame_dct = 0 (interlaced frame pictures, frame-based
                                      prediction only: frame_motion_type = 2) and this % of the intra / pattern macroblocks use
                                      dct_type = 1 (field DCT, mb_decoder.cpp:172-195, 357-360)                                */
    int32_t  intra_vlc_table0;     /* 1: intra_vlc_format = 0 -- intra blocks code their AC coefficients with table B.14 like non-intra
                                      blocks (default: intra_vlc_format = 1, table B.15)                                        */
} mp2v_gen_params_t;

typedef struct mp2v_gen mp2v_gen_t;

/* fills p with the defaults of SURVEY.md 8(d) fuzz mode for the given geometry */
MP2V_API void mp2v_gen_default_params(mp2v_gen_params_t* p, int width, int height, int chroma_format);
MP2V_API mp2v_gen_t* mp2v_gen_create(const mp2v_gen_params_t* p);
MP2V_API void mp2v_gen_destroy(mp2v_gen_t* g);
MP2V_API const char* mp2v_gen_error(mp2v_gen_t* g);

/* the elementary stream; the buffer is followed by >= 256 zero bytes (decoder.h convention) */
MP2V_API size_t mp2v_gen_stream(mp2v_gen_t* g, const uint8_t** data);
/* byte offset where GOP i's sequence header starts (i == n_gops: end of the last GOP's pictures) */
MP2V_API size_t mp2v_gen_gop_offset(mp2v_gen_t* g, int gop);

typedef struct mp2v_gen_picture {
    mp2v_pic_params_t params;      /* dst/l0/l1 hold CODED-ORDER picture indices of the refs (-1 none) */
    const mp2v_mb_info_t* mb;
    const mp2v_coef_t* coef;
    uint32_t mb_count;
    uint32_t n_coef;
    int32_t  display_index;        /* position in display order over the whole stream               */
    int32_t  gop;
    int32_t  q_scale_type;
    int32_t  intra_dc_precision;
    uint8_t  tx[4][64];            /* matrices as transmitted (zig-zag order)                       */
    int32_t  tx_loaded[4];
} mp2v_gen_picture_t;

MP2V_API int mp2v_gen_num_pictures(mp2v_gen_t* g);
MP2V_API int mp2v_gen_picture(mp2v_gen_t* g, int coded_index, mp2v_gen_picture_t* out);

#ifdef __cplusplus
}
#endif
#endif
