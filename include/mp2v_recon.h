/*
 * mp2v_recon.h -- C ABI of the B200 MPEG-2 reconstruction back end (libmp2v_b200.so).
 *
 * This is the batched replacement of the reference decoder's per-block / per-macroblock
 * reconstruction seam (fxslava/tiny_mp2v_dec):
 *
 *   reference call site (synchronous, on a CPU worker thread)           replaced by
 *   ------------------------------------------------------------------  --------------------------------
 *   parse_block<>: dequant + saturate + mismatch  mb_decoder.cpp:74-155   mp2v_coef_t stream  -> kernel
 *   inverse_dct_template<add>(plane,F,stride)     idct_sse2.hpp:96-120    fused in the same kernel
 *   mc_pred_16xh/8xh[4], mc_bidir_16xh/8xh[16]    mc.h:6-12, mc.cpp:4-25  mp2v_mb_info_t      -> kernel
 *   base_motion_compensation<>                    mb_decoder.cpp:291-339  (per-MB flags + vectors)
 *   mp2v_picture_c::init() quantiser_matrices     decoder.cpp:154-192     mp2v_pic_params_t.W
 *   frame_c planes / strides / pool               decoder.cpp:44-105      device frame pool (frame ids)
 *   picture dependencies + display hand-off       threads.cpp, decoder.cpp:346-379   submit order + map_frame
 *   decode_slice -> parse_macroblock / parse_block decoder.cpp:107-152,              mp2v_recon_submit_slices:
 *     (VLC, DC / PMV prediction, skipped MBs)      mb_decoder.cpp:74-155, 521-641     slice-parallel parser kernel
 *   scan_start_codes over decode()'s buffer       start_codes_search.hpp:7-26,       mp2v_recon_stream_begin / _codes / _add:
 *                                                 decoder.cpp:283-288                scan kernels over the device-resident stream
 *   decode()'s slice start-code switch            decoder.cpp:318-326                mp2v_recon_submit_stream_picture (slice offsets)
 *   sample's planar YUV write from host memory    tiny_mp2v_dec.cpp:11-17            frame_device_ptrs / convert_frames (NV12, P010, UYVY)
 *
 * Three ways in.  (1) A host slice parser (VLC, DC prediction, motion-vector prediction, skipped-
 * macroblock resolution) fills one mp2v_picture_t per coded picture in pinned memory and
 * mp2v_recon_submit() copies it to the device.  With MP2V_RECON_DEVICE_VLC a kernel produces the same
 * records on the device: (2) the caller hands over one picture's coded slices (mp2v_recon_submit_slices,
 * staged and copied per picture), or (3) the whole elementary stream once (mp2v_recon_stream_begin) and
 * then pictures as slice offsets into that resident copy (mp2v_recon_submit_stream_picture).  Either way
 * the whole picture is reconstructed (batched with any other submitted pictures that do not depend
 * on each other) by hand-written sm_100a kernels.  There is no CPU fallback: every entry point fails
 * with MP2V_ERR_CUDA when no device / kernel image is usable.
 *
 * Plain C: pointers and sizes only, no C++ or torch types.  Thread-safety: one context may be used
 * from several threads for acquire / fill (disjoint pictures); submit, map and sync are serialised
 * internally.
 */
#ifndef MP2V_RECON_H
#define MP2V_RECON_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MP2V_API __attribute__((visibility("default")))
#else
#define MP2V_API
#endif

/* ---- status codes --------------------------------------------------------------------------- */
enum {
    MP2V_OK = 0,
    MP2V_ERR_ARG = -1,      /* bad argument / geometry / id                                    */
    MP2V_ERR_CUDA = -2,     /* CUDA runtime or launch failure (no device, no sm_100a image...)  */
    MP2V_ERR_NOMEM = -3,
    MP2V_ERR_STATE = -4,    /* call order violated (e.g. reference frame never written)          */
    MP2V_ERR_RANGE = -5     /* picture data fails validation (motion vector outside the frame,
                               coefficient index past the arena, ...)                           */
};

/* ---- per-macroblock record (16 bytes) --------------------------------------------------------
 * One per macroblock in raster order (skipped macroblocks included, resolved by the host to a
 * prediction-only record, mb_decoder.cpp:541-550).
 *   coef_off : [30:0] index of this macroblock's first mp2v_coef_t in the picture's coefficient arena
 *              [31]   field_dct  dct_type = 1 (frame pictures with frame_pred_frame_dct = 0): the luma blocks
 *                                hold the two fields -- blocks 0,1 the even rows of the macroblock, 2,3 the
 *                                odd rows -- and so do the chroma blocks of 4:2:2 / 4:4:4
 *                                (mb_decoder.cpp:172-195; 4:2:0 chroma is always frame organised)
 *   bits     : [ 9: 0] n_coef   number of coefficient records (<= 12*64)
 *              [16:10] qscale   quantiser_scale 1..112 after q_scale_type mapping
 *                               (decoder.cpp:140-145, mb_decoder.cpp:555-563)
 *              [28:17] cbp      bit i = block i coded; block order Y0 Y1 Y2 Y3 Cb Cr, then
 *                               Cb' Cr' (4:2:2 lower halves), then 4:4:4 right halves
 *                               (mb_decoder.cpp:177-195)
 *              [29]    intra    blocks overwrite (add=false); no prediction
 *              [30]    fwd      prediction from L0 with mv[0]
 *              [31]    bwd      prediction from L1 with mv[1]; fwd|bwd = rounding average
 *              non-intra with neither fwd nor bwd is not emitted: the host resolves P "no-MC"
 *              and P-skipped macroblocks to fwd with a zero vector (mb_decoder.cpp:329-338).
 *   mv[s][t] : luma half-pel units, s = 0 forward / 1 backward, t = 0 x / 1 y
 */
typedef struct mp2v_mb_info {
    uint32_t coef_off;
    uint32_t bits;
    int16_t  mv[2][2];
} mp2v_mb_info_t;

#define MP2V_MB_FIELD_DCT   (1u << 31)                       /* in coef_off */
#define MP2V_MB_COEF_OFF(o) ((o) & 0x7fffffffu)
#define MP2V_MB_NCOEF(b)   ((b) & 0x3ffu)
#define MP2V_MB_QSCALE(b)  (((b) >> 10) & 0x7fu)
#define MP2V_MB_CBP(b)     (((b) >> 17) & 0xfffu)
#define MP2V_MB_INTRA      (1u << 29)
#define MP2V_MB_FWD        (1u << 30)
#define MP2V_MB_BWD        (1u << 31)
#define MP2V_MB_BITS(ncoef, qscale, cbp, flags) \
    (((uint32_t)(ncoef) & 0x3ffu) | (((uint32_t)(qscale) & 0x7fu) << 10) | (((uint32_t)(cbp) & 0xfffu) << 17) | (flags))

/* ---- coefficient record (4 bytes) ------------------------------------------------------------
 * Run/level-decoded, NOT dequantised ("ship levels, not products").  Records of one macroblock are
 * contiguous, block by block in coding order, each block's records in scan order.
 *   [15: 0] level   signed; for MP2V_COEF_RAW the final intra DC value
 *                   wrap16(dc_pred << (3 - intra_dc_precision)) (mb_decoder.cpp:46-72)
 *   [21:16] pos     scan position i (index into W[set][], mb_decoder.cpp:140-143)
 *   [25:22] blk     block index 0..11 inside the macroblock
 *   [26]    RAW     intra DC: stored as is, excluded from the mismatch sum (mb_decoder.cpp:160)
 *   [27]    FIRST   non-intra first coefficient coded "1s" (mb_decoder.cpp:79-88):
 *                   val = (3*W[0]*qscale)>>5 with sign, NOT clamped, included in the mismatch sum
 *   [31:28] mbx     the macroblock's column modulo 16: the reconstruction kernel walks the records of up to 16
 *                   consecutive macroblocks of a row as one flat list and finds each record's macroblock from it
 */
typedef uint32_t mp2v_coef_t;
#define MP2V_COEF_RAW    (1u << 26)
#define MP2V_COEF_FIRST  (1u << 27)
#define MP2V_COEF(level, pos, blk, flags) \
    (((uint32_t)(uint16_t)(int16_t)(level)) | ((uint32_t)(pos) << 16) | ((uint32_t)(blk) << 22) | (flags))
#define MP2V_COEF_MB(mbx)  (((uint32_t)(mbx) & 15u) << 28)
#define MP2V_COEF_LEVEL(c) ((int)(int16_t)((c) & 0xffffu))
#define MP2V_COEF_POS(c)   (((c) >> 16) & 63u)
#define MP2V_COEF_BLK(c)   (((c) >> 22) & 15u)

/* ---- per-picture parameters ------------------------------------------------------------------ */
typedef struct mp2v_pic_params {
    uint8_t  W[4][64];             /* quantiser matrices indexed by SCAN POSITION, exactly the
                                      reference's quantiser_matrices (decoder.cpp:185-191):
                                      0 intra, 1 non-intra, 2 chroma intra, 3 chroma non-intra.
                                      Blocks 0..5 use 0/1, blocks 6..11 use 2/3 (reference quirk,
                                      mb_decoder.cpp:177-195).                                   */
    int32_t  picture_coding_type;  /* 1 I, 2 P, 3 B                                              */
    int32_t  alternate_scan;       /* selects g_scan_trans[alt] (scan_c.cpp:4-21)                 */
    int32_t  dst_frame;            /* frame id written by this picture                           */
    int32_t  l0_frame;             /* forward reference frame id, -1 if none                     */
    int32_t  l1_frame;             /* backward reference frame id, -1 if none                    */
    uint32_t n_coef;               /* coefficient records used (arena high-water mark)           */
    uint32_t reserved[2];
} mp2v_pic_params_t;

/* ---- a picture's pinned SoA buffers (owned by the context) ------------------------------------ */
typedef struct mp2v_picture {
    mp2v_pic_params_t* params;     /* one                                                        */
    mp2v_mb_info_t*    mb;         /* mb_count records, raster order                             */
    mp2v_coef_t*       coef;       /* coef_capacity records                                      */
    uint32_t           mb_count;
    uint32_t           coef_capacity;
    int32_t            slot;       /* context-internal                                           */
    int32_t            reserved;
} mp2v_picture_t;

typedef struct mp2v_recon_config {
    int32_t device;                /* CUDA device ordinal                                        */
    int32_t width;                 /* coded luma width, multiple of 16 (decoder_config_t.width)   */
    int32_t height;                /* coded luma height, multiple of 16                          */
    int32_t chroma_format;         /* 1 = 4:2:0, 2 = 4:2:2, 3 = 4:4:4 (mp2v_hdr.h:56-58)          */
    int32_t n_frames;              /* device frame pool size (>= 3)                              */
    int32_t n_pictures;            /* picture slots in flight (pinned + device arenas)           */
    int32_t max_batch;             /* max pictures fused into one launch (0 = default, <= 128)    */
    int32_t flags;                 /* MP2V_RECON_* below                                         */
    uint32_t coef_capacity;        /* coefficient records per picture slot; 0 = worst case
                                      (mb_count * blocks * 64).  submit fails with MP2V_ERR_RANGE
                                      when a picture needs more.                                 */
    uint32_t bitstream_capacity;   /* MP2V_RECON_DEVICE_VLC: bytes of coded slice data one picture
                                      may carry; 0 = max(2 MiB, 128 bytes per macroblock)         */
} mp2v_recon_config_t;

#define MP2V_RECON_VALIDATE   1    /* check vectors / offsets on the host before launch          */
#define MP2V_RECON_DEVICE_VLC 2    /* pictures are handed over as coded slices
                                      (mp2v_recon_submit_slices) and parsed on the device; every
                                      slot's device arena then has the worst-case capacity and
                                      mp2v_picture_t.coef is NULL (no host-side coefficient arena) */
#define MP2V_RECON_AUTO_DOWNLOAD 4 /* queue the copy of every submitted picture's frame into its
                                      pinned mirror right behind the picture's launch, so that the
                                      copies overlap later launches; mp2v_recon_map_frame then only
                                      waits for that copy (for decoders that show every frame)     */

#define MP2V_RECON_THROUGHPUT 8    /* launch in full lots from the first picture on.  By default the lot size ramps up
                                      (1, 2, 4, ... pictures) after every sync so that the first frames of a decode leave
                                      early; consumers that only care about the rate (frames staying on the device) set this */

typedef struct mp2v_recon mp2v_recon_t;

/* frame geometry: the reference's frame_c rule (decoder.cpp:44-66) */
typedef struct mp2v_frame_layout {
    int32_t width[3], height[3], stride[3];
    size_t  plane_offset[3];
    size_t  bytes;                 /* one frame, all planes, 256-byte aligned planes             */
} mp2v_frame_layout_t;

MP2V_API int  mp2v_frame_layout(int width, int height, int chroma_format, mp2v_frame_layout_t* out);

MP2V_API int  mp2v_recon_create(const mp2v_recon_config_t* cfg, mp2v_recon_t** out);
MP2V_API void mp2v_recon_destroy(mp2v_recon_t* ctx);
MP2V_API const char* mp2v_recon_last_error(mp2v_recon_t* ctx);   /* ctx may be NULL: creation errors */

/* Blocks until a picture slot is free; the returned buffers stay valid until the picture's
 * reconstruction has been issued (submit) or it is given back with release_picture. */
MP2V_API int  mp2v_recon_acquire_picture(mp2v_recon_t* ctx, mp2v_picture_t** out);
MP2V_API int  mp2v_recon_release_picture(mp2v_recon_t* ctx, mp2v_picture_t* pic);

/* Queue one filled picture: H2D of its records + reconstruction.  Pictures must be submitted in
 * coded order (references before the pictures that use them).  Asynchronous; consecutive
 * submissions that do not depend on one another are fused into one launch, and launches are issued once
 * max_batch pictures are queued (or by flush / sync / a wait for one of their frames). */
MP2V_API int  mp2v_recon_submit(mp2v_recon_t* ctx, mp2v_picture_t* pic);
/* Optional: run submit's record validation + byte accounting now, from any thread, without taking the
 * context lock (the picture still belongs to the caller); submit then skips it. */
MP2V_API int  mp2v_recon_precheck(mp2v_recon_t* ctx, mp2v_picture_t* pic);

/* ---- device-side slice parsing (contexts created with MP2V_RECON_DEVICE_VLC) ------------------
 * Instead of filling mb[] / coef[] on the host, hand over the picture's coded slices: the library
 * copies the bytes to the device and a kernel does what the reference's parse_macroblock /
 * parse_block do up to (not including) dequantisation (mb_decoder.cpp:74-155, 521-641) -- one
 * thread per slice, all slices of all pictures in flight at once -- writing the same records the
 * host parser would.  The caller fills params (W, picture_coding_type, alternate_scan, frame ids)
 * as for mp2v_recon_submit and passes what the slice layer needs of picture_coding_extension.
 * Envelope: frame pictures with frame prediction, at most one slice per macroblock row.
 * Asynchronous like submit; a syntax error inside a slice (or a vector leaving the frame) is
 * reported by the next sync / map_frame / download_frame / acquire as MP2V_ERR_RANGE, and the
 * macroblocks after the error reconstruct as blank intra macroblocks.                            */
typedef struct mp2v_pic_syntax {
    int32_t f_code[2][2];          /* [forward, backward][horizontal, vertical]                   */
    int32_t intra_dc_precision;    /* 0..3                                                        */
    int32_t q_scale_type;
    int32_t intra_vlc_format;
    int32_t field_dct_syntax;      /* 1: frame_pred_frame_dct = 0 -- macroblock_modes carry frame_motion_type (only
                                      frame-based prediction is accepted) and dct_type (mb_decoder.cpp:349-360)   */
} mp2v_pic_syntax_t;
typedef struct mp2v_slice_ref {
    const uint8_t* payload;        /* first byte after the 4-byte slice start code               */
    uint32_t bytes;                /* up to the next start code prefix                            */
    int32_t  code;                 /* slice_start_code value 0x01..0xAF                           */
} mp2v_slice_ref_t;
MP2V_API int  mp2v_recon_submit_slices(mp2v_recon_t* ctx, mp2v_picture_t* pic, const mp2v_pic_syntax_t* syntax,
                                       const mp2v_slice_ref_t* slices, int n_slices);
/* The same in two steps, for callers that prepare pictures on several threads: stage_slices
 * validates the arguments and copies the coded bytes into the picture's pinned staging buffer (no
 * device work, no context lock: any thread, any order, disjoint pictures); submit_staged issues
 * the copy + parse + reconstruction and, like submit, must be called in coded order.             */
MP2V_API int  mp2v_recon_stage_slices(mp2v_recon_t* ctx, mp2v_picture_t* pic, const mp2v_pic_syntax_t* syntax,
                                      const mp2v_slice_ref_t* slices, int n_slices);
MP2V_API int  mp2v_recon_submit_staged(mp2v_recon_t* ctx, mp2v_picture_t* pic);

/* ---- stream-resident front end (contexts created with MP2V_RECON_DEVICE_VLC) ------------------
 * For callers that hold a whole elementary stream (the reference's decode(buffer, len), decoder.h:100): the stream is
 * copied to the device ONCE, a kernel lists its start codes -- the reference's scan_start_codes (start_codes_search.hpp:7-26)
 * -- and pictures are then handed over as byte offsets of their slices' start codes in that resident copy.  No byte of
 * the stream is touched by the host besides the headers it parses, there is no per-picture staging copy, and one kernel
 * launch parses the slices of every picture handed over since the previous launch.
 *   stream_begin   copies data[0, bytes) (or only the given ranges of it) and scans for start codes: scan = 1 returns the
 *                  ascending byte offsets of every 00 00 01 prefix (host memory owned by the context, valid until the next
 *                  stream_begin); scan = 2 only launches the scan -- fetch the list with mp2v_recon_stream_codes (several
 *                  devices then copy and scan their parts of one stream at the same time); scan = 0 copies only.  With
 *                  ranges, the scan covers ranges[0] alone (it must start on a 16-byte boundary): the codes that BEGIN in
 *                  it, offsets relative to data.  `data` must stay valid until the pictures of this stream have been
 *                  submitted (slice start codes are validated from it).  Pass page-locked memory for the full copy rate.
 *                  Waits for the device parses of the previous stream.
 *   stream_codes   waits for a scan launched with scan = 2 and returns its list.
 *   stream_add     copies further ranges of the same stream (a device's own pictures, once the stream is indexed); the
 *                  parse launches that follow wait for them.
 *   submit_stream_picture   like mp2v_recon_submit_slices with the slices named by offset; asynchronous, coded order.
 * GOP sharding over N devices (what the bundled decoder does): split the stream into N parts; stream_begin(part d, scan = 2)
 * on every device, stream_codes on every device, concatenate the lists (they are ascending), index, then stream_add to every
 * device the byte ranges of the pictures it will decode. */
typedef struct mp2v_byte_range { size_t offset, bytes; } mp2v_byte_range_t;
MP2V_API int  mp2v_recon_stream_begin(mp2v_recon_t* ctx, const uint8_t* data, size_t bytes, const mp2v_byte_range_t* ranges, int n_ranges,
                                      int scan, const uint32_t** codes, uint32_t* n_codes);
MP2V_API int  mp2v_recon_stream_codes(mp2v_recon_t* ctx, const uint32_t** codes, uint32_t* n_codes);
MP2V_API int  mp2v_recon_stream_add(mp2v_recon_t* ctx, const mp2v_byte_range_t* ranges, int n_ranges);
MP2V_API int  mp2v_recon_submit_stream_picture(mp2v_recon_t* ctx, mp2v_picture_t* pic, const mp2v_pic_syntax_t* syntax,
                                               const uint32_t* slice_offsets, int n_slices);

MP2V_API int  mp2v_recon_flush(mp2v_recon_t* ctx);               /* launch whatever is queued       */
MP2V_API int  mp2v_recon_sync(mp2v_recon_t* ctx);                /* flush + wait for the device     */
/* After an error: wait for whatever was issued, give every picture slot back, forget queued pictures,
 * frame contents and the pending error -- the context is as after creation, without re-allocating. */
MP2V_API int  mp2v_recon_reset(mp2v_recon_t* ctx);

/* Device-resident mode (benchmarks, re-decode): copy a filled picture to its device arena once,
 * keep the slot, then reconstruct any list of resident pictures without host traffic.
 * levels[] (optional) gives each picture's dependency level: pictures of equal level are fused into
 * one launch, levels run in increasing order. */
MP2V_API int  mp2v_recon_upload(mp2v_recon_t* ctx, mp2v_picture_t* pic);
MP2V_API int  mp2v_recon_run_resident(mp2v_recon_t* ctx, mp2v_picture_t* const* pics, const int32_t* levels, int n);

/* Copy a reconstructed frame to host memory (waits for the pictures that write it).
 * dst[p] / dst_stride[p]: caller buffers (pinned or pageable); rows are width[p] bytes. */
MP2V_API int  mp2v_recon_download_frame(mp2v_recon_t* ctx, int frame_id, uint8_t* const dst[3], const int32_t dst_stride[3]);
/* Same, into the context's own pinned mirror of that frame; planes[] valid until the next map of
 * the same frame id. */
MP2V_API int  mp2v_recon_map_frame(mp2v_recon_t* ctx, int frame_id, uint8_t* planes[3], int32_t strides[3]);
/* Test hook: fill a device frame from host planes (used to seed references in kernel unit tests). */
MP2V_API int  mp2v_recon_upload_frame(mp2v_recon_t* ctx, int frame_id, const uint8_t* const src[3], const int32_t src_stride[3]);

/* Device pointers of a frame's planes (zero-copy consumers, e.g. a CUDA renderer). */
MP2V_API int  mp2v_recon_frame_device_ptrs(mp2v_recon_t* ctx, int frame_id, void* planes[3], int32_t strides[3]);

/* Output path for GPU consumers: convert reconstructed frames into the layout the next stage on
 * the GPU takes, into the caller's DEVICE buffers (one pointer per frame, rows `dst_pitch` bytes apart), ordered on
 * the context's compute stream behind the pictures that write the frames; mp2v_recon_sync (or any later work on that
 * stream) completes it.  Replaces the host-side planar write of the reference's sample
 * (tiny_decoder/tiny_mp2v_dec.cpp:11-17); with mp2v_b200_options_t::download_frames = false no frame crosses PCIe.
 *   MP2V_OUT_NV12  4:2:0: `height` rows of Y, then `height / 2` rows of interleaved Cb/Cr pairs (row = width bytes)
 *   MP2V_OUT_P010  4:2:0: the same layout with 16-bit little-endian samples, the decoded 8 bits in the high byte
 *                  (row = 2 * width bytes)
 *   MP2V_OUT_UYVY  4:2:2: `height` rows of packed Cb Y0 Cr Y1 (row = 2 * width bytes)                              */
enum { MP2V_OUT_NV12 = 0, MP2V_OUT_P010 = 1, MP2V_OUT_UYVY = 2 };
MP2V_API int  mp2v_recon_convert_frames(mp2v_recon_t* ctx, int format, const int32_t* frame_ids, void* const* dst_device, int n, int32_t dst_pitch);
MP2V_API int  mp2v_recon_convert_frame_nv12(mp2v_recon_t* ctx, int frame_id, void* dst_device, int32_t dst_pitch);
MP2V_API int  mp2v_recon_convert_frames_nv12(mp2v_recon_t* ctx, const int32_t* frame_ids, void* const* dst_device, int n, int32_t dst_pitch);
/* Block until the reconstruction of a frame has finished on the device (consumers that read
 * mp2v_recon_frame_device_ptrs' planes from their own streams); reports a slice error of its picture. */
MP2V_API int  mp2v_recon_wait_frame(mp2v_recon_t* ctx, int frame_id);

/* Statistics of the context since creation / last reset. */
typedef struct mp2v_recon_stats {
    uint64_t pictures;             /* pictures reconstructed                                      */
    uint64_t launches;             /* kernel launches issued                                      */
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    uint64_t algorithmic_bytes;    /* OUT + REF + COEF(128 B/coded block) + META   */
    double   kernel_ms;            /* CUDA-event time of the reconstruction launches (when
                                      timing is enabled)                                          */
    uint64_t vlc_launches;         /* slice parser kernel launches (one per picture handed over with submit_slices / submit_staged,
                                      one per batch of pictures handed over with submit_stream_picture) */
    uint64_t vlc_slices;           /* slices handed to the device parser                          */
    uint64_t vlc_coefs;            /* coefficient records it wrote (parses completed so far)      */
    uint64_t idct_batches;         /* batches of <= 24 coded blocks transformed (when timing is enabled) ...            */
    uint64_t idct_exact_pass2;     /* ... of them those whose range bound forced the saturating arithmetic in pass 2 ...  */
    uint64_t idct_exact_pass1;     /* ... and in pass 1 too (an intra DC outside the analysed range)                      */
} mp2v_recon_stats_t;
MP2V_API int  mp2v_recon_set_timing(mp2v_recon_t* ctx, int enable);
/* CUDA-event stopwatch on the context's compute stream (the stream every kernel is launched on):
 * start records an event behind the work queued so far, stop records another, waits for it and
 * returns the elapsed device time in milliseconds. */
MP2V_API int  mp2v_recon_timer_start(mp2v_recon_t* ctx);
MP2V_API int  mp2v_recon_timer_stop(mp2v_recon_t* ctx, double* elapsed_ms);
MP2V_API int  mp2v_recon_get_stats(mp2v_recon_t* ctx, mp2v_recon_stats_t* out, int reset);

/* NUMA placement (multi-socket hosts): the context's pinned memory is allocated on the NUMA node of the device's PCIe
 * root; threads that feed the context or read its mapped frames should run there too (the bundled decoder binds its
 * feeder and output threads).  Returns that node, or -1 on a single-node host / when the kernel does not say /
 * with MP2V_NUMA=0 -- then nothing is bound.
 * mp2v_numa_parse_cpu_list: "0-3,8,10-11" -> 0 1 2 3 8 10 11 (the sysfs cpulist format); returns the number of CPUs
 * (at most cap are stored), 0 for malformed input. */
MP2V_API int  mp2v_recon_numa_node(mp2v_recon_t* ctx);
MP2V_API int  mp2v_numa_parse_cpu_list(const char* list, int32_t* cpus, int cap);

#ifdef __cplusplus
}
#endif
#endif /* MP2V_RECON_H */
