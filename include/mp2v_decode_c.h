/*
 * mp2v_decode_c.h -- plain-C entry points over the C++ decode API (mp2v_decoder.hpp), for bindings
 * that cannot hold a C++ class (ctypes, cgo, JNI).  They wrap exactly what a user of the reference's
 * sample does (tiny_decoder/tiny_mp2v_dec.cpp:36-59): construct mp2v_decoder_c with a
 * decoder_config_t and a renderer, call decode(buffer, len) once, collect the frames.
 */
#ifndef MP2V_DECODE_C_H
#define MP2V_DECODE_C_H
#include <stddef.h>
#include <stdint.h>
#include "mp2v_recon.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mp2v_decode_params {
    int32_t width, height, chroma_format;   /* decoder_config_t (decoder.h:25-32)                          */
    int32_t pictures_pool_size, num_threads, reordering;
    int32_t n_devices;                      /* 0 or 1: devices[0]; > 1: closed GOPs round-robin            */
    int32_t devices[8];
    int32_t max_batch, output_lag;          /* 0 = defaults                                                 */
    int32_t download_frames;                /* 0: reconstruct only (frames stay on the device)             */
    int32_t hash_output;                    /* 1: FNV-1a of the output into stats->hash (slow; tests only)  */
    int32_t host_parser;                    /* 1: parse slices on the host (mp2v_b200_options_t.gpu_vlc = false);
                                               0: on the device whenever the stream is inside its envelope   */
} mp2v_decode_params_t;

typedef struct mp2v_decode_stats {
    uint64_t frames, pictures, launches, h2d_bytes, d2h_bytes, algorithmic_bytes;
    double kernel_ms, parse_cpu_seconds, wall_seconds;
    uint64_t hash;                          /* FNV-1a 64 of the cropped planar output, display order        */
    uint64_t vlc_launches;                  /* slice parser kernel launches (0: the host parser was used)   */
    double device_ms;                       /* CUDA-event time between the call's first and last device work  */
} mp2v_decode_stats_t;

/* per-frame callback, invoked on the decoder's output thread in display order */
typedef void (*mp2v_frame_fn)(void* user, uint8_t* const planes[3], const int32_t strides[3],
                              const int32_t widths[3], const int32_t heights[3]);

/* per-frame callback with the frame's planes in DEVICE memory (mp2v_b200_options_t::device_renderer): complete and valid
 * during the call; recon / frame_id name the frame for mp2v_recon_convert_frames */
typedef void (*mp2v_device_frame_fn)(void* user, void* const planes[3], const int32_t strides[3], const int32_t widths[3],
                                     const int32_t heights[3], int32_t device, int32_t frame_id, mp2v_recon_t* recon);

/* Decode a whole elementary stream (buffer padded with >= 64 readable bytes).  Frames go to `fn` when
 * given; when `out` is given the cropped planar YUV (Y, Cb, Cr per frame, display order) is also
 * stored there (up to out_cap bytes; *out_bytes = bytes produced).  Returns MP2V_OK or an error code;
 * err (optional) receives the message. */
MP2V_API int mp2v_decode_stream(const mp2v_decode_params_t* params, uint8_t* buffer, int len,
                                mp2v_frame_fn fn, void* user, uint8_t* out, size_t out_cap, size_t* out_bytes,
                                mp2v_decode_stats_t* stats, char* err, size_t err_len);

/* The same with a decoder that outlives the call: create allocates the device contexts (the
 * reference allocates its frame pool in the constructor), decode may be called repeatedly. */
typedef struct mp2v_decoder mp2v_decoder_t;
MP2V_API int mp2v_decoder_create(const mp2v_decode_params_t* params, mp2v_decoder_t** out, char* err, size_t err_len);
MP2V_API int mp2v_decoder_decode(mp2v_decoder_t* dec, uint8_t* buffer, int len, mp2v_frame_fn fn, void* user,
                                 uint8_t* out, size_t out_cap, size_t* out_bytes, mp2v_decode_stats_t* stats,
                                 char* err, size_t err_len);
/* decode once more the stream the last mp2v_decoder_decode left resident on the device (mp2v_decoder_c::decode_resident) */
MP2V_API int mp2v_decoder_decode_resident(mp2v_decoder_t* dec, mp2v_frame_fn fn, void* user, uint8_t* out, size_t out_cap, size_t* out_bytes,
                                          mp2v_decode_stats_t* stats, char* err, size_t err_len);
/* install (fn != NULL) or remove the device-side consumer of the decoder's frames; keeps the device contexts */
MP2V_API int mp2v_decoder_set_device_renderer(mp2v_decoder_t* dec, mp2v_device_frame_fn fn, void* user);
MP2V_API void mp2v_decoder_destroy(mp2v_decoder_t* dec);

/* Host-only: index + slice-parse a stream into reconstruction records on `threads` threads, no GPU
 * involved (parser tests; the "host parse" figure of the benchmark). */
typedef struct mp2v_parsed mp2v_parsed_t;
MP2V_API int mp2v_parse_stream(const uint8_t* buffer, int len, int width, int height, int chroma_format,
                               int threads, mp2v_parsed_t** out, char* err, size_t err_len);
MP2V_API int mp2v_parsed_num_pictures(const mp2v_parsed_t* p);
MP2V_API int mp2v_parsed_picture(const mp2v_parsed_t* p, int coded_index, mp2v_pic_params_t* params,
                                 const mp2v_mb_info_t** mb, const mp2v_coef_t** coef, uint32_t* n_coef,
                                 int32_t* temporal_reference, int32_t* gop);
MP2V_API double mp2v_parsed_wall_seconds(const mp2v_parsed_t* p);     /* slice parsing, wall clock          */
MP2V_API double mp2v_parsed_cpu_seconds(const mp2v_parsed_t* p);      /* summed over threads                */
MP2V_API void mp2v_parsed_free(mp2v_parsed_t* p);

#ifdef __cplusplus
}
#endif
#endif
