// Drop-in C++ decode API of the B200 back end: the same three public types as the reference's
// src/core/decoder.h -- decoder_config_t (:25-32), frame_c (:34-49), mp2v_decoder_c (:82-131) -- with
// the same constructor / decode / renderer conventions, implemented on top of the C ABI in
// mp2v_recon.h (host slice parsing on num_threads threads, reconstruction on the GPU).
//
// Conventions kept from the reference:
//   * the caller owns `buffer` (a whole elementary stream) and pads it with >= 64 readable bytes;
//     it must stay valid until decode() returns; `len` is int
//   * decode() is one-shot and blocks until every frame has been rendered
//   * `renderer` runs on an internal thread, in display order when `reordering`, and receives a
//     frame_c* exposing HOST plane pointers that are valid only during the call
//   * geometry and chroma_format come from decoder_config_t, not from the sequence header
// Differences: decode() returns false (and last_error() says why) instead of undefined behaviour on
// malformed input; streams without a quant_matrix_extension decode with the sequence-header /
// default matrices instead of crashing (decoder.cpp:187 dereferences nullptr there).
#pragma once
#include <cstdint>
#include <fstream>      // the reference's decoder.h pulls this in (bitstream.h:3); its sample relies on it
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "mp2v_stream_headers.h"

#if defined(__GNUC__)
#define MP2V_CXX_API __attribute__((visibility("default")))
#else
#define MP2V_CXX_API
#endif

constexpr int MAX_NUM_THREADS = 256;

struct decoder_config_t {
    int width;
    int height;
    int chroma_format;
    int pictures_pool_size;
    int num_threads;
    bool reordering;
};

class MP2V_CXX_API frame_c {
public:
    frame_c(int width, int height, int chroma_format);   // owning host frame, reference geometry (decoder.cpp:44-87)
    ~frame_c();
    frame_c(const frame_c&) = delete;
    frame_c& operator=(const frame_c&) = delete;

    uint8_t* get_planes (int plane_idx) { return m_planes[plane_idx]; }
    int      get_strides(int plane_idx) { return (int)m_stride[plane_idx]; }
    int      get_width  (int plane_idx) { return (int)m_width [plane_idx]; }
    int      get_height (int plane_idx) { return (int)m_height[plane_idx]; }

    // view over planes owned elsewhere (the decoder's pinned frame mirrors)
    frame_c(int width, int height, int chroma_format, uint8_t* const planes[3], const int strides[3]);
private:
    uint32_t m_width [3] = { 0 };
    uint32_t m_height[3] = { 0 };
    uint32_t m_stride[3] = { 0 };
    uint8_t* m_planes[3] = { 0 };
    bool m_owner = false;
};

// A decoded frame where it was reconstructed: plane pointers in DEVICE memory of CUDA device `device`, complete (the
// reconstruction has finished) and valid until the callback returns.  `recon` / `frame_id` name it for
// mp2v_recon_convert_frames (NV12 / P010 / UYVY into the consumer's own device buffer, mp2v_recon.h).
struct mp2v_recon;
struct mp2v_device_frame_t {
    void* planes[3];
    int strides[3], width[3], height[3];
    int device;
    int frame_id;
    mp2v_recon* recon;
};

// Knobs the reference does not have; all optional.
struct mp2v_b200_options_t {
    std::vector<int> devices = {0};   // CUDA ordinals; more than one = closed GOPs round-robin over the devices
    int max_batch = 8;                // pictures fused into one launch
    int output_lag = 4;               // pictures the display side stays behind the submit side (lets launches batch)
    bool download_frames = true;      // false: renderer gets frames whose planes were not copied back (benchmarks)
    // Consumer on the GPU: called on the output thread for every frame in display order, like the renderer, but with
    // the frame's DEVICE planes -- set download_frames = false and no decoded pixel crosses PCIe (the reference's only
    // output path is the host-side planar write of its sample, tiny_decoder/tiny_mp2v_dec.cpp:11-17).
    std::function<void(const mp2v_device_frame_t&)> device_renderer;
    bool gpu_vlc = true;              // slices are parsed on the device (mp2v_recon_submit_slices): the host only finds
                                      // start codes and parses headers.  Streams outside the device parser's envelope
                                      // (several slices in one macroblock row, oversized pictures) and gpu_vlc = false
                                      // take the host slice parser on num_threads worker threads instead.
};

class mp2v_picture_c;   // the reference's per-picture task object (decoder.h:56-80); only ever named by flush()

class MP2V_CXX_API mp2v_decoder_c {
public:
    mp2v_decoder_c();
    mp2v_decoder_c(const decoder_config_t& config, std::function<void(frame_c*)> renderer);
    ~mp2v_decoder_c();
    bool decoder_init(const decoder_config_t& config, std::function<void(frame_c*)> renderer);
    bool decode(uint8_t* buffer, int len);
    // decoder.h:101.  decode() is one-shot and drains everything itself (the reference's decode() ends with
    // flush(cur_pic), decoder.cpp:326-327), so there is never anything left to flush; the argument is ignored.
    void flush(mp2v_picture_c* cur_pic = nullptr);

    // headers & user data of the stream last given to decode() (decoder.h:120-131): the last occurrence of
    // each header, all user_data() bytes in stream order; the optional ones stay nullptr when the stream has none
    std::vector<uint8_t> user_data;
    sequence_header_t m_sequence_header = {};
    sequence_extension_t m_sequence_extension = {};
    sequence_display_extension_t* m_sequence_display_extension = nullptr;
    sequence_scalable_extension_t* m_sequence_scalable_extension = nullptr;
    group_of_pictures_header_t* m_group_of_pictures_header = nullptr;

    // extensions
    // Decode once more the stream the last successful decode() left resident on the device(s) (device slice parsing
    // only): no upload, no start-code scan, the renderer is called for every frame again.  The buffer given to that
    // decode() must still be valid.  (Re-decode, and the device-side figure of the benchmark.)
    bool decode_resident();
    void set_options(const mp2v_b200_options_t& opt);
    void set_device_renderer(std::function<void(const mp2v_device_frame_t&)> r);   // mp2v_b200_options_t::device_renderer alone; keeps the device contexts
    bool prepare();                   // allocate the device contexts now (otherwise the first decode() does)
    const char* last_error() const;
    struct stats_t {
        uint64_t pictures = 0, launches = 0, h2d_bytes = 0, d2h_bytes = 0, algorithmic_bytes = 0;
        double kernel_ms = 0;          // CUDA-event time of the reconstruction launches
        uint64_t vlc_launches = 0;     // slice parser kernel launches (0: the host parser was used)
        double parse_cpu_seconds = 0;  // summed over worker threads: slice parsing only
        double wall_seconds = 0;       // decode() wall clock
        double device_ms = 0;          // CUDA-event time from the call's first to its last device work (max over devices)
    };
    stats_t stats() const;

private:
    struct impl_t;
    std::unique_ptr<impl_t> m;
};
