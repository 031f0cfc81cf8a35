// Sequence-level headers of an MPEG-2 video elementary stream as the decode API exposes them
// (public members of mp2v_decoder_c, reference: src/core/decoder.h:120-131).  Field names are the syntax
// element names of ISO/IEC 13818-2 6.2.2.1 / 6.2.2.3-6.2.2.6, which is also what the reference's
// src/core/mp2v_hdr.h:60-131 calls them, so client code reading e.g.
// `dec.m_sequence_header.frame_rate_code` compiles against either library.  Every element is held in a
// uint32_t whatever its width in the stream; marker bits are not stored.
#pragma once
#include <cstdint>

struct sequence_header_t {                    // 6.2.2.1
    uint32_t sequence_header_code;            // 0x000001B3
    uint32_t horizontal_size_value;           // 12 bits
    uint32_t vertical_size_value;             // 12
    uint32_t aspect_ratio_information;        // 4
    uint32_t frame_rate_code;                 // 4
    uint32_t bit_rate_value;                  // 18
    uint32_t vbv_buffer_size_value;           // 10
    uint32_t constrained_parameters_flag;     // 1
    uint32_t load_intra_quantiser_matrix;     // 1
    uint8_t  intra_quantiser_matrix[64];      // zig-zag order, valid when loaded
    uint32_t load_non_intra_quantiser_matrix; // 1
    uint8_t  non_intra_quantiser_matrix[64];
};

struct sequence_extension_t {                 // 6.2.2.3
    uint32_t extension_start_code;            // 0x000001B5
    uint32_t extension_start_code_identifier; // 4 bits, = 1
    uint32_t profile_and_level_indication;    // 8
    uint32_t progressive_sequence;            // 1
    uint32_t chroma_format;                   // 2
    uint32_t horizontal_size_extension;       // 2
    uint32_t vertical_size_extension;         // 2
    uint32_t bit_rate_extension;              // 12
    uint32_t vbv_buffer_size_extension;       // 8
    uint32_t low_delay;                       // 1
    uint32_t frame_rate_extension_n;          // 2
    uint32_t frame_rate_extension_d;          // 5
};

struct sequence_display_extension_t {         // 6.2.2.4
    uint32_t extension_start_code_identifier; // = 2
    uint32_t video_format;                    // 3
    uint32_t colour_description;              // 1
    uint32_t colour_primaries;                // 8, present when colour_description
    uint32_t transfer_characteristics;        // 8
    uint32_t matrix_coefficients;             // 8
    uint32_t display_horizontal_size;         // 14
    uint32_t display_vertical_size;           // 14
};

struct sequence_scalable_extension_t {        // 6.2.2.5
    uint32_t extension_start_code_identifier; // = 5
    uint32_t scalable_mode;                   // 2
    uint32_t layer_id;                        // 4
    uint32_t lower_layer_prediction_horizontal_size;   // 14, spatial scalability
    uint32_t lower_layer_prediction_vertical_size;     // 14
    uint32_t horizontal_subsampling_factor_m; // 5
    uint32_t horizontal_subsampling_factor_n; // 5
    uint32_t vertical_subsampling_factor_m;   // 5
    uint32_t vertical_subsampling_factor_n;   // 5
    uint32_t picture_mux_enable;              // 1, temporal scalability
    uint32_t mux_to_progressive_sequence;     // 1
    uint32_t picture_mux_order;               // 3
    uint32_t picture_mux_factor;              // 3
};

struct group_of_pictures_header_t {           // 6.2.2.6
    uint32_t group_start_code;                // 0x000001B8
    uint32_t time_code;                       // 25 bits
    uint32_t closed_gop;                      // 1
    uint32_t broken_link;                     // 1
};
