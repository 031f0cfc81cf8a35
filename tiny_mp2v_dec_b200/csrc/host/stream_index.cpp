#include "stream_index.h"

#include <thread>

#include "bitreader.h"

namespace mp2v {

// offsets of every start code prefix in [buffer, buffer+len), ascending
static std::vector<uint32_t> scan_start_codes(const uint8_t* buffer, size_t len, int threads) {
    const size_t kMinChunk = 256 * 1024;
    size_t n = threads < 1 ? 1 : (size_t)threads;
    if (len / kMinChunk + 1 < n) n = len / kMinChunk + 1;
    std::vector<std::vector<uint32_t>> found(n);
    auto scan = [&](size_t part) {
        const size_t lo = len * part / n, hi = len * (part + 1) / n;
        // a prefix starting in [lo, hi) may extend two bytes past hi; one starting before lo belongs to the previous part
        const uint8_t* end = buffer + (hi + 2 < len ? hi + 2 : len);
        for (const uint8_t* p = find_start_code(buffer + lo, end); p + 3 <= end && p < buffer + hi; p = find_start_code(p + 3, end))
            found[part].push_back((uint32_t)(p - buffer));
    };
    std::vector<std::thread> th;
    for (size_t i = 1; i < n; i++) th.emplace_back(scan, i);
    scan(0);
    for (auto& t : th) t.join();
    std::vector<uint32_t> all;
    size_t total = 0;
    for (auto& f : found) total += f.size();
    all.reserve(total);
    for (auto& f : found) all.insert(all.end(), f.begin(), f.end());
    return all;
}

// ---- the published copies of the sequence-level headers (ISO/IEC 13818-2 6.2.2; element widths in mp2v_stream_headers.h)
static void read_sequence_header(const uint8_t* payload, sequence_header_t& h) {
    bitreader_t br(payload);
    h = sequence_header_t{};
    h.sequence_header_code = 0x000001B3u;
    h.horizontal_size_value = br.get(12); h.vertical_size_value = br.get(12);
    h.aspect_ratio_information = br.get(4); h.frame_rate_code = br.get(4);
    h.bit_rate_value = br.get(18); br.get(1);
    h.vbv_buffer_size_value = br.get(10); h.constrained_parameters_flag = br.get(1);
    if ((h.load_intra_quantiser_matrix = br.get(1)) != 0) for (auto& v : h.intra_quantiser_matrix) v = (uint8_t)br.get(8);
    if ((h.load_non_intra_quantiser_matrix = br.get(1)) != 0) for (auto& v : h.non_intra_quantiser_matrix) v = (uint8_t)br.get(8);
}
static void read_sequence_level_extension(const uint8_t* payload, stream_headers_t& hd) {
    bitreader_t br(payload);
    const uint32_t id = br.get(4);
    if (id == 1) {
        sequence_extension_t& e = hd.sequence_extension;
        e = sequence_extension_t{};
        e.extension_start_code = 0x000001B5u; e.extension_start_code_identifier = id;
        e.profile_and_level_indication = br.get(8); e.progressive_sequence = br.get(1); e.chroma_format = br.get(2);
        e.horizontal_size_extension = br.get(2); e.vertical_size_extension = br.get(2);
        e.bit_rate_extension = br.get(12); br.get(1);
        e.vbv_buffer_size_extension = br.get(8); e.low_delay = br.get(1);
        e.frame_rate_extension_n = br.get(2); e.frame_rate_extension_d = br.get(5);
    } else if (id == 2) {
        sequence_display_extension_t& e = hd.sequence_display_extension;
        e = sequence_display_extension_t{};
        e.extension_start_code_identifier = id;
        e.video_format = br.get(3);
        if ((e.colour_description = br.get(1)) != 0) { e.colour_primaries = br.get(8); e.transfer_characteristics = br.get(8); e.matrix_coefficients = br.get(8); }
        e.display_horizontal_size = br.get(14); br.get(1);
        e.display_vertical_size = br.get(14);
        hd.have_display_extension = true;
    } else if (id == 5) {
        sequence_scalable_extension_t& e = hd.sequence_scalable_extension;
        e = sequence_scalable_extension_t{};
        e.extension_start_code_identifier = id;
        e.scalable_mode = br.get(2); e.layer_id = br.get(4);
        if (e.scalable_mode == 1) {            // spatial scalability
            e.lower_layer_prediction_horizontal_size = br.get(14); br.get(1);
            e.lower_layer_prediction_vertical_size = br.get(14);
            e.horizontal_subsampling_factor_m = br.get(5); e.horizontal_subsampling_factor_n = br.get(5);
            e.vertical_subsampling_factor_m = br.get(5); e.vertical_subsampling_factor_n = br.get(5);
        } else if (e.scalable_mode == 3) {     // temporal scalability
            if ((e.picture_mux_enable = br.get(1)) != 0) e.mux_to_progressive_sequence = br.get(1);
            e.picture_mux_order = br.get(3); e.picture_mux_factor = br.get(3);
        }
        hd.have_scalable_extension = true;
    }
}

bool index_stream(const uint8_t* buffer, size_t len, stream_index_t& out, int threads) {
    const std::vector<uint32_t> codes = scan_start_codes(buffer, len, threads);
    return index_stream_from_codes(buffer, len, codes.data(), codes.size(), out);
}

bool index_stream_from_codes(const uint8_t* buffer, size_t len, const uint32_t* codes, size_t n_codes, stream_index_t& out) {
    out.pictures.clear();
    out.headers = stream_headers_t();
    out.error.clear();
    const uint8_t* end = buffer + len;
    sequence_info_t seq;
    coded_picture_t* cur = nullptr;
    int gop = 0;
    bool have_picture = false;      // a picture has been seen since the last chain boundary
    for (size_t ci = 0; ci < n_codes; ci++) {
        if (codes[ci] >= len || (ci && codes[ci] <= codes[ci - 1])) { out.error = "start code list is not ascending / inside the stream"; return false; }
        const uint8_t* p = buffer + codes[ci];
        if (p + 4 > end) break;
        const int code = p[3];
        const uint8_t* payload = p + 4;
        if (code == 0xB3) {                              // sequence_header
            if (!parse_sequence_header(payload, seq)) { out.error = "bad sequence_header"; return false; }
            read_sequence_header(payload, out.headers.sequence_header);
            cur = nullptr;
        } else if (code == 0xB5) {                       // extension_start_code
            if (!parse_extension(payload, seq, cur ? &cur->info : nullptr)) {
                // picture-level extension before any picture header, or a bad value
                if (cur || (payload[0] >> 4) == 1) { out.error = "bad extension data"; return false; }
            }
            if (cur) cur->seq = seq;
            else read_sequence_level_extension(payload, out.headers);
        } else if (code == 0xB2) {                       // user_data_start_code: bytes up to the next start code
            const uint8_t* next = ci + 1 < n_codes ? buffer + codes[ci + 1] : end;
            out.headers.user_data.insert(out.headers.user_data.end(), payload, next);
        } else if (code == 0xB8) {                       // group_start_code: time_code(25) closed_gop(1) broken_link(1)
            bitreader_t br(payload);
            group_of_pictures_header_t& g = out.headers.group_of_pictures_header;
            g.group_start_code = 0x000001B8u;
            g.time_code = br.get(25);
            const bool closed = br.get1() != 0;
            g.closed_gop = closed; g.broken_link = br.get(1);
            out.headers.have_gop_header = true;
            // a closed GOP starts a new independent chain: no picture after it references one before it
            if (closed && have_picture) { gop++; have_picture = false; }
            cur = nullptr;
        } else if (code == 0x00) {                       // picture_start_code
            out.pictures.emplace_back();
            cur = &out.pictures.back();
            cur->seq = seq;
            cur->gop = gop;
            if (!parse_picture_header(payload, seq, cur->info)) { out.error = "bad picture_header"; return false; }
            have_picture = true;
        } else if (code >= 0x01 && code <= 0xAF) {       // slice
            if (!cur) { out.error = "slice before any picture header"; return false; }
            const uint8_t* next = ci + 1 < n_codes ? buffer + codes[ci + 1] : end;
            cur->slices.push_back({payload, code, (uint32_t)(next - payload)});
        } else if (code == 0xB7 || code == 0xB4) {       // sequence_end / sequence_error
            cur = nullptr;
        }
    }
    out.n_gops = out.pictures.empty() ? 0 : gop + 1;
    for (auto& pic : out.pictures)
        if (!pic.info.have_coding_extension) { out.error = "picture without picture_coding_extension (MPEG-1 streams are not supported)"; return false; }
    return true;
}

}  // namespace mp2v
