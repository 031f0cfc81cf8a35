#include "stream_index.h"

#include "bitreader.h"

namespace mp2v {

bool index_stream(const uint8_t* buffer, size_t len, stream_index_t& out) {
    out.pictures.clear();
    out.error.clear();
    const uint8_t* end = buffer + len;
    sequence_info_t seq;
    coded_picture_t* cur = nullptr;
    int gop = 0;
    bool have_picture = false;      // a picture has been seen since the last chain boundary
    const uint8_t* p = find_start_code(buffer, end);
    while (p + 4 <= end) {
        const int code = p[3];
        const uint8_t* payload = p + 4;
        if (code == 0xB3) {                              // sequence_header
            if (!parse_sequence_header(payload, seq)) { out.error = "bad sequence_header"; return false; }
            cur = nullptr;
        } else if (code == 0xB5) {                       // extension_start_code
            if (!parse_extension(payload, seq, cur ? &cur->info : nullptr)) {
                // picture-level extension before any picture header, or a bad value
                if (cur || (payload[0] >> 4) == 1) { out.error = "bad extension data"; return false; }
            }
            if (cur) cur->seq = seq;
        } else if (code == 0xB8) {                       // group_start_code: time_code(25) closed_gop(1) broken_link(1)
            bitreader_t br(payload);
            br.get(25);
            const bool closed = br.get1() != 0;
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
            cur->slices.push_back({payload, code});
        } else if (code == 0xB7 || code == 0xB4) {       // sequence_end / sequence_error
            cur = nullptr;
        }
        p = find_start_code(p + 3, end);
    }
    out.n_gops = out.pictures.empty() ? 0 : gop + 1;
    for (auto& pic : out.pictures)
        if (!pic.info.have_coding_extension) { out.error = "picture without picture_coding_extension (MPEG-1 streams are not supported)"; return false; }
    return true;
}

}  // namespace mp2v
