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

bool index_stream(const uint8_t* buffer, size_t len, stream_index_t& out, int threads) {
    out.pictures.clear();
    out.error.clear();
    const uint8_t* end = buffer + len;
    const std::vector<uint32_t> codes = scan_start_codes(buffer, len, threads);
    sequence_info_t seq;
    coded_picture_t* cur = nullptr;
    int gop = 0;
    bool have_picture = false;      // a picture has been seen since the last chain boundary
    for (size_t ci = 0; ci < codes.size(); ci++) {
        const uint8_t* p = buffer + codes[ci];
        if (p + 4 > end) break;
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
            const uint8_t* next = ci + 1 < codes.size() ? buffer + codes[ci + 1] : end;
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
