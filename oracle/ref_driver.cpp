// TEST INFRASTRUCTURE ONLY -- never linked into, or called by, the product path.
//
// Drivers over the UNMODIFIED reference library (fxslava/tiny_mp2v_dec), compiled where its sources
// lie under /root/reference by oracle/Makefile into oracle/_ref/.  No reference source is copied into
// this repository; this file only *calls* the reference's public API (src/core/decoder.h).
//
//  * serial decode  -- deterministic parity oracle.  The reference's multi-threaded scheduler has a
//    data race on small pictures (get_decoded() can hand out a not-yet-decoded
//    slot), so parity runs re-walk decode()'s start-code switch (decoder.cpp:278-329) on ONE
//    thread, calling the reference's own public parse_* / mp2v_picture_c::init() /
//    mp2v_picture_c::decode_slice() (decoder.h:57-80) and doing the I/P/B reorder of
//    decoder_output_scheduler (decoder.cpp:346-379) inline.
//  * MT decode      -- the reference exactly as its sample uses it (tiny_mp2v_dec.cpp:48-52):
//    mp2v_decoder_c(cfg, renderer) + decode(); this is the timed CPU baseline.
//
// Exposed both as a C ABI (libmp2v_ref.so, used from Python via ctypes) and as a CLI (ref_decode).
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <chrono>
#include <vector>
#include <string>
#include <functional>
#include "core/decoder.h"

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {

struct yuv_sink_t {
    uint8_t* out = nullptr;
    size_t cap = 0;
    size_t pos = 0;
    uint64_t hash = 1469598103934665603ull;  // FNV-1a 64 over the cropped planar output
    int frames = 0;
    void put(frame_c* f) {
        for (int p = 0; p < 3; p++) {
            uint8_t* row = f->get_planes(p);
            int w = f->get_width(p), h = f->get_height(p), s = f->get_strides(p);
            for (int y = 0; y < h; y++, row += s) {
                for (int x = 0; x < w; x++) { hash ^= row[x]; hash *= 1099511628211ull; }
                if (out && pos + w <= cap) memcpy(out + pos, row, w);
                pos += w;
            }
        }
        frames++;
    }
};

// Serial re-walk of mp2v_decoder_c::decode().  Subclassing gives access to the protected bit reader
// and extension parser; the default constructor starts no threads and leaves task_queue == nullptr.
class serial_decoder_c : public mp2v_decoder_c {
public:
    serial_decoder_c(int w, int h, int cf) {
        for (auto& p : pool) p = new mp2v_picture_c(this, new frame_c(w, h, cf));
    }
    ~serial_decoder_c() {
        for (auto* p : pool) { delete p->get_frame(); delete p; }
    }
    void run(uint8_t* buf, int len, yuv_sink_t& sink) {
        m_bs.set_bitstream_buffer(buf);
        mp2v_picture_c* cur = nullptr;
        mp2v_picture_c* refs[2] = { nullptr, nullptr };
        bool new_picture = false;
        for (int i = 0; i + 3 < len; i++) {
            if (buf[i] != 0 || buf[i + 1] != 0 || buf[i + 2] != 1) continue;
            uint8_t* ptr = buf + i;
            // same re-seat of the bit reader as the reference's lambda (decoder.cpp:285-288)
            m_bs.get_idx() = 32;
            m_bs.get_ptr() = (uint32_t*)(ptr + 4);
            m_bs.get_buf() = (uint64_t)__builtin_bswap32(*((uint32_t*)ptr));
            uint8_t code = ptr[3];
            if (code == sequence_header_code) parse_sequence_header(&m_bs, m_sequence_header);
            else if (code == extension_start_code) decode_extension_data(cur);
            else if (code == group_start_code) { group_of_pictures_header_t g; parse_group_of_pictures_header(&m_bs, g); }
            else if (code == picture_start_code) {
                finish(cur, sink);
                new_picture = true;
                // a pool slot that is neither of the two live references (4 slots: 2 refs + held + cur)
                mp2v_picture_c* slot = nullptr;
                for (auto* p : pool) if (p != refs[0] && p != refs[1]) { slot = p; break; }
                cur = slot;
                cur->reset();
                parse_picture_header(&m_bs, cur->m_picture_header);
                int t = cur->m_picture_header.picture_coding_type;
                if (t == picture_coding_type_pred || t == picture_coding_type_intra) {   // decoder.cpp:298-302
                    cur->add_dependency(refs[1]);
                    refs[0] = refs[1];
                    refs[1] = cur;
                } else {                                                                 // decoder.cpp:303
                    cur->add_dependency(refs[0]);
                    cur->add_dependency(refs[1]);
                }
            }
            else if (code >= slice_start_code_min && code <= slice_start_code_max) {
                if (new_picture) cur->init();
                new_picture = false;
                cur->decode_slice(m_bs);
            }
        }
        finish(cur, sink);
        if (held) sink.put(held->get_frame());
    }
private:
    // Display reorder of decoder_output_scheduler (decoder.cpp:350-378): a B picture is shown when
    // complete, an I/P picture when the NEXT I/P picture is complete (or at the end of the stream).
    void finish(mp2v_picture_c* pic, yuv_sink_t& sink) {
        if (!pic) return;
        if (pic->m_picture_header.picture_coding_type == picture_coding_type_bidir) sink.put(pic->get_frame());
        else { if (held) sink.put(held->get_frame()); held = pic; }
    }
    mp2v_picture_c* pool[4];
    mp2v_picture_c* held = nullptr;
};

}  // namespace

// Decode `len` bytes (caller pads >= 64 zero bytes after len) serially.  Writes cropped planar YUV
// frames in display order to out (if non-null), returns the number of frames; *out_bytes = total
// bytes produced (may exceed cap; then only the first cap bytes were stored), *out_hash = FNV-1a.
REF_API int ref_decode_serial(uint8_t* buf, int len, int width, int height, int chroma_format,
                              uint8_t* out, size_t cap, size_t* out_bytes, uint64_t* out_hash) {
    yuv_sink_t sink; sink.out = out; sink.cap = cap;
    {
        serial_decoder_c dec(width, height, chroma_format);
        dec.run(buf, len, sink);
    }
    if (out_bytes) *out_bytes = sink.pos;
    if (out_hash) *out_hash = sink.hash;
    return sink.frames;
}

// The reference's own multi-threaded decoder, used exactly like tiny_mp2v_dec.cpp:48-52.
// Returns frames rendered; *seconds = steady_clock time around decode() (includes flush/drain).
REF_API int ref_decode_mt(uint8_t* buf, int len, int width, int height, int chroma_format,
                          int pool_size, int num_threads, int want_output,
                          uint8_t* out, size_t cap, size_t* out_bytes, uint64_t* out_hash, double* seconds) {
    yuv_sink_t sink; sink.out = out; sink.cap = cap;
    int frames = 0;
    double secs = 0;
    {
        decoder_config_t cfg = { width, height, chroma_format, pool_size, num_threads, true };
        std::function<void(frame_c*)> renderer;
        if (want_output) renderer = [&sink](frame_c* f) { sink.put(f); };
        else             renderer = [&frames](frame_c*) { frames++; };   // README.md:48: no file output when timing
        mp2v_decoder_c dec(cfg, renderer);
        auto t0 = std::chrono::steady_clock::now();
        dec.decode(buf, len);
        secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    if (want_output) frames = sink.frames;
    if (out_bytes) *out_bytes = sink.pos;
    if (out_hash) *out_hash = sink.hash;
    if (seconds) *seconds = secs;
    return frames;
}

#ifdef REF_DRIVER_MAIN
static std::vector<uint8_t> load(const char* path) {
    std::vector<uint8_t> v;
    FILE* f = fopen(path, "rb");
    if (!f) return v;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    v.assign((size_t)n + 256, 0);
    if (fread(v.data(), 1, n, f) != (size_t)n) v.clear();
    fclose(f);
    return v;
}

// ref_decode <serial|mt> in.m2v width height chroma_format [out.yuv | - | hash] [threads] [pool] [repeat]
int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: %s serial|mt in.m2v W H CF [out.yuv|-] [threads] [pool] [repeat]\n", argv[0]); return 2; }
    std::string mode = argv[1];
    auto bs = load(argv[2]);
    if (bs.empty()) { fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
    int len = (int)bs.size() - 256;
    int w = atoi(argv[3]), h = atoi(argv[4]), cf = atoi(argv[5]);
    // "hash": the renderer hashes every frame (FNV-1a over the cropped planes) but nothing is stored or written
    const bool hash_only = argc > 6 && !strcmp(argv[6], "hash");
    const char* outp = (argc > 6 && strcmp(argv[6], "-") && !hash_only) ? argv[6] : nullptr;
    int threads = argc > 7 ? atoi(argv[7]) : 8, pool = argc > 8 ? atoi(argv[8]) : 10, repeat = argc > 9 ? atoi(argv[9]) : 1;
    size_t fb = (size_t)w * h * (cf == 1 ? 3 : cf == 2 ? 4 : 6) / 2;
    std::vector<uint8_t> out;
    if (outp) out.resize(fb * 4096 < ((size_t)1 << 33) ? fb * 1024 : fb * 64);
    size_t nbytes = 0; uint64_t hash = 0; int frames = 0; double best = 1e30;
    for (int r = 0; r < repeat; r++) {
        double s = 0;
        if (mode == "serial") {
            auto t0 = std::chrono::steady_clock::now();
            frames = ref_decode_serial(bs.data(), len, w, h, cf, outp ? out.data() : nullptr, out.size(), &nbytes, &hash);
            s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        } else {
            frames = ref_decode_mt(bs.data(), len, w, h, cf, pool, threads, (outp || hash_only) ? 1 : 0, outp ? out.data() : nullptr, out.size(), &nbytes, &hash, &s);
        }
        if (s < best) best = s;
    }
    if (outp) { FILE* f = fopen(outp, "wb"); if (f) { fwrite(out.data(), 1, nbytes < out.size() ? nbytes : out.size(), f); fclose(f); } }
    printf("{\"mode\":\"%s\",\"frames\":%d,\"bytes\":%zu,\"hash\":\"%016llx\",\"seconds\":%.6f,\"fps\":%.2f}\n",
           mode.c_str(), frames, nbytes, (unsigned long long)hash, best, frames / best);
    return 0;
}
#endif
