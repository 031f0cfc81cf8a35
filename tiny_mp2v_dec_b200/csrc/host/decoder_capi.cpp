// Plain-C wrappers over mp2v_decoder_c and the host parser -- see include/mp2v_decode_c.h.
#include "mp2v_decode_c.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mp2v_decoder.hpp"
#include "mp2v_parser.h"
#include "stream_index.h"

using namespace mp2v;

static void set_err(char* err, size_t n, const std::string& s) {
    if (err && n) { snprintf(err, n, "%s", s.c_str()); }
}

// a decoder that outlives one call: device contexts are allocated at create (as the reference
// allocates its frame pool in the constructor, decoder.cpp:381-406) and reused by every decode
struct mp2v_decoder {
    mp2v_decode_params_t p{};
    mp2v_decoder_c dec;
    // per-call sink state, read by the renderer closure
    mp2v_frame_fn fn = nullptr;
    void* user = nullptr;
    uint8_t* out = nullptr;
    size_t out_cap = 0, pos = 0;
    uint64_t hash = 0, frames = 0;

    void render(frame_c* f) {
        frames++;
        if (!p.download_frames) return;
        int32_t strides[3], widths[3], heights[3];
        uint8_t* planes[3];
        for (int i = 0; i < 3; i++) { planes[i] = f->get_planes(i); strides[i] = f->get_strides(i); widths[i] = f->get_width(i); heights[i] = f->get_height(i); }
        if (fn) fn(user, planes, strides, widths, heights);
        if (!out && !p.hash_output) return;
        for (int i = 0; i < 3; i++) {
            const uint8_t* row = planes[i];
            for (int y = 0; y < heights[i]; y++, row += strides[i]) {
                if (out && pos + (size_t)widths[i] <= out_cap) memcpy(out + pos, row, (size_t)widths[i]);
                if (p.hash_output) for (int x = 0; x < widths[i]; x++) { hash ^= row[x]; hash *= 1099511628211ull; }
                pos += (size_t)widths[i];
            }
        }
    }
};

extern "C" MP2V_API int mp2v_decoder_create(const mp2v_decode_params_t* p, mp2v_decoder_t** out, char* err, size_t err_len) {
    if (!p || !out) return MP2V_ERR_ARG;
    std::unique_ptr<mp2v_decoder> d(new mp2v_decoder);
    d->p = *p;
    mp2v_decoder* raw = d.get();
    decoder_config_t cfg = {p->width, p->height, p->chroma_format, p->pictures_pool_size > 0 ? p->pictures_pool_size : 10,
                            p->num_threads > 0 ? p->num_threads : 1, p->reordering != 0};
    if (!d->dec.decoder_init(cfg, [raw](frame_c* f) { raw->render(f); })) { set_err(err, err_len, d->dec.last_error()); return MP2V_ERR_ARG; }
    mp2v_b200_options_t opt;
    if (p->n_devices > 0) opt.devices.assign(p->devices, p->devices + (p->n_devices > 8 ? 8 : p->n_devices));
    if (p->max_batch > 0) opt.max_batch = p->max_batch;
    if (p->output_lag > 0) opt.output_lag = p->output_lag;
    opt.download_frames = p->download_frames != 0;
    opt.gpu_vlc = p->host_parser == 0;
    d->dec.set_options(opt);
    if (!d->dec.prepare()) { set_err(err, err_len, d->dec.last_error()); return MP2V_ERR_CUDA; }
    *out = d.release();
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_decoder_set_device_renderer(mp2v_decoder_t* d, mp2v_device_frame_fn fn, void* user) {
    if (!d) return MP2V_ERR_ARG;
    std::function<void(const mp2v_device_frame_t&)> r;
    if (fn) r = [fn, user](const mp2v_device_frame_t& f) {
        const int32_t st[3] = {f.strides[0], f.strides[1], f.strides[2]}, w[3] = {f.width[0], f.width[1], f.width[2]}, h[3] = {f.height[0], f.height[1], f.height[2]};
        fn(user, f.planes, st, w, h, f.device, f.frame_id, f.recon);
    };
    d->dec.set_device_renderer(r);
    return MP2V_OK;
}

extern "C" MP2V_API void mp2v_decoder_destroy(mp2v_decoder_t* d) { delete d; }

static int run_decode(mp2v_decoder_t* d, uint8_t* buffer, int len, bool resident, mp2v_frame_fn fn, void* user,
                      uint8_t* out, size_t out_cap, size_t* out_bytes, mp2v_decode_stats_t* stats, char* err, size_t err_len) {
    d->fn = fn; d->user = user; d->out = out; d->out_cap = out_cap; d->pos = 0;
    d->hash = 1469598103934665603ull; d->frames = 0;
    const bool ok = resident ? d->dec.decode_resident() : d->dec.decode(buffer, len);
    if (out_bytes) *out_bytes = d->pos;
    if (stats) {
        const auto s = d->dec.stats();
        stats->frames = d->frames; stats->pictures = s.pictures; stats->launches = s.launches;
        stats->h2d_bytes = s.h2d_bytes; stats->d2h_bytes = s.d2h_bytes; stats->algorithmic_bytes = s.algorithmic_bytes;
        stats->kernel_ms = s.kernel_ms; stats->parse_cpu_seconds = s.parse_cpu_seconds; stats->wall_seconds = s.wall_seconds;
        stats->hash = d->hash;
        stats->vlc_launches = s.vlc_launches;
        stats->device_ms = s.device_ms;
    }
    if (!ok) {
        set_err(err, err_len, d->dec.last_error());
        return (strstr(d->dec.last_error(), "CUDA") || strstr(d->dec.last_error(), "recon_create")) ? MP2V_ERR_CUDA : MP2V_ERR_RANGE;
    }
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_decoder_decode(mp2v_decoder_t* d, uint8_t* buffer, int len, mp2v_frame_fn fn, void* user,
                                            uint8_t* out, size_t out_cap, size_t* out_bytes, mp2v_decode_stats_t* stats,
                                            char* err, size_t err_len) {
    if (!d || !buffer || len < 0) return MP2V_ERR_ARG;
    return run_decode(d, buffer, len, false, fn, user, out, out_cap, out_bytes, stats, err, err_len);
}

extern "C" MP2V_API int mp2v_decoder_decode_resident(mp2v_decoder_t* d, mp2v_frame_fn fn, void* user, uint8_t* out, size_t out_cap, size_t* out_bytes,
                                                     mp2v_decode_stats_t* stats, char* err, size_t err_len) {
    if (!d) return MP2V_ERR_ARG;
    return run_decode(d, nullptr, 0, true, fn, user, out, out_cap, out_bytes, stats, err, err_len);
}

extern "C" MP2V_API int mp2v_decode_stream(const mp2v_decode_params_t* p, uint8_t* buffer, int len,
                                           mp2v_frame_fn fn, void* user, uint8_t* out, size_t out_cap, size_t* out_bytes,
                                           mp2v_decode_stats_t* stats, char* err, size_t err_len) {
    mp2v_decoder_t* d = nullptr;
    int rc = mp2v_decoder_create(p, &d, err, err_len);
    if (rc != MP2V_OK) return rc;
    rc = mp2v_decoder_decode(d, buffer, len, fn, user, out, out_cap, out_bytes, stats, err, err_len);
    mp2v_decoder_destroy(d);
    return rc;
}

// ------------------------------------------------------------------------------------------------ host-only parse

struct mp2v_parsed {
    stream_index_t index;
    struct pic_t {
        mp2v_pic_params_t params{};
        std::vector<mp2v_mb_info_t> mb;
        std::unique_ptr<mp2v_coef_t[]> coef;
        coef_arena_t arena;
    };
    std::vector<std::unique_ptr<pic_t>> pics;
    double wall = 0, cpu = 0;
};

extern "C" MP2V_API int mp2v_parse_stream(const uint8_t* buffer, int len, int width, int height, int cf, int threads,
                                          mp2v_parsed_t** out, char* err, size_t err_len) {
    if (!buffer || !out || width <= 0 || height <= 0 || (width & 15) || (height & 15) || cf < 1 || cf > 3) return MP2V_ERR_ARG;
    std::unique_ptr<mp2v_parsed> P(new mp2v_parsed);
    if (!index_stream(buffer, (size_t)len, P->index, threads)) { set_err(err, err_len, P->index.error); return MP2V_ERR_RANGE; }
    const int mbw = width / 16, mbh = height / 16, nblk = cf == 1 ? 6 : cf == 2 ? 8 : 12;
    struct job_t { int pic, slice; };
    std::vector<job_t> jobs;
    for (size_t i = 0; i < P->index.pictures.size(); i++) {
        const coded_picture_t& src = P->index.pictures[i];
        if (src.seq.chroma_format != cf) { set_err(err, err_len, "stream chroma_format differs from the requested one"); return MP2V_ERR_ARG; }
        P->pics.emplace_back(new mp2v_parsed::pic_t);
        auto& pic = *P->pics.back();
        const mp2v_mb_info_t blank = {0u, MP2V_MB_BITS(0, 1, 0, MP2V_MB_INTRA), {{0, 0}, {0, 0}}};
        pic.mb.assign((size_t)mbw * mbh, blank);
        // bits bound the records: a coefficient costs >= 2 bits + sign; plus per-row chunk slack
        size_t bytes = 0;
        for (size_t s = 0; s < src.slices.size(); s++) {
            const uint8_t* a = src.slices[s].payload;
            const uint8_t* b = find_start_code(a, buffer + len);
            bytes += (size_t)(b - a);
        }
        // bits bound the records: a coefficient costs >= 3 bits, an intra DC + end of block >= 4
        uint64_t cap = bytes * 8 / 3 + (uint64_t)mbw * mbh * nblk;
        const uint64_t worst = (uint64_t)mbw * mbh * nblk * 64u;
        if (cap > worst) cap = worst;
        pic.coef.reset(new mp2v_coef_t[cap]);
        pic.arena.base = pic.coef.get();
        pic.arena.capacity = (uint32_t)cap;
        build_picture_matrices(src.info, pic.params.W);
        pic.params.picture_coding_type = src.info.picture_coding_type;
        pic.params.alternate_scan = src.info.alternate_scan;
        pic.params.dst_frame = (int)i; pic.params.l0_frame = pic.params.l1_frame = -1;
        for (size_t s = 0; s < src.slices.size(); s++) jobs.push_back({(int)i, (int)s});
    }
    std::atomic<size_t> next{0};
    std::atomic<int64_t> cpu_ns{0};
    std::atomic<bool> bad{false};
    std::string first_error;
    std::mutex emu;
    auto work = [&] {
        for (;;) {
            const size_t j = next.fetch_add(1);
            if (j >= jobs.size()) return;
            const coded_picture_t& src = P->index.pictures[jobs[j].pic];
            auto& pic = *P->pics[jobs[j].pic];
            const auto t0 = std::chrono::steady_clock::now();
            const slice_ref_t& sr = src.slices[jobs[j].slice];
            const slice_result_t r = parse_slice(sr.payload, sr.code, src.seq, src.info, mbw, mbh, pic.mb.data(), pic.arena);
            cpu_ns.fetch_add(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count());
            if (!r.ok) { bad.store(true); std::lock_guard<std::mutex> lk(emu); if (first_error.empty()) first_error = r.error ? r.error : "slice parse error"; }
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int i = 1; i < (threads < 1 ? 1 : threads); i++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    P->wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    P->cpu = cpu_ns.load() * 1e-9;
    if (bad.load()) { set_err(err, err_len, first_error); return MP2V_ERR_RANGE; }
    // reference bookkeeping in coded order (decoder.cpp:294-305), as coded indices
    int refs[2] = {-1, -1};
    for (size_t i = 0; i < P->pics.size(); i++) {
        auto& pp = P->pics[i]->params;
        pp.n_coef = P->pics[i]->arena.next.load();
        if (pp.picture_coding_type == 3) { pp.l0_frame = refs[0]; pp.l1_frame = refs[1]; }
        else { pp.l0_frame = pp.picture_coding_type == 2 ? refs[1] : -1; refs[0] = refs[1]; refs[1] = (int)i; }
    }
    *out = P.release();
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_parsed_num_pictures(const mp2v_parsed_t* p) { return p ? (int)p->pics.size() : 0; }

extern "C" MP2V_API int mp2v_parsed_picture(const mp2v_parsed_t* p, int i, mp2v_pic_params_t* params, const mp2v_mb_info_t** mb,
                                            const mp2v_coef_t** coef, uint32_t* n_coef, int32_t* temporal_reference, int32_t* gop) {
    if (!p || i < 0 || i >= (int)p->pics.size()) return MP2V_ERR_ARG;
    const auto& pic = *p->pics[i];
    if (params) *params = pic.params;
    if (mb) *mb = pic.mb.data();
    if (coef) *coef = pic.coef.get();
    if (n_coef) *n_coef = pic.params.n_coef;
    if (temporal_reference) *temporal_reference = p->index.pictures[i].info.temporal_reference;
    if (gop) *gop = p->index.pictures[i].gop;
    return MP2V_OK;
}

extern "C" MP2V_API double mp2v_parsed_wall_seconds(const mp2v_parsed_t* p) { return p ? p->wall : 0; }
extern "C" MP2V_API double mp2v_parsed_cpu_seconds(const mp2v_parsed_t* p) { return p ? p->cpu : 0; }
extern "C" MP2V_API void mp2v_parsed_free(mp2v_parsed_t* p) { delete p; }
