// mp2v_decoder_c on the B200 back end -- see include/mp2v_decoder.hpp.
//
// Reference roles and where they went (src/core/decoder.cpp, threads.cpp):
//   decode()'s start-code switch          -> index_stream()                       (caller's thread)
//   task_queue_c ring + busy-spin workers  -> slice queue with condition variables (num_threads workers)
//   mp2v_picture_c::decode_slice           -> parse_slice(): emits records, no pixel work
//   picture dependencies (add_dependency)  -> coded-order submission to mp2v_recon; the device stream is the order
//   decoder_output_scheduler               -> per-device output thread, same I/P/B display reorder
//   frame_c pool                           -> device frame pool (ids) + pinned host mirrors for the renderer
// Closed GOPs are independent chains: with several devices they are dealt round-robin to one
// pipeline per GPU and stitched back in display order; nothing is exchanged between GPUs.
#include "mp2v_decoder.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>

#include "mp2v_parser.h"
#include "numa.h"
#include "mp2v_recon.h"
#include "stream_index.h"

using namespace mp2v;

// ------------------------------------------------------------------------------------------------ frame_c

static void frame_geometry(int width, int height, int cf, uint32_t w[3], uint32_t h[3], uint32_t s[3]) {
    mp2v_frame_layout_t lay;
    if (mp2v_frame_layout(width, height, cf, &lay) != MP2V_OK) { for (int p = 0; p < 3; p++) w[p] = h[p] = s[p] = 0; return; }
    for (int p = 0; p < 3; p++) { w[p] = (uint32_t)lay.width[p]; h[p] = (uint32_t)lay.height[p]; s[p] = (uint32_t)lay.stride[p]; }
}

frame_c::frame_c(int width, int height, int chroma_format) {
    frame_geometry(width, height, chroma_format, m_width, m_height, m_stride);
    for (int p = 0; p < 3; p++) m_planes[p] = (uint8_t*)aligned_alloc(64, ((size_t)m_height[p] * m_stride[p] + 63) & ~(size_t)63);
    m_owner = true;
}

frame_c::frame_c(int width, int height, int chroma_format, uint8_t* const planes[3], const int strides[3]) {
    frame_geometry(width, height, chroma_format, m_width, m_height, m_stride);
    for (int p = 0; p < 3; p++) { m_planes[p] = planes[p]; m_stride[p] = (uint32_t)strides[p]; }
}

frame_c::~frame_c() {
    if (m_owner) for (auto* p : m_planes) free(p);
}

// ------------------------------------------------------------------------------------------------ internals

namespace {

using clock_t_ = std::chrono::steady_clock;

struct pipeline_t;

struct pic_task_t {
    const coded_picture_t* src = nullptr;
    pipeline_t* pipe = nullptr;
    int local_index = 0;               // position in the pipeline's coded order
    int dst = -1, l0 = -1, l1 = -1;    // device frame ids
    mp2v_picture_t* rp = nullptr;
    coef_arena_t arena;
    int n_jobs = 0;                    // worker jobs of this picture: its slices (host parser) or 1 (stage for the device parser)
    std::atomic<int> next_slice{0};    // next job index a worker may claim
    std::atomic<int> remaining{0};
    std::atomic<bool> ok{true};
    const char* error = nullptr;
    std::string error_text;            // storage for messages that are not literals
    double ts_queued = 0, ts_staged = 0, ts_submitted = 0, ts_mapped = 0;   // MP2V_PROFILE=2: milliseconds since decode() began
    bool parsed = false, submitted = false;
};

struct shared_t {
    decoder_config_t cfg{};
    mp2v_b200_options_t opt;
    std::function<void(frame_c*)> renderer;
    int mbw = 0, mbh = 0;
    bool stream_mode = false;                // the stream is resident on the device(s): pictures are handed over as slice offsets
    const uint8_t* stream_base = nullptr;
    // work distribution: pictures whose slices may be claimed (an atomic index per picture); workers
    // only take the lock to move on to the next picture or to sleep
    std::mutex qmu;
    std::condition_variable qcv;
    std::deque<pic_task_t*> queue;
    bool stop = false;
    // errors
    std::atomic<bool> failed{false};
    std::mutex emu;
    std::string error;
    // display order across GOP chains
    std::mutex gmu;
    std::condition_variable gcv;
    int emit_gop = 0;
    std::vector<int> gop_size, gop_emitted;
    // statistics (nanoseconds, summed over threads)
    std::atomic<int64_t> parse_ns{0};
    std::atomic<int64_t> feeder_wait_ns{0}, feeder_work_ns{0}, submit_ns{0}, precheck_ns{0}, out_wait_ns{0}, out_map_ns{0}, worker_idle_ns{0};
    std::vector<pipeline_t*> pipes;
    clock_t_::time_point t_origin;
    double now_ms() const { return std::chrono::duration<double, std::milli>(clock_t_::now() - t_origin).count(); }

    void fail(const std::string& why);
};

struct pipeline_t {
    shared_t* sh = nullptr;
    int device = 0;
    mp2v_recon_t* recon = nullptr;
    std::deque<pic_task_t> tasks;            // coded order, this device's GOP chains only
    // frame pool: use count = 1 while awaiting display + 1 while it is one of the two live references
    std::vector<int> frame_use;
    int n_frames = 0, n_slots = 0, parse_window = 0;
    std::mutex mu;
    std::condition_variable cv;
    int in_parse = 0;                        // acquired, not yet submitted
    int next_submit = 0;
    int frame_cursor = 0;                    // next-fit: consecutive pictures get consecutive frame ids, so their copies to the host merge
    std::thread feeder_thread, output_thread;

    void feeder();
    void output();
    void on_parsed(pic_task_t* t);
    void release_frame_use(int f) {
        std::lock_guard<std::mutex> lk(mu);
        if (f >= 0 && --frame_use[f] == 0) cv.notify_all();
    }
    void wake() { { std::lock_guard<std::mutex> lk(mu); } cv.notify_all(); }
};

void shared_t::fail(const std::string& why) {
    { std::lock_guard<std::mutex> lk(emu); if (error.empty()) error = why; }
    failed.store(true);
    { std::lock_guard<std::mutex> lk(qmu); }
    qcv.notify_all();
    { std::lock_guard<std::mutex> lk(gmu); }
    gcv.notify_all();
    for (auto* p : pipes) p->cv.notify_all();   // waiters re-check `failed` (they hold p->mu only while testing)
}

void pipeline_t::feeder() {
    bind_this_thread_to_numa_node(mp2v_recon_numa_node(recon));      // next to the device's pinned memory (no-op on single-node hosts)
    int refs[2] = {-1, -1};   // local task indices of the two live references
    std::vector<uint32_t> slice_offsets;
    auto unref = [&](int k) { if (k >= 0) release_frame_use(tasks[k].dst); };
    for (size_t k = 0; k < tasks.size() && !sh->failed.load(); k++) {
        pic_task_t& t = tasks[k];
        const picture_info_t& info = t.src->info;
        const auto tf0 = clock_t_::now();
        {   // a free device frame and room in the parse window
            std::unique_lock<std::mutex> lk(mu);
            int f = -1;
            cv.wait(lk, [&] {
                if (sh->failed.load()) return true;
                if (in_parse >= parse_window) return false;
                for (int i = 0; i < n_frames; i++) { const int c = (frame_cursor + i) % n_frames; if (frame_use[c] == 0) { f = c; return true; } }
                return false;
            });
            if (sh->failed.load()) break;
            t.dst = f;
            frame_cursor = (f + 1) % n_frames;
            frame_use[f] = 1;               // awaiting display
            in_parse++;
        }
        const auto tf1 = clock_t_::now();
        sh->feeder_wait_ns.fetch_add(std::chrono::duration_cast<std::chrono::nanoseconds>(tf1 - tf0).count(), std::memory_order_relaxed);
        if (info.picture_coding_type == 3) {                       // B: both references (decoder.cpp:303)
            t.l0 = refs[0] >= 0 ? tasks[refs[0]].dst : -1;
            t.l1 = refs[1] >= 0 ? tasks[refs[1]].dst : -1;
            if (t.l0 < 0 || t.l1 < 0) { sh->fail("B picture without two reference pictures"); break; }
        } else {                                                   // I/P: previous reference (decoder.cpp:298-302)
            t.l0 = (info.picture_coding_type == 2 && refs[1] >= 0) ? tasks[refs[1]].dst : -1;
            if (info.picture_coding_type == 2 && t.l0 < 0) { sh->fail("P picture without a reference picture"); break; }
            unref(refs[0]);
            refs[0] = refs[1];
            refs[1] = (int)k;
            { std::lock_guard<std::mutex> lk(mu); frame_use[t.dst]++; }
        }
        if (mp2v_recon_acquire_picture(recon, &t.rp) != MP2V_OK) { sh->fail(std::string("acquire_picture: ") + mp2v_recon_last_error(recon)); break; }
        mp2v_pic_params_t& pp = *t.rp->params;
        build_picture_matrices(info, pp.W);
        pp.picture_coding_type = info.picture_coding_type;
        pp.alternate_scan = info.alternate_scan;
        pp.dst_frame = t.dst; pp.l0_frame = t.l0; pp.l1_frame = t.l1;
        if (sh->stream_mode) {
            // stream-resident device parsing: the coded bytes are already on the device; the picture is its slices' offsets
            if (!picture_in_envelope(info)) { sh->fail("only frame pictures without concealment vectors are supported (the reference's envelope)"); break; }
            mp2v_pic_syntax_t sy{};
            for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) sy.f_code[a][b] = info.f_code[a][b];
            sy.intra_dc_precision = info.intra_dc_precision;
            sy.q_scale_type = info.q_scale_type;
            sy.intra_vlc_format = info.intra_vlc_format;
            sy.field_dct_syntax = info.frame_pred_frame_dct ? 0 : 1;
            slice_offsets.clear();
            for (const slice_ref_t& sr : t.src->slices) slice_offsets.push_back((uint32_t)(sr.payload - 4 - sh->stream_base));
            if (mp2v_recon_submit_stream_picture(recon, t.rp, &sy, slice_offsets.data(), (int)slice_offsets.size()) != MP2V_OK) {
                sh->fail(std::string("picture ") + std::to_string(k) + ": " + mp2v_recon_last_error(recon));
                break;
            }
            t.ts_queued = t.ts_submitted = sh->now_ms();
            {
                std::lock_guard<std::mutex> lk(mu);
                t.parsed = t.submitted = true;
                next_submit = (int)k + 1;
                in_parse--;
            }
            cv.notify_all();
            sh->feeder_work_ns.fetch_add(std::chrono::duration_cast<std::chrono::nanoseconds>(clock_t_::now() - tf1).count(), std::memory_order_relaxed);
            continue;
        }
        if (sh->opt.gpu_vlc) {
            // device-side parsing: a worker stages the coded slices (one job per picture), submission stays in coded order
            if (!picture_in_envelope(info)) { sh->fail("only frame pictures without concealment vectors are supported (the reference's envelope)"); break; }
            t.n_jobs = 1;
            t.ts_queued = sh->now_ms();
            t.remaining.store(1);
            t.next_slice.store(0);
            {
                std::lock_guard<std::mutex> lk(sh->qmu);
                sh->queue.push_back(&t);
            }
            sh->qcv.notify_one();
            sh->feeder_work_ns.fetch_add(std::chrono::duration_cast<std::chrono::nanoseconds>(clock_t_::now() - tf1).count(), std::memory_order_relaxed);
            continue;
        }
        // macroblocks no slice covers: intra with no coded block (reconstructs to 0); never the case in valid streams
        const mp2v_mb_info_t blank = {0u, MP2V_MB_BITS(0, 1, 0, MP2V_MB_INTRA), {{0, 0}, {0, 0}}};
        for (uint32_t i = 0; i < t.rp->mb_count; i++) t.rp->mb[i] = blank;
        t.arena.base = t.rp->coef;
        t.arena.capacity = t.rp->coef_capacity;
        t.arena.next.store(0);
        t.arena.overflow.store(false);
        const int ns = (int)t.src->slices.size();
        if (ns == 0) { on_parsed(&t); continue; }
        t.n_jobs = ns;
        t.remaining.store(ns);
        t.next_slice.store(0);
        {
            std::lock_guard<std::mutex> lk(sh->qmu);
            sh->queue.push_back(&t);
        }
        sh->qcv.notify_all();
        sh->feeder_work_ns.fetch_add(std::chrono::duration_cast<std::chrono::nanoseconds>(clock_t_::now() - tf1).count(), std::memory_order_relaxed);
    }
    unref(refs[0]);
    unref(refs[1]);
    // the last pictures rarely fill a launch lot: hand them over now instead of when the output thread asks for them
    if (sh->stream_mode && !sh->failed.load() && mp2v_recon_flush(recon) != MP2V_OK) sh->fail(std::string("flush: ") + mp2v_recon_last_error(recon));
}

// called by the worker that finished the last slice of a picture
void pipeline_t::on_parsed(pic_task_t* t) {
    // record validation / byte accounting is per picture: do it before taking the submission lock
    const auto ts0 = clock_t_::now();
    if (!sh->opt.gpu_vlc && t->ok.load() && !t->arena.overflow.load() && t->rp) {
        t->rp->params->n_coef = t->arena.next.load();
        if (mp2v_recon_precheck(recon, t->rp) != MP2V_OK) { t->error = "records failed validation (motion vector outside the frame or bad offsets)"; t->ok.store(false); }
    }
    const auto ts1 = clock_t_::now();
    sh->precheck_ns.fetch_add(std::chrono::duration_cast<std::chrono::nanoseconds>(ts1 - ts0).count(), std::memory_order_relaxed);
    struct submit_timer_t { shared_t* sh; clock_t_::time_point t0; ~submit_timer_t() { sh->submit_ns.fetch_add(std::chrono::duration_cast<std::chrono::nanoseconds>(clock_t_::now() - t0).count(), std::memory_order_relaxed); } } submit_timer{sh, ts1};
    std::lock_guard<std::mutex> lk(mu);
    t->parsed = true;
    while (next_submit < (int)tasks.size() && tasks[next_submit].parsed && !sh->failed.load()) {
        pic_task_t& s = tasks[next_submit];
        if (!s.ok.load() || s.arena.overflow.load()) {
            sh->fail(std::string("picture ") + std::to_string(next_submit) + ": " + (s.error ? s.error : "coefficient arena exhausted"));
            break;
        }
        if (sh->opt.gpu_vlc) {
            if (mp2v_recon_submit_staged(recon, s.rp) != MP2V_OK) { sh->fail(std::string("picture ") + std::to_string(next_submit) + ": " + mp2v_recon_last_error(recon)); break; }
        } else {
            s.rp->params->n_coef = s.arena.next.load();
            if (mp2v_recon_submit(recon, s.rp) != MP2V_OK) { sh->fail(std::string("submit: ") + mp2v_recon_last_error(recon)); break; }
        }
        s.submitted = true;
        s.ts_submitted = sh->now_ms();
        next_submit++;
        in_parse--;
    }
    cv.notify_all();
}

void pipeline_t::output() {
    bind_this_thread_to_numa_node(mp2v_recon_numa_node(recon));
    const int n = (int)tasks.size();
    const int lag = sh->opt.output_lag < 0 ? 0 : sh->opt.output_lag;
    auto emit = [&](int k) -> bool {
        pic_task_t& t = tasks[k];
        const auto to0 = clock_t_::now();
        {   // stay `lag` pictures behind the submit side so that launches can batch
            std::unique_lock<std::mutex> lk(mu);
            const int need = (k + lag < n - 1 ? k + lag : n - 1) + 1;
            cv.wait(lk, [&] { return sh->failed.load() || next_submit >= need; });
            if (sh->failed.load()) return false;
        }
        uint8_t* planes[3] = {nullptr, nullptr, nullptr};
        int32_t strides[3] = {0, 0, 0};
        const auto to1 = clock_t_::now();
        sh->out_wait_ns.fetch_add(std::chrono::duration_cast<std::chrono::nanoseconds>(to1 - to0).count(), std::memory_order_relaxed);
        if (sh->opt.download_frames) {
            if (mp2v_recon_map_frame(recon, t.dst, planes, strides) != MP2V_OK) { sh->fail(std::string("map_frame: ") + mp2v_recon_last_error(recon)); return false; }
        }
        sh->out_map_ns.fetch_add(std::chrono::duration_cast<std::chrono::nanoseconds>(clock_t_::now() - to1).count(), std::memory_order_relaxed);
        t.ts_mapped = sh->now_ms();
        {   // display order across GOP chains: chain g is shown after chain g-1
            std::unique_lock<std::mutex> lk(sh->gmu);
            sh->gcv.wait(lk, [&] { return sh->failed.load() || sh->emit_gop == t.src->gop; });
            if (sh->failed.load()) return false;
        }
        if (sh->opt.device_renderer) {
            mp2v_device_frame_t df{};
            int32_t dstr[3] = {0, 0, 0};
            if (mp2v_recon_wait_frame(recon, t.dst) != MP2V_OK || mp2v_recon_frame_device_ptrs(recon, t.dst, df.planes, dstr) != MP2V_OK) {
                sh->fail(std::string("device frame: ") + mp2v_recon_last_error(recon));
                return false;
            }
            mp2v_frame_layout_t lay;
            mp2v_frame_layout(sh->cfg.width, sh->cfg.height, sh->cfg.chroma_format, &lay);
            for (int p = 0; p < 3; p++) { df.strides[p] = dstr[p]; df.width[p] = lay.width[p]; df.height[p] = lay.height[p]; }
            df.device = device; df.frame_id = t.dst; df.recon = recon;
            sh->opt.device_renderer(df);
        }
        if (sh->renderer) {
            const int st[3] = {strides[0], strides[1], strides[2]};
            frame_c view(sh->cfg.width, sh->cfg.height, sh->cfg.chroma_format, planes, st);
            sh->renderer(&view);
        }
        {
            std::lock_guard<std::mutex> lk(sh->gmu);
            if (++sh->gop_emitted[t.src->gop] == sh->gop_size[t.src->gop]) { sh->emit_gop++; sh->gcv.notify_all(); }
        }
        release_frame_use(t.dst);
        return true;
    };
    // display reorder of decoder_output_scheduler (decoder.cpp:350-378)
    int held = -1;
    for (int k = 0; k < n; k++) {
        // end of a GOP chain: its held reference is the last frame of that chain
        if (k > 0 && held >= 0 && tasks[k].src->gop != tasks[k - 1].src->gop) { if (!emit(held)) return; held = -1; }
        const bool is_b = tasks[k].src->info.picture_coding_type == 3;
        if (is_b || !sh->cfg.reordering) { if (!emit(k)) return; }
        else {
            if (held >= 0 && !emit(held)) return;
            held = k;
        }
    }
    if (held >= 0) emit(held);
}

void worker_main(shared_t* sh) {
    std::vector<mp2v_slice_ref_t> slice_refs;
    for (;;) {
        pic_task_t* t = nullptr;
        int slice = -1;
        {
            std::unique_lock<std::mutex> lk(sh->qmu);
            for (;;) {
                // drop pictures whose slices have all been claimed, claim one from the first that has any
                while (!sh->queue.empty() && sh->queue.front()->next_slice.load(std::memory_order_relaxed) >= sh->queue.front()->n_jobs)
                    sh->queue.pop_front();
                if (!sh->queue.empty()) { t = sh->queue.front(); break; }
                if (sh->stop) return;
                const auto ti0 = clock_t_::now();
                sh->qcv.wait(lk);
                sh->worker_idle_ns.fetch_add(std::chrono::duration_cast<std::chrono::nanoseconds>(clock_t_::now() - ti0).count(), std::memory_order_relaxed);
            }
        }
        // claim slices of this picture without the lock until it runs dry
        const int ns = t->n_jobs;
        while ((slice = t->next_slice.fetch_add(1, std::memory_order_relaxed)) < ns) {
            if (sh->opt.gpu_vlc) {
                if (!sh->failed.load(std::memory_order_relaxed)) {
                    // stage the coded picture for the device parser: bytes into the slot's pinned buffer, no device work
                    const picture_info_t& info = t->src->info;
                    mp2v_pic_syntax_t sy{};
                    for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) sy.f_code[a][b] = info.f_code[a][b];
                    sy.intra_dc_precision = info.intra_dc_precision;
                    sy.q_scale_type = info.q_scale_type;
                    sy.intra_vlc_format = info.intra_vlc_format;
                    sy.field_dct_syntax = info.frame_pred_frame_dct ? 0 : 1;
                    std::vector<mp2v_slice_ref_t>& refs = slice_refs;
                    refs.clear();
                    for (const slice_ref_t& sr : t->src->slices) refs.push_back({sr.payload, sr.bytes, sr.code});
                    const int stage_rc = mp2v_recon_stage_slices(t->pipe->recon, t->rp, &sy, refs.data(), (int)refs.size());
                    t->ts_staged = sh->now_ms();
                    if (stage_rc != MP2V_OK) {
                        t->error_text = mp2v_recon_last_error(t->pipe->recon);
                        t->error = t->error_text.c_str();
                        t->ok.store(false);
                    }
                }
            } else if (!sh->failed.load(std::memory_order_relaxed)) {
                const auto t0 = clock_t_::now();
                const slice_ref_t& sr = t->src->slices[slice];
                const slice_result_t r = parse_slice(sr.payload, sr.code, t->src->seq, t->src->info, sh->mbw, sh->mbh, t->rp->mb, t->arena);
                sh->parse_ns.fetch_add(std::chrono::duration_cast<std::chrono::nanoseconds>(clock_t_::now() - t0).count(), std::memory_order_relaxed);
                if (!r.ok) { t->error = r.error; t->ok.store(false); }
            }
            if (t->remaining.fetch_sub(1) == 1) { t->pipe->on_parsed(t); break; }   // t may be recycled after this
        }
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ mp2v_decoder_c

// per-device reconstruction context, created once (prepare) and reused by every decode() call
struct device_ctx_t {
    int device = 0;
    mp2v_recon_t* recon = nullptr;
    int n_frames = 0, n_slots = 0;
};

struct mp2v_decoder_c::impl_t {
    decoder_config_t cfg{};
    mp2v_b200_options_t opt;
    std::function<void(frame_c*)> renderer;
    bool initialised = false;
    std::string error;
    stats_t stats;
    // [0] contexts fed by the host slice parser, [1] contexts that parse on the device; each set is
    // created when first needed (a stream outside the device parser's envelope takes the host parser)
    std::vector<device_ctx_t> dev_sets[2];

    void release() {
        for (auto& devs : dev_sets) {
            for (auto& d : devs) if (d.recon) mp2v_recon_destroy(d.recon);
            devs.clear();
        }
    }
    bool prepare(bool gpu_vlc);
    bool device_parser_can_take(const stream_index_t& index, bool staged) const;

    // the stream of the last decode(): its index, and whether it is resident on the device(s) (decode_resident re-decodes it)
    struct stream_state_t {
        stream_index_t index;
        uint8_t* buffer = nullptr;
        int len = 0;
        bool resident = false;        // uploaded to and scanned on the first device
        bool gpu_vlc = false;         // slices are parsed on the device
        bool stream_mode = false;     // ... straight out of the resident copy
        double index_ms = 0;
    };
    std::unique_ptr<stream_state_t> last;
    bool open_stream(mp2v_decoder_c& owner, uint8_t* buffer, int len);
    bool run(mp2v_decoder_c& owner, clock_t_::time_point t_begin, bool fresh_stats);
};

// mp2v_recon_submit_slices' envelope: at most one slice per macroblock row, coded picture within the staging capacity
bool mp2v_decoder_c::impl_t::device_parser_can_take(const stream_index_t& index, bool staged) const {
    const int mbh = cfg.height / 16;
    const size_t cap = std::max<size_t>(2u << 20, (size_t)(cfg.width / 16) * mbh * 128u);   // the library's default bitstream_capacity
    std::vector<uint8_t> seen((size_t)mbh);
    for (const auto& pic : index.pictures) {
        if ((int)pic.slices.size() > mbh) return false;
        if (pic.slices.empty()) continue;
        std::fill(seen.begin(), seen.end(), 0);
        const uint8_t* lo = pic.slices[0].payload;
        const uint8_t* hi = lo;
        for (const auto& sl : pic.slices) {
            int row = sl.code - 1;
            if (cfg.height > 2800 && sl.bytes > 0) row += (sl.payload[0] >> 5) << 7;
            if (row < 0 || row >= mbh) continue;          // reported as an error by either parser
            if (seen[row]) return false;
            seen[row] = 1;
            lo = std::min(lo, sl.payload);
            hi = std::max(hi, sl.payload + sl.bytes);
        }
        if (staged && (size_t)(hi - lo) + 64 > cap) return false;      // (a resident stream needs no staging capacity)
    }
    return true;
}

bool mp2v_decoder_c::impl_t::prepare(bool gpu_vlc) {
    std::vector<device_ctx_t>& devs = dev_sets[gpu_vlc ? 1 : 0];
    if (!devs.empty()) return true;
    const decoder_config_t& c = cfg;
    const int mbw = c.width / 16, mbh = c.height / 16;
    const int lag = opt.output_lag < 0 ? 0 : opt.output_lag;
    const uint32_t nblk = c.chroma_format == 1 ? 6 : c.chroma_format == 2 ? 8 : 12;
    const uint64_t worst = (uint64_t)mbw * mbh * nblk * 64u;
    // a picture needs the worst case only when every coefficient of every block is coded: size the
    // slots for a quarter of it, but never below 1 Mi records (or the worst case itself if smaller)
    uint64_t cap = worst / 4 > (1u << 20) ? worst / 4 : (worst < (1u << 20) ? worst : (1u << 20));
    std::vector<int> ids = opt.devices.empty() ? std::vector<int>{0} : opt.devices;
    for (int id : ids) {
        device_ctx_t d;
        d.device = id;
        d.n_frames = (c.pictures_pool_size > 4 ? c.pictures_pool_size : 4) + lag + 2;
        d.n_slots = 2 * (opt.max_batch > 0 ? opt.max_batch : 8);
        if (d.n_slots < 6) d.n_slots = 6;
        // a slice parses at the speed of one GPU thread: throughput comes from the number of pictures in flight
        if (gpu_vlc) {
            // ... about 128 pictures of 1080p, fewer of larger ones (a slot's worst-case coefficient arena is 1.5 KiB per macroblock)
            int slots = (int)std::min<int64_t>(128, std::max<int64_t>(24, 128ll * 8160 / ((int64_t)mbw * mbh)));
            int extra_frames = slots;
            if (const char* v = getenv("MP2V_VLC_SLOTS")) slots = atoi(v) > 0 ? atoi(v) : slots;             // dev knobs
            if (const char* v = getenv("MP2V_VLC_FRAMES")) extra_frames = atoi(v) > 0 ? atoi(v) : extra_frames;
            d.n_frames += extra_frames;
            if (d.n_slots < slots || getenv("MP2V_VLC_SLOTS")) d.n_slots = slots;
        }
        mp2v_recon_config_t rc{};
        rc.device = id; rc.width = c.width; rc.height = c.height; rc.chroma_format = c.chroma_format;
        rc.n_frames = d.n_frames; rc.n_pictures = d.n_slots; rc.max_batch = opt.max_batch;
        if (gpu_vlc && rc.max_batch < 64) rc.max_batch = 64;     // device-parsed pictures are launched in lots of up to 64, one launch per dependency level
        rc.flags = MP2V_RECON_VALIDATE | (gpu_vlc ? MP2V_RECON_DEVICE_VLC : 0) | (opt.download_frames ? MP2V_RECON_AUTO_DOWNLOAD : MP2V_RECON_THROUGHPUT);
        rc.coef_capacity = (uint32_t)(cap > 0xffffffffull ? 0xffffffffull : cap);
        if (mp2v_recon_create(&rc, &d.recon) != MP2V_OK) {
            error = std::string("mp2v_recon_create (CUDA device ") + std::to_string(id) + "): " + mp2v_recon_last_error(nullptr);
            for (auto& x : devs) if (x.recon) mp2v_recon_destroy(x.recon);
            devs.clear();
            return false;
        }
        mp2v_recon_set_timing(d.recon, 1);
        devs.push_back(d);
    }
    return true;
}

mp2v_decoder_c::mp2v_decoder_c() : m(new impl_t) {
    if (const char* d = getenv("MP2V_DEVICE")) m->opt.devices = {atoi(d)};
    if (const char* v = getenv("MP2V_GPU_VLC")) m->opt.gpu_vlc = atoi(v) != 0;
}
mp2v_decoder_c::mp2v_decoder_c(const decoder_config_t& config, std::function<void(frame_c*)> renderer) : mp2v_decoder_c() {
    decoder_init(config, renderer);
}
mp2v_decoder_c::~mp2v_decoder_c() {
    m->release();
    delete m_sequence_display_extension;
    delete m_sequence_scalable_extension;
    delete m_group_of_pictures_header;
}

// the public header members (decoder.h:120-131) of the stream just indexed
static void publish_headers(mp2v_decoder_c& d, const stream_headers_t& h) {
    d.user_data = h.user_data;
    d.m_sequence_header = h.sequence_header;
    d.m_sequence_extension = h.sequence_extension;
    auto set = [](auto*& dst, bool have, const auto& src) {
        if (!have) { delete dst; dst = nullptr; return; }
        if (!dst) dst = new typename std::remove_reference<decltype(*dst)>::type;
        *dst = src;
    };
    set(d.m_sequence_display_extension, h.have_display_extension, h.sequence_display_extension);
    set(d.m_sequence_scalable_extension, h.have_scalable_extension, h.sequence_scalable_extension);
    set(d.m_group_of_pictures_header, h.have_gop_header, h.group_of_pictures_header);
}

bool mp2v_decoder_c::decoder_init(const decoder_config_t& config, std::function<void(frame_c*)> renderer) {
    m->cfg = config;
    m->renderer = renderer;
    mp2v_frame_layout_t lay;
    m->initialised = mp2v_frame_layout(config.width, config.height, config.chroma_format, &lay) == MP2V_OK && config.num_threads >= 1;
    if (!m->initialised) m->error = "bad decoder_config_t (width/height must be multiples of 16, chroma_format 1..3, num_threads >= 1)";
    return m->initialised;
}

void mp2v_decoder_c::set_options(const mp2v_b200_options_t& opt) { m->release(); m->opt = opt; }
void mp2v_decoder_c::set_device_renderer(std::function<void(const mp2v_device_frame_t&)> r) { m->opt.device_renderer = std::move(r); }
bool mp2v_decoder_c::prepare() { return m->initialised && m->prepare(m->opt.gpu_vlc); }
const char* mp2v_decoder_c::last_error() const { return m->error.c_str(); }
mp2v_decoder_c::stats_t mp2v_decoder_c::stats() const { return m->stats; }
void mp2v_decoder_c::flush(mp2v_picture_c*) {}   // decode() is one-shot and drains everything itself (as the reference's always does)

bool mp2v_decoder_c::decode(uint8_t* buffer, int len) {
    if (!m->initialised) return false;
    const auto t_begin = clock_t_::now();
    m->error.clear();
    for (auto& devs : m->dev_sets) for (auto& d : devs) { mp2v_recon_stats_t st; mp2v_recon_get_stats(d.recon, &st, 1); }   // statistics are per decode() call
    if (!m->open_stream(*this, buffer, len)) return false;
    return m->run(*this, t_begin, false);
}

bool mp2v_decoder_c::decode_resident() {
    if (!m->initialised) return false;
    m->error.clear();
    if (!m->last || !m->last->stream_mode) { m->error = "decode_resident: the last decode() did not leave a stream resident on the device"; return false; }
    return m->run(*this, clock_t_::now(), true);
}

// index the stream (on the device when slices are parsed there), publish its headers, check it against the configuration
bool mp2v_decoder_c::impl_t::open_stream(mp2v_decoder_c& owner, uint8_t* buffer, int len) {
    const auto t_begin = clock_t_::now();
    last.reset(new stream_state_t);
    stream_state_t& S = *last;
    S.buffer = buffer; S.len = len;
    stream_index_t& index = S.index;
    // Device front end: the stream goes to the device in one copy and kernels list its start codes
    // (start_codes_search.hpp:7-26); the host only parses the headers those offsets point at.  With several devices
    // (GOP sharding) every device copies and scans ONE PART of the stream, all at the same time over their own PCIe
    // links, and later receives the byte ranges of the pictures it decodes: no device ever holds the whole stream.
    std::vector<uint32_t> merged_codes;
    if (opt.gpu_vlc && len > 0) {
        if (!prepare(true)) { last.reset(); return false; }
        const std::vector<device_ctx_t>& vdevs = dev_sets[1];
        const uint32_t* codes = nullptr;
        uint32_t n_codes = 0;
        int rc = MP2V_OK;
        mp2v_recon_t* failed_on = vdevs[0].recon;
        if (vdevs.size() == 1) {
            rc = mp2v_recon_stream_begin(vdevs[0].recon, buffer, (size_t)len, nullptr, 0, 1, &codes, &n_codes);
        } else {
            const size_t nd = vdevs.size();
            const size_t part = ((((size_t)len + nd - 1) / nd) + 4095) & ~(size_t)4095;
            for (size_t d = 0; d < nd && rc == MP2V_OK; d++) {
                const size_t lo = std::min((size_t)len, d * part), hi = std::min((size_t)len, lo + part);
                const mp2v_byte_range_t r{lo, hi - lo};
                rc = mp2v_recon_stream_begin(vdevs[d].recon, buffer, (size_t)len, &r, 1, 2, nullptr, nullptr);
                if (rc != MP2V_OK) failed_on = vdevs[d].recon;
            }
            for (size_t d = 0; d < nd && rc == MP2V_OK; d++) {
                rc = mp2v_recon_stream_codes(vdevs[d].recon, &codes, &n_codes);
                if (rc != MP2V_OK) { failed_on = vdevs[d].recon; break; }
                merged_codes.insert(merged_codes.end(), codes, codes + n_codes);      // parts are in stream order: the list stays ascending
            }
            codes = merged_codes.data();
            n_codes = (uint32_t)merged_codes.size();
        }
        if (rc == MP2V_OK) {
            if (!index_stream_from_codes(buffer, (size_t)len, codes, n_codes, index)) { error = index.error; last.reset(); return false; }
            S.resident = true;
        } else if (rc != MP2V_ERR_RANGE) {        // (RANGE: a pathological number of start codes -- the host scan takes it)
            error = std::string("stream upload: ") + mp2v_recon_last_error(failed_on);
            last.reset();
            return false;
        }
    }
    if (!S.resident && !index_stream(buffer, (size_t)len, index, cfg.num_threads)) { error = index.error; last.reset(); return false; }
    publish_headers(owner, index.headers);
    for (const auto& pic : index.pictures) {
        if (pic.seq.chroma_format != cfg.chroma_format) { error = "stream chroma_format differs from decoder_config_t.chroma_format"; last.reset(); return false; }
        // the reference trusts decoder_config_t blindly (decoder.cpp:44-66); a stream of another coded size would be
        // reconstructed into the wrong geometry, so it is refused here
        if (pic.seq.have_sequence_header && (((pic.seq.horizontal_size + 15) & ~15) != cfg.width || ((pic.seq.vertical_size + 15) & ~15) != cfg.height)) {
            error = "stream coded size " + std::to_string(pic.seq.horizontal_size) + "x" + std::to_string(pic.seq.vertical_size) +
                    " (rounded up to macroblocks) differs from decoder_config_t " + std::to_string(cfg.width) + "x" + std::to_string(cfg.height);
            last.reset();
            return false;
        }
    }
    S.gpu_vlc = opt.gpu_vlc && device_parser_can_take(index, !S.resident);
    S.stream_mode = S.gpu_vlc && S.resident;
    if (!prepare(S.gpu_vlc)) { last.reset(); return false; }
    const std::vector<device_ctx_t>& devs = dev_sets[S.gpu_vlc ? 1 : 0];
    if (S.stream_mode && devs.size() > 1) {
        // GOP sharding (chain g -> device g mod N): every device receives the byte ranges of its own pictures, at the same offsets
        for (size_t d = 0; d < devs.size(); d++) {
            std::vector<mp2v_byte_range_t> ranges;
            for (const coded_picture_t& pic : index.pictures) {
                if ((size_t)pic.gop % devs.size() != d || pic.slices.empty()) continue;
                const size_t lo = (size_t)(pic.slices.front().payload - 4 - buffer);
                const size_t hi = (size_t)(pic.slices.back().payload + pic.slices.back().bytes - buffer);
                if (!ranges.empty() && lo <= ranges.back().offset + ranges.back().bytes + 4096) ranges.back().bytes = hi - ranges.back().offset;   // (picture headers in between)
                else ranges.push_back({lo, hi - lo});
            }
            if (ranges.empty()) continue;
            if (mp2v_recon_stream_add(devs[d].recon, ranges.data(), (int)ranges.size()) != MP2V_OK) {
                error = std::string("stream upload (CUDA device ") + std::to_string(devs[d].device) + "): " + mp2v_recon_last_error(devs[d].recon);
                last.reset();
                return false;
            }
        }
    }
    S.index_ms = std::chrono::duration<double, std::milli>(clock_t_::now() - t_begin).count();
    return true;
}

// decode the opened stream: one pipeline per device; GOP chain g -> device g mod N
bool mp2v_decoder_c::impl_t::run(mp2v_decoder_c& owner, clock_t_::time_point t_begin, bool fresh_stats) {
    (void)owner;
    impl_t* m = this;
    const stream_state_t& S = *last;
    const stream_index_t& index = S.index;
    const auto t_indexed = clock_t_::now();
    shared_t sh;
    sh.t_origin = t_begin;
    sh.cfg = cfg; sh.opt = opt; sh.renderer = renderer;
    sh.mbw = cfg.width / 16; sh.mbh = cfg.height / 16;
    sh.gop_size.assign(index.n_gops > 0 ? index.n_gops : 1, 0);
    sh.gop_emitted.assign(sh.gop_size.size(), 0);
    for (const auto& pic : index.pictures) sh.gop_size[pic.gop]++;
    sh.opt.gpu_vlc = S.gpu_vlc;
    sh.stream_mode = S.stream_mode;
    sh.stream_base = S.buffer;
    const std::vector<device_ctx_t>& devs = dev_sets[sh.opt.gpu_vlc ? 1 : 0];
    std::deque<pipeline_t> pipes(devs.size());
    for (size_t d = 0; d < pipes.size(); d++) {
        pipeline_t& p = pipes[d];
        p.sh = &sh; p.device = devs[d].device; p.recon = devs[d].recon;
        p.n_frames = devs[d].n_frames; p.n_slots = devs[d].n_slots; p.parse_window = p.n_slots / 2;
        p.frame_use.assign(p.n_frames, 0);
        sh.pipes.push_back(&p);
        mp2v_recon_stats_t st;
        if (fresh_stats) mp2v_recon_get_stats(p.recon, &st, 1);   // statistics are per call (decode() has reset them before its upload)
        mp2v_recon_timer_start(p.recon);         // device time of the call: CUDA events on the stream every launch waits on / runs in
    }
    for (const auto& pic : index.pictures) {
        pipeline_t& p = pipes[(size_t)pic.gop % pipes.size()];
        p.tasks.emplace_back();
        pic_task_t& t = p.tasks.back();
        t.src = &pic; t.pipe = &p; t.local_index = (int)p.tasks.size() - 1;
    }
    const auto t_setup = clock_t_::now();
    std::vector<std::thread> workers;
    {
        int nthreads = m->cfg.num_threads > MAX_NUM_THREADS ? MAX_NUM_THREADS : m->cfg.num_threads;
        if (sh.opt.gpu_vlc && nthreads > 4) nthreads = 4;        // no host slice parsing: workers only stage the coded bytes
        if (sh.stream_mode) nthreads = 0;                        // ... and with the stream resident on the device there is nothing to stage
        for (int i = 0; i < nthreads; i++) workers.emplace_back(worker_main, &sh);
        for (auto& p : pipes) if (!p.tasks.empty()) {
            p.feeder_thread = std::thread(&pipeline_t::feeder, &p);
            p.output_thread = std::thread(&pipeline_t::output, &p);
        }
        for (auto& p : pipes) {
            if (p.feeder_thread.joinable()) p.feeder_thread.join();
            if (p.output_thread.joinable()) p.output_thread.join();
        }
        { std::lock_guard<std::mutex> lk(sh.qmu); sh.stop = true; }
        sh.qcv.notify_all();
        for (auto& w : workers) w.join();
    }
    const auto t_joined = clock_t_::now();
    m->stats = stats_t();
    for (auto& p : pipes) {
        if (mp2v_recon_sync(p.recon) != MP2V_OK && !sh.failed.load()) sh.fail(std::string("sync: ") + mp2v_recon_last_error(p.recon));
        double dev_ms = 0;
        if (mp2v_recon_timer_stop(p.recon, &dev_ms) == MP2V_OK && dev_ms > m->stats.device_ms) m->stats.device_ms = dev_ms;
        mp2v_recon_stats_t st;
        if (mp2v_recon_get_stats(p.recon, &st, 0) == MP2V_OK) {
            m->stats.pictures += st.pictures; m->stats.launches += st.launches; m->stats.h2d_bytes += st.h2d_bytes;
            m->stats.d2h_bytes += st.d2h_bytes; m->stats.algorithmic_bytes += st.algorithmic_bytes; m->stats.kernel_ms += st.kernel_ms;
            m->stats.vlc_launches += st.vlc_launches;
        }
    }
    m->stats.parse_cpu_seconds = sh.parse_ns.load() * 1e-9;
    if (getenv("MP2V_PROFILE")) {
        auto ms = [](clock_t_::time_point a, clock_t_::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[mp2v profile] index %.2f ms  setup %.2f ms  pipeline %.2f ms  drain+stats %.2f ms\n", S.index_ms, ms(t_indexed, t_setup),
                ms(t_setup, t_joined), ms(t_joined, clock_t_::now()));
    }
    if (getenv("MP2V_PROFILE"))
        fprintf(stderr, "[mp2v profile] pictures %zu  parse %.1f ms(cpu)  worker idle %.1f ms(sum)  feeder wait %.1f work %.1f ms  precheck %.1f ms  submit(+lock) %.1f ms  output wait %.1f map %.1f ms\n",
                index.pictures.size(), sh.parse_ns.load() * 1e-6, sh.worker_idle_ns.load() * 1e-6, sh.feeder_wait_ns.load() * 1e-6, sh.feeder_work_ns.load() * 1e-6,
                sh.precheck_ns.load() * 1e-6, sh.submit_ns.load() * 1e-6, sh.out_wait_ns.load() * 1e-6, sh.out_map_ns.load() * 1e-6);
    if (const char* pv = getenv("MP2V_PROFILE")) if (atoi(pv) >= 2)
        for (auto& p : pipes)
            for (size_t k = 0; k < p.tasks.size(); k++) {
                const pic_task_t& t = p.tasks[k];
                fprintf(stderr, "[mp2v timeline] pic %3zu type %d  queued %7.3f  staged %7.3f  submitted %7.3f  mapped %7.3f ms\n", k, t.src->info.picture_coding_type,
                        t.ts_queued, t.ts_staged, t.ts_submitted, t.ts_mapped);
            }
    m->stats.wall_seconds = std::chrono::duration<double>(clock_t_::now() - t_begin).count();
    if (sh.failed.load()) {
        // pictures may be left acquired / queued inside the contexts: give every slot back (no re-allocation);
        // a context that cannot even do that (a CUDA error) is rebuilt by the next call
        bool reset_ok = true;
        for (auto& p : pipes) reset_ok = reset_ok && mp2v_recon_reset(p.recon) == MP2V_OK;
        if (!reset_ok) m->release();
        m->error = sh.error;
        return false;
    }
    return true;
}
