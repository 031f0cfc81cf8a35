// C ABI of the reconstruction back end (include/mp2v_recon.h): device frame pool, pinned + device
// picture arenas, launch batching and stream / event scheduling.  Replaces the roles of the
// reference's frame_c pool (decoder.cpp:44-105) and task_queue_c dependency tracking
// (threads.cpp) with CUDA streams and events.  No CPU reconstruction path exists in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

#include "host/numa.h"
#include "mp2v_recon.h"
#include "recon_kernels.cuh"
#include "vlc_kernel.cuh"

using namespace mp2v;

namespace {

thread_local std::string g_create_error;

enum slot_state_t { SLOT_FREE = 0, SLOT_FILLING, SLOT_QUEUED, SLOT_INFLIGHT, SLOT_RESIDENT };

struct slot_t {
    mp2v_picture_t pub{};
    uint8_t* h_arena = nullptr;        // (slices of the context's pools: one pinned, one device allocation for all slots)
    uint8_t* d_arena = nullptr;
    const uint8_t* d_params = nullptr; // where the picture's parameter block lies on the device: the arena's head, or the staged block's
    slot_state_t state = SLOT_FREE;
    cudaEvent_t done = nullptr;        // the event of the launch that consumed the slot (an entry of the context's launch-event ring)
    uint64_t alg_bytes = 0;
    bool prechecked = false;           // account_and_validate already ran for the records now in the slot
    uint64_t seq = 0;                  // submission order, to recycle the oldest in-flight slot first
    // device-side slice parsing (MP2V_RECON_DEVICE_VLC): staged bitstream + own stream (allocated by the first
    // stage_slices of the slot: the stream-resident front end needs neither), parse result
    uint8_t* h_staged = nullptr;
    uint8_t* d_staged = nullptr;
    vlc_slice_status_t* h_status = nullptr;   // pinned + mapped: one entry per slice, written by the kernel
    vlc_slice_status_t* d_status = nullptr;   // the device's address of h_status
    int n_slices = 0, rows_covered = 0;
    size_t staged_end = 0;
    bool staged = false;               // stage_slices done, submit_staged pending
    int trace_idx = -1;
    cudaStream_t s_vlc = nullptr;
    cudaEvent_t vlc_done = nullptr;    // H2D + parse + status read-back of the picture now in the slot
    bool vlc = false;                  // the slot's records come from the device parser
    cudaEvent_t vlc_wait = nullptr;    // what the reconstruction launch waits for: vlc_done, or the event of the batched parse launch
    bool stream_pic = false;           // handed over with mp2v_recon_submit_stream_picture: parsed by the next batched parse launch
    bool need_blank = false;           // ... and some macroblock row has no slice: blank records first
    bool status_pending = false;       // h_status not yet folded into the statistics / error state
    uint64_t picture_no = 0;
};

constexpr size_t kParamsBytes = 512;   // sizeof(mp2v_pic_params_t) rounded up; mb records follow

}  // namespace

struct mp2v_recon {
    mp2v_recon_config_t cfg{};
    mp2v_frame_layout_t lay{};
    int nblk = 0, mbw = 0, mbh = 0, mb_count = 0, max_batch = 0;
    size_t frame_alloc = 0, arena_bytes = 0, coef_off = 0;
    uint8_t* h_arena_pool = nullptr; uint8_t* d_arena_pool = nullptr; uint8_t* h_status_pool = nullptr;
    uint8_t* d_frames = nullptr;
    recon_tmaps_t tmaps{};                     // TMA descriptors of the frame pool (reference windows)
    std::vector<uint8_t*> h_frames;            // pinned mirrors, allocated on first map
    // One event per reconstruction LAUNCH (a ring): frames and slots remember which launch wrote / consumed them.
    // An entry re-recorded by a later launch only makes a waiter wait longer (same stream), never less.
    std::vector<cudaEvent_t> launch_ev;
    uint64_t launch_seq = 0;
    std::vector<int> frame_launch;             // frame id -> index into launch_ev of its last writer (-1: none / an upload)
    std::vector<cudaEvent_t> frame_ev;         // last writer of a frame filled by mp2v_recon_upload_frame
    std::vector<uint8_t> frame_written;
    // MP2V_RECON_AUTO_DOWNLOAD: every submitted picture's frame is copied to its pinned mirror right behind its launch
    bool auto_dl = false;
    std::vector<cudaEvent_t> mirror_ev;        // ring of per-LAUNCH events: the mirror copies of all frames of one launch
    std::vector<int> mirror_ev_of;             // frame id -> index into mirror_ev
    uint64_t mirror_batches = 0;
    std::vector<uint8_t> mirror_valid;         // mirror_ev covers the frame's current content
    // a copy of a frame to the host (D2H stream) must finish before a later picture overwrites the frame (compute stream)
    std::vector<cudaEvent_t> read_ev;          // per frame: behind its last queued copy
    std::vector<uint8_t> read_pending;         // since the frame's last writer was launched: 1 = read_ev[f] recorded, 2 = a launch's mirror event covers a copy
    uint8_t* h_pool = nullptr;                 // the pinned mirrors as ONE block laid out like the device pool: runs of frame ids copy as one piece
    cudaStream_t s_copy = nullptr, s_compute = nullptr, s_d2h = nullptr, s_d2h2 = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    std::vector<slot_t> slots;
    std::vector<int> pending;                  // queued slots, submit order; launched by dependency level (flush_locked)
    int queued = 0;                            // = pending.size()
    std::mutex mu;
    std::string err;
    uint64_t seq = 0;
    // statistics
    mp2v_recon_stats_t stats{};
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;   // (start, stop) of launches not yet summed
    std::vector<cudaEvent_t> ev_pool;
    std::vector<cudaEvent_t> timing_pool;      // timing-enabled events for the per-launch stopwatch
    // device-side slice parsing
    bool vlc = false;
    int vlc_lanes = 1;
    size_t staged_bytes = 0;                   // per-slot staging capacity (header + slice table + bitstream)
    uint32_t slice_region = 0;
    void* d_tables = nullptr;
    uint8_t* d_blank_mb = nullptr;             // mb_count blank records, copied over a slot's records before each parse
    // MP2V_TRACE=1 (development): device timestamps per picture, dumped to stderr by mp2v_recon_sync
    struct trace_rec_t { uint64_t picture_no; cudaEvent_t h2d, vlc, recon, d2h; };
    struct launch_trace_t { int n; cudaEvent_t begin, end, copied; };      // MP2V_TRACE: one per reconstruction launch
    std::vector<launch_trace_t> launch_log;
    std::vector<launch_trace_t> parse_log;      // (begin, end of a parse launch; copied = the resident stream's arrival)
    bool trace = false;
    cudaEvent_t trace_base = nullptr;
    std::vector<trace_rec_t> trace_log;
    std::deque<int> status_fifo;               // slots whose parse status has not been folded in yet, submission order
    std::string vlc_error;                     // sticky: first slice error reported by the device parser
    // stream-resident front end (mp2v_recon_stream_begin / mp2v_recon_submit_stream_picture)
    static constexpr int kParseStreams = 4, kParseBufs = 8;
    uint8_t* d_stream = nullptr;               // device copy of the elementary stream
    size_t stream_cap = 0, stream_len = 0;
    const uint8_t* h_stream = nullptr;         // the caller's copy (valid until the next stream_begin): slice start codes are validated from it
    uint32_t* d_codes = nullptr; uint32_t* h_codes = nullptr; uint32_t codes_cap = 0;
    uint32_t* d_counts = nullptr; size_t counts_cap = 0;
    uint32_t* h_total = nullptr; uint32_t* d_total = nullptr;   // pinned + mapped
    cudaStream_t s_parse[kParseStreams] = {};
    int sm_count = 148;
    unsigned long long* d_counters = nullptr;  // batch_desc_t::counters
    int numa_node = -1;                        // of the device's PCIe root; -1 on single-node hosts
    int parse_rr = 0, n_parse_streams = 2, lot_cap = 0, parse_lanes = 0;      // dev knobs MP2V_PARSE_STREAMS / MP2V_LOT
    cudaEvent_t ev_stream = nullptr;           // the upload (and scan) of the resident stream
    bool scan_pending = false;                 // a start-code scan has been launched and its list not fetched yet
    cudaEvent_t ev_stream_timed = nullptr;     // MP2V_TRACE: the same moment with a timestamp
    struct parse_buf_t { uint8_t* h = nullptr; uint8_t* d = nullptr; cudaEvent_t done = nullptr; bool used = false; } parse_buf[kParseBufs];
    int parse_buf_rr = 0;
    size_t desc_stride = 0;
    std::vector<int> parse_pending;            // stream pictures whose descriptors sit in parse_buf[parse_buf_rr], parse not launched yet
    uint64_t pictures_submitted = 0;
    int batch_ramp = 1;                        // launch batch limit right after a sync: 1, 2, 4, ... max_batch (first frames out early)

    int fail(int code, const std::string& what) { err = what; return code; }
    int cuda_fail(cudaError_t e, const char* what) {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return MP2V_ERR_CUDA;
    }
    uint8_t* frame_ptr(int id, int plane) const { return d_frames + (size_t)id * frame_alloc + lay.plane_offset[plane]; }
    cudaEvent_t get_event() {
        if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        return e;
    }
};

extern "C" MP2V_API int mp2v_frame_layout(int width, int height, int chroma_format, mp2v_frame_layout_t* out) {
    // frame_c::frame_c (decoder.cpp:44-66): 64-byte aligned strides, chroma per format
    if (!out || width <= 0 || height <= 0 || (width & 15) || (height & 15) || chroma_format < 1 || chroma_format > 3) return MP2V_ERR_ARG;
    out->width[0] = width; out->height[0] = height;
    out->stride[0] = (width + 63) & ~63;
    const bool full = chroma_format == 3;
    out->width[1] = full ? width : width >> 1;
    out->height[1] = chroma_format == 1 ? height >> 1 : height;
    out->stride[1] = full ? out->stride[0] : ((out->stride[0] >> 1) + 63) & ~63;
    out->width[2] = out->width[1]; out->height[2] = out->height[1]; out->stride[2] = out->stride[1];
    size_t off = 0;
    for (int p = 0; p < 3; p++) {
        out->plane_offset[p] = off;
        off += ((size_t)out->stride[p] * out->height[p] + 255) & ~(size_t)255;
    }
    out->bytes = off;
    return MP2V_OK;
}

#define CK(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ctx->cuda_fail(e_, what); } while (0)

static void destroy_ctx(mp2v_recon* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->s_compute) cudaStreamSynchronize(ctx->s_compute);
    if (ctx->s_copy) cudaStreamSynchronize(ctx->s_copy);
    if (ctx->s_d2h) cudaStreamSynchronize(ctx->s_d2h);
    if (ctx->s_d2h2) cudaStreamSynchronize(ctx->s_d2h2);
    for (auto st : ctx->s_parse) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (auto& b : ctx->parse_buf) { if (b.h) cudaFreeHost(b.h); if (b.d) cudaFree(b.d); if (b.done) cudaEventDestroy(b.done); }
    if (ctx->d_stream) cudaFree(ctx->d_stream);
    if (ctx->d_codes) cudaFree(ctx->d_codes);
    if (ctx->h_codes) cudaFreeHost(ctx->h_codes);
    if (ctx->d_counts) cudaFree(ctx->d_counts);
    if (ctx->h_total) cudaFreeHost(ctx->h_total);
    if (ctx->ev_stream) cudaEventDestroy(ctx->ev_stream);
    if (ctx->ev_stream_timed) cudaEventDestroy(ctx->ev_stream_timed);
    for (auto e : ctx->launch_ev) if (e) cudaEventDestroy(e);
    for (auto& s : ctx->slots) {
        if (s.s_vlc) { cudaStreamSynchronize(s.s_vlc); cudaStreamDestroy(s.s_vlc); }
        if (s.vlc_done) cudaEventDestroy(s.vlc_done);
        if (s.h_staged) cudaFreeHost(s.h_staged);
        if (s.d_staged) cudaFree(s.d_staged);
    }
    if (ctx->h_arena_pool) cudaFreeHost(ctx->h_arena_pool);
    if (ctx->d_arena_pool) cudaFree(ctx->d_arena_pool);
    if (ctx->h_status_pool) cudaFreeHost(ctx->h_status_pool);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->d_tables) cudaFree(ctx->d_tables);
    if (ctx->d_blank_mb) cudaFree(ctx->d_blank_mb);
    if (ctx->h_pool) cudaFreeHost(ctx->h_pool);
    else for (auto* h : ctx->h_frames) if (h) cudaFreeHost(h);
    for (auto e : ctx->frame_ev) if (e) cudaEventDestroy(e);
    for (auto e : ctx->mirror_ev) if (e) cudaEventDestroy(e);
    for (auto e : ctx->read_ev) if (e) cudaEventDestroy(e);
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    for (auto e : ctx->timing_pool) cudaEventDestroy(e);
    for (auto& pr : ctx->timed) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    if (ctx->ev_h2d) cudaEventDestroy(ctx->ev_h2d);
    if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
    if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
    if (ctx->d_frames) cudaFree(ctx->d_frames);
    if (ctx->s_copy) cudaStreamDestroy(ctx->s_copy);
    if (ctx->s_compute) cudaStreamDestroy(ctx->s_compute);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    if (ctx->s_d2h2) cudaStreamDestroy(ctx->s_d2h2);
    delete ctx;
}

static int create_impl(mp2v_recon* ctx) {
    const mp2v_recon_config_t& c = ctx->cfg;
    if (mp2v_frame_layout(c.width, c.height, c.chroma_format, &ctx->lay) != MP2V_OK) return ctx->fail(MP2V_ERR_ARG, "bad geometry");
    if (c.n_frames < 1 || c.n_pictures < 1 || c.max_batch < 0 || c.max_batch > kMaxBatch) return ctx->fail(MP2V_ERR_ARG, "bad pool sizes");
    ctx->nblk = c.chroma_format == 1 ? 6 : c.chroma_format == 2 ? 8 : 12;
    ctx->mbw = c.width / 16; ctx->mbh = c.height / 16; ctx->mb_count = ctx->mbw * ctx->mbh;
    ctx->max_batch = c.max_batch ? c.max_batch : 8;
    // One stream per picture slot parses concurrently: with the default 8 hardware work queues the
    // streams alias and independent parses queue up behind one another (measured: 8 at a time).  Only
    // effective when this is the process's first CUDA call; otherwise the embedding application sets it.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return ctx->fail(MP2V_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (c.device < 0 || c.device >= ndev) return ctx->fail(MP2V_ERR_ARG, "device ordinal out of range");
    CK(cudaSetDevice(c.device), "cudaSetDevice");
    // the pinned memory of this context (record arenas, frame mirrors, parse status) is allocated while the calling thread
    // sits on the device's NUMA node; a no-op on single-node hosts (host/numa.h)
    char bus_id[32] = {0};
    if (cudaDeviceGetPCIBusId(bus_id, (int)sizeof(bus_id), c.device) == cudaSuccess) ctx->numa_node = numa_node_of_pci_device(bus_id);
    cudaGetLastError();
    numa_scope_t numa_scope(ctx->numa_node);
    cudaFuncAttributes fa;
    e = recon_kernel_attributes(c.chroma_format, &fa);   // fails loudly when the sm_100a image cannot load on this device
    if (e != cudaSuccess) return ctx->cuda_fail(e, "reconstruction kernel image not usable on this device (built for sm_100a only)");
    // Priorities: frames leaving the device first, then reconstruction, then (default priority) the slice
    // parses -- of which dozens are resident at any time and which otherwise crowd out the stages that
    // free their slots and frames.
    int prio_least = 0, prio_greatest = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest), "stream priority range");
    const int prio_mid = prio_greatest < prio_least - 1 ? prio_greatest + 1 : prio_greatest;
    CK(cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking), "stream");
    CK(cudaStreamCreateWithPriority(&ctx->s_compute, cudaStreamNonBlocking, prio_mid), "stream");
    CK(cudaStreamCreateWithPriority(&ctx->s_d2h, cudaStreamNonBlocking, prio_greatest), "stream");
    CK(cudaStreamCreateWithPriority(&ctx->s_d2h2, cudaStreamNonBlocking, prio_greatest), "stream");
    CK(cudaEventCreateWithFlags(&ctx->ev_h2d, cudaEventDisableTiming), "event");
    CK(cudaEventCreate(&ctx->ev_t0), "event");
    CK(cudaEventCreate(&ctx->ev_t1), "event");
    // frames: the window staging over-reads one row below and 16 bytes right of a block (always inside
    // this slack), so every frame carries two luma rows + 256 bytes of tail
    ctx->frame_alloc = (ctx->lay.bytes + 2 * (size_t)ctx->lay.stride[0] + 256 + 255) & ~(size_t)255;
    CK(cudaMalloc(&ctx->d_frames, ctx->frame_alloc * c.n_frames), "cudaMalloc frames");
    CK(cudaMemset(ctx->d_frames, 0, ctx->frame_alloc * c.n_frames), "cudaMemset frames");
    CK(make_frame_tmaps(c.chroma_format, ctx->d_frames, ctx->lay, ctx->frame_alloc, c.n_frames, &ctx->tmaps), "TMA descriptors of the frame pool");
    ctx->h_frames.assign(c.n_frames, nullptr);
    ctx->frame_ev.assign(c.n_frames, nullptr);
    ctx->frame_written.assign(c.n_frames, 0);
    for (auto& ev : ctx->frame_ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "event");
    ctx->frame_launch.assign(c.n_frames, -1);
    ctx->launch_ev.assign((size_t)c.n_pictures + c.n_frames + 16, nullptr);
    for (auto& ev : ctx->launch_ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "event");
    if (const char* v = getenv("MP2V_TRACE")) ctx->trace = atoi(v) != 0;
    CK(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, c.device), "device attribute");
    CK(cudaMalloc(&ctx->d_counters, 3 * sizeof(unsigned long long)), "cudaMalloc counters");
    CK(cudaMemset(ctx->d_counters, 0, 3 * sizeof(unsigned long long)), "cudaMemset counters");
    ctx->auto_dl = (c.flags & MP2V_RECON_AUTO_DOWNLOAD) != 0;
    ctx->mirror_valid.assign(c.n_frames, 0);
    ctx->read_ev.assign(c.n_frames, nullptr);
    ctx->read_pending.assign(c.n_frames, 0);
    for (auto& ev : ctx->read_ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "event");
    if (ctx->auto_dl) {
        // one event per launch, not per frame: an event between two copies costs the copy engine a bubble of
        // several microseconds (measured 67 vs 60 us per 3 MB frame).  At most n_frames launches are outstanding.
        ctx->mirror_ev.assign(c.n_frames + 1, nullptr);
        ctx->mirror_ev_of.assign(c.n_frames, 0);
        for (auto& ev : ctx->mirror_ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "event");
        CK(cudaHostAlloc(&ctx->h_pool, ctx->frame_alloc * c.n_frames, cudaHostAllocDefault), "cudaHostAlloc frame mirrors");
        for (int f = 0; f < c.n_frames; f++) ctx->h_frames[f] = ctx->h_pool + (size_t)f * ctx->frame_alloc;
    }
    // picture slots
    const uint64_t worst = (uint64_t)ctx->mb_count * ctx->nblk * 64u;
    if (worst > 0xffffffffull) return ctx->fail(MP2V_ERR_ARG, "picture too large");
    ctx->vlc = (c.flags & MP2V_RECON_DEVICE_VLC) != 0;
    // With the device parser every slice owns a worst-case region of the device arena (sparse use of
    // plentiful HBM instead of a second counting pass), and the host side of a slot holds no
    // coefficient records at all.
    const uint32_t cap = ctx->vlc ? (uint32_t)worst : (c.coef_capacity ? c.coef_capacity : (uint32_t)worst);
    const uint32_t host_cap = ctx->vlc ? 0u : cap;
    ctx->coef_off = kParamsBytes + (((size_t)ctx->mb_count * sizeof(mp2v_mb_info_t) + 255) & ~(size_t)255);
    // + 256: the kernel requests up to 32 records past a macroblock's last one (prefetch of the next batch)
    ctx->arena_bytes = ctx->coef_off + (size_t)cap * sizeof(mp2v_coef_t) + 256;
    const size_t host_arena_bytes = ctx->coef_off + (size_t)host_cap * sizeof(mp2v_coef_t);
    if (ctx->vlc) {
        cudaFuncAttributes va;
        e = vlc_kernel_attributes(&va);
        if (e != cudaSuccess) return ctx->cuda_fail(e, "slice parser kernel image not usable on this device (built for sm_100a only)");
        CK(vlc_upload_tables(&ctx->d_tables), "slice parser tables");
        ctx->slice_region = (uint32_t)ctx->mbw * ctx->nblk * 64u;
        const size_t bits_cap = c.bitstream_capacity ? c.bitstream_capacity : std::max<size_t>(2u << 20, (size_t)ctx->mb_count * 128u);
        ctx->staged_bytes = kVlcParamsBytes + sizeof(vlc_pic_header_t) + (size_t)ctx->mbh * sizeof(vlc_slice_t) + 16 + bits_cap + 16;
        if (const char* v = getenv("MP2V_VLC_LANES")) { const int l = atoi(v); if (l >= 1 && l <= 32) ctx->vlc_lanes = l; }
        std::vector<mp2v_mb_info_t> blank((size_t)ctx->mb_count, mp2v_mb_info_t{0u, MP2V_MB_BITS(0, 1, 0, MP2V_MB_INTRA), {{0, 0}, {0, 0}}});
        CK(cudaMalloc(&ctx->d_blank_mb, blank.size() * sizeof(mp2v_mb_info_t)), "cudaMalloc blank records");
        CK(cudaMemcpy(ctx->d_blank_mb, blank.data(), blank.size() * sizeof(mp2v_mb_info_t), cudaMemcpyHostToDevice), "H2D blank records");
    }
    // One pinned and one device allocation hold the arenas of all slots (a context of 128 slots used to make 640 driver
    // calls and pin 256 MB of staging nobody might use: 0.1 s per decoder object, measured with 64 decoders in a process).
    ctx->slots.resize(c.n_pictures);
    ctx->arena_bytes = (ctx->arena_bytes + 255) & ~(size_t)255;
    const size_t host_arena_stride = (host_arena_bytes + 255) & ~(size_t)255;
    const size_t status_stride = ((size_t)ctx->mbh * sizeof(vlc_slice_status_t) + 63) & ~(size_t)63;
    CK(cudaHostAlloc(&ctx->h_arena_pool, host_arena_stride * c.n_pictures, cudaHostAllocDefault), "cudaHostAlloc picture arenas");
    CK(cudaMalloc(&ctx->d_arena_pool, ctx->arena_bytes * c.n_pictures), "cudaMalloc picture arenas");
    uint8_t* d_status_pool = nullptr;
    if (ctx->vlc) {
        CK(cudaHostAlloc(&ctx->h_status_pool, status_stride * c.n_pictures, cudaHostAllocMapped), "cudaHostAlloc parse status");
        CK(cudaHostGetDevicePointer(&d_status_pool, ctx->h_status_pool, 0), "cudaHostGetDevicePointer");
    }
    for (int i = 0; i < c.n_pictures; i++) {
        slot_t& s = ctx->slots[i];
        s.h_arena = ctx->h_arena_pool + host_arena_stride * i;
        s.d_arena = ctx->d_arena_pool + ctx->arena_bytes * i;
        s.d_params = s.d_arena;
        s.pub.params = reinterpret_cast<mp2v_pic_params_t*>(s.h_arena);
        s.pub.mb = reinterpret_cast<mp2v_mb_info_t*>(s.h_arena + kParamsBytes);
        s.pub.coef = host_cap ? reinterpret_cast<mp2v_coef_t*>(s.h_arena + ctx->coef_off) : nullptr;
        s.pub.mb_count = (uint32_t)ctx->mb_count;
        s.pub.coef_capacity = host_cap;
        s.pub.slot = i;
        if (ctx->vlc) {
            s.h_status = reinterpret_cast<vlc_slice_status_t*>(ctx->h_status_pool + status_stride * i);
            s.d_status = reinterpret_cast<vlc_slice_status_t*>(d_status_pool + status_stride * i);
        }
    }
    if (ctx->vlc) {
        // stream-resident front end: parse streams, descriptor buffers (one per batched parse launch in flight)
        // (development knobs of tools/dev/*_sweep.sh: concurrent parse launches, slices per warp, pictures per parse launch)
        if (const char* v = getenv("MP2V_PARSE_STREAMS")) { const int k = atoi(v); if (k >= 1 && k <= mp2v_recon::kParseStreams) ctx->n_parse_streams = k; }
        if (const char* v = getenv("MP2V_PARSE_LANES")) { const int k = atoi(v); if (k >= 0 && k <= 32) ctx->parse_lanes = k; }
        if (const char* v = getenv("MP2V_LOT")) { const int k = atoi(v); if (k >= 1 && k <= kMaxStreamBatch) ctx->lot_cap = k; }
        for (auto& st : ctx->s_parse) CK(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio_least), "stream");
        CK(cudaEventCreateWithFlags(&ctx->ev_stream, cudaEventDisableTiming), "event");
        ctx->desc_stride = vlc_stream_desc_bytes(ctx->mbh);
        for (auto& b : ctx->parse_buf) {
            CK(cudaHostAlloc(&b.h, ctx->desc_stride * kMaxStreamBatch, cudaHostAllocDefault), "cudaHostAlloc parse descriptors");
            CK(cudaMalloc(&b.d, ctx->desc_stride * kMaxStreamBatch), "cudaMalloc parse descriptors");
            CK(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming), "event");
        }
        CK(cudaHostAlloc(&ctx->h_total, 64, cudaHostAllocMapped), "cudaHostAlloc scan total");
        CK(cudaHostGetDevicePointer(&ctx->d_total, ctx->h_total, 0), "cudaHostGetDevicePointer");
    }
    CK(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_recon_create(const mp2v_recon_config_t* cfg, mp2v_recon_t** out) {
    if (!cfg || !out) return MP2V_ERR_ARG;
    static_assert(sizeof(mp2v_pic_params_t) <= kParamsBytes && kVlcParamsBytes == kParamsBytes, "params block");
    static_assert(sizeof(mp2v_mb_info_t) == 16, "mb record");
    mp2v_recon* ctx = new mp2v_recon();
    ctx->cfg = *cfg;
    const int rc = create_impl(ctx);
    if (rc != MP2V_OK) {
        g_create_error = ctx->err;
        destroy_ctx(ctx);
        *out = nullptr;
        return rc;
    }
    *out = ctx;
    return MP2V_OK;
}

extern "C" MP2V_API void mp2v_recon_destroy(mp2v_recon_t* ctx) { destroy_ctx(ctx); }

extern "C" MP2V_API const char* mp2v_recon_last_error(mp2v_recon_t* ctx) {
    if (!ctx) return g_create_error.c_str();
    // several threads use one context: hand out a per-thread copy taken under the lock, not the shared string
    thread_local std::string copy;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        copy = ctx->err;
    }
    return copy.c_str();
}

// ---------------------------------------------------------------------------------------------
// launching

static void fill_desc(mp2v_recon* ctx, const slot_t& s, pic_desc_t& d) {
    const mp2v_pic_params_t& pp = *s.pub.params;
    d.params = reinterpret_cast<const mp2v_pic_params_t*>(s.d_params);
    d.mb = reinterpret_cast<const mp2v_mb_info_t*>(s.d_arena + kParamsBytes);
    d.coef = reinterpret_cast<const mp2v_coef_t*>(s.d_arena + ctx->coef_off);
    for (int p = 0; p < 3; p++) {
        d.dst[p] = ctx->frame_ptr(pp.dst_frame, p);
        d.l0[p] = pp.l0_frame >= 0 ? ctx->frame_ptr(pp.l0_frame, p) : nullptr;
        d.l1[p] = pp.l1_frame >= 0 ? ctx->frame_ptr(pp.l1_frame, p) : nullptr;
    }
    d.l0_id = pp.l0_frame; d.l1_id = pp.l1_frame;
}

// one launch over `ids` (<= max_batch slots whose records are already on, or on their way to, the device)
static int launch_slots(mp2v_recon* ctx, const int* ids, int n, bool download = false) {
    batch_desc_t b{};
    b.n_pics = n;
    b.mbw = ctx->mbw; b.mbh = ctx->mbh; b.mb_count = ctx->mb_count;
    for (int p = 0; p < 3; p++) b.stride[p] = ctx->lay.stride[p];
    b.mbs_per_warp = choose_mbs_per_warp(ctx->cfg.chroma_format, (long long)n * ctx->mb_count);
    b.counters = ctx->timing ? ctx->d_counters : nullptr;
    const int g = b.mbs_per_warp * (kCtaThreads / 32);
    b.ctas_per_pic = (ctx->mb_count + g - 1) / g;
    for (int i = 0; i < n; i++) fill_desc(ctx, ctx->slots[ids[i]], b.pic[i]);
    // a frame whose copy to the host is still queued must not be overwritten under it
    for (int i = 0; i < n; i++) {
        const int f = ctx->slots[ids[i]].pub.params->dst_frame;
        if (ctx->read_pending[f] & 1) CK(cudaStreamWaitEvent(ctx->s_compute, ctx->read_ev[f], 0), "stream wait");
        if (ctx->read_pending[f] & 2) CK(cudaStreamWaitEvent(ctx->s_compute, ctx->mirror_ev[ctx->mirror_ev_of[f]], 0), "stream wait");
        ctx->read_pending[f] = 0;
    }
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (ctx->timing) {
        // timing events are recycled (get_stats hands them back): creating a pair per launch cost ~10 us under the lock
        for (cudaEvent_t* e : {&t0, &t1}) {
            if (!ctx->timing_pool.empty()) { *e = ctx->timing_pool.back(); ctx->timing_pool.pop_back(); }
            else CK(cudaEventCreate(e), "event");
        }
        CK(cudaEventRecord(t0, ctx->s_compute), "event record");
    }
    mp2v_recon::launch_trace_t lt{n, nullptr, nullptr, nullptr};
    if (ctx->trace) {
        if (!ctx->trace_base) { CK(cudaEventCreate(&ctx->trace_base), "event"); CK(cudaEventRecord(ctx->trace_base, ctx->s_compute), "event record"); }
        CK(cudaEventCreate(&lt.begin), "event"); CK(cudaEventCreate(&lt.end), "event"); CK(cudaEventCreate(&lt.copied), "event");
        CK(cudaEventRecord(lt.begin, ctx->s_compute), "event record");
    }
    CK(launch_recon(ctx->cfg.chroma_format, b, ctx->tmaps, ctx->s_compute), "reconstruction kernel launch");
    if (ctx->trace) CK(cudaEventRecord(lt.end, ctx->s_compute), "event record");
    if (ctx->timing) {
        CK(cudaEventRecord(t1, ctx->s_compute), "event record");
        ctx->timed.emplace_back(t0, t1);
    }
    // ONE event for the whole launch: its frames and slots point at it
    const int lidx = (int)(ctx->launch_seq++ % ctx->launch_ev.size());
    CK(cudaEventRecord(ctx->launch_ev[lidx], ctx->s_compute), "event record");
    for (int i = 0; i < n; i++) {
        slot_t& s = ctx->slots[ids[i]];
        const int f = s.pub.params->dst_frame;
        ctx->frame_launch[f] = lidx;
        s.done = ctx->launch_ev[lidx];
        ctx->frame_written[f] = 1;
        ctx->mirror_valid[f] = 0;
        ctx->stats.algorithmic_bytes += s.alg_bytes;
        if (ctx->trace && s.vlc && s.trace_idx >= 0) CK(cudaEventRecord(ctx->trace_log[s.trace_idx].recon, ctx->s_compute), "event record");
    }
    if (download) {
        // The copies queue up behind this launch on one of two D2H streams (alternating, so that the copies of
        // consecutive launches overlap their gaps) and overlap the launches that follow.  The mirrors are laid out
        // like the device pool: a run of consecutive frame ids is ONE copy.
        cudaStream_t sd = (ctx->mirror_batches & 1) ? ctx->s_d2h2 : ctx->s_d2h;
        CK(cudaStreamWaitEvent(sd, ctx->launch_ev[lidx], 0), "stream wait");
        const int ev = (int)(ctx->mirror_batches++ % ctx->mirror_ev.size());
        int fr[kMaxBatch];
        for (int i = 0; i < n; i++) fr[i] = ctx->slots[ids[i]].pub.params->dst_frame;
        std::sort(fr, fr + n);
        for (int i = 0; i < n;) {
            int j = i + 1;
            while (j < n && fr[j] == fr[j - 1] + 1) j++;
            const size_t bytes = (size_t)(j - i - 1) * ctx->frame_alloc + ctx->lay.bytes;
            CK(cudaMemcpyAsync(ctx->h_frames[fr[i]], ctx->frame_ptr(fr[i], 0), bytes, cudaMemcpyDeviceToHost, sd), "D2H frames");
            ctx->stats.d2h_bytes += (uint64_t)(j - i) * ctx->lay.bytes;
            i = j;
        }
        for (int i = 0; i < n; i++) {
            const int f = fr[i];
            ctx->mirror_ev_of[f] = ev;
            ctx->mirror_valid[f] = 1;
            ctx->read_pending[f] |= 2;                     // covered by this launch's mirror event
        }
        if (ctx->trace) for (int i = 0; i < n; i++) if (ctx->slots[ids[i]].vlc && ctx->slots[ids[i]].trace_idx >= 0) CK(cudaEventRecord(ctx->trace_log[ctx->slots[ids[i]].trace_idx].d2h, sd), "event record");
        CK(cudaEventRecord(ctx->mirror_ev[ev], sd), "event record");
        if (ctx->trace) CK(cudaEventRecord(lt.copied, sd), "event record");
    }
    if (ctx->trace) ctx->launch_log.push_back(lt);
    ctx->stats.pictures += n;
    ctx->stats.launches += 1;
    return MP2V_OK;
}

// one parse launch for every stream picture handed over since the last one (ctx->mu held)
static int launch_parse_batch(mp2v_recon* ctx) {
    if (ctx->parse_pending.empty()) return MP2V_OK;
    mp2v_recon::parse_buf_t& b = ctx->parse_buf[ctx->parse_buf_rr];
    cudaStream_t st = ctx->s_parse[ctx->parse_rr];
    ctx->parse_rr = (ctx->parse_rr + 1) % ctx->n_parse_streams;
    const int n = (int)ctx->parse_pending.size();
    CK(cudaStreamWaitEvent(st, ctx->ev_stream, 0), "stream wait");      // the resident stream has arrived
    // rows without any slice (never in valid streams): blank records first
    for (int id : ctx->parse_pending) {
        slot_t& s = ctx->slots[id];
        if (s.need_blank)
            CK(cudaMemcpyAsync(s.d_arena + kParamsBytes, ctx->d_blank_mb, (size_t)ctx->mb_count * sizeof(mp2v_mb_info_t), cudaMemcpyDeviceToDevice, st), "blank records");
    }
    CK(cudaMemcpyAsync(b.d, b.h, (size_t)n * ctx->desc_stride, cudaMemcpyHostToDevice, st), "H2D parse descriptors");
    if (ctx->trace && ctx->trace_base) {
        mp2v_recon::launch_trace_t lt{n, nullptr, nullptr, nullptr};
        CK(cudaEventCreate(&lt.begin), "event"); CK(cudaEventCreate(&lt.end), "event");
        CK(cudaEventRecord(lt.begin, st), "event record");
        ctx->parse_log.push_back(lt);
    }
    // Slices per warp: a parser thread's registers are held by its whole warp, and a lot of single-lane warps fills every
    // SM's register file (8 CTAs of 128 threads x 64 registers) for a millisecond -- the reconstruction launches of the
    // lots before it then wait for room.  Lots of more than 16 pictures of 1080p put two slices in a warp (lanes of a warp
    // that sit on different slices serialise where their paths differ; measured on the 120-picture call: 2 lanes 3.08 ms,
    // 1 lane 3.5, 4 lanes 3.2, 8 lanes 3.6).  Launching the parse of a lot in parts ahead of the lot measured slower
    // (parts queue up behind one another on the parse streams).
    int lanes = ctx->parse_lanes;
    if (lanes == 0) lanes = n * ctx->mbh <= 2 * ctx->sm_count * 4 ? 1 : 2;
    CK(launch_vlc_stream(ctx->d_stream, b.d, ctx->desc_stride, n, ctx->mbh, lanes, ctx->d_tables, st), "slice parser kernel launch");
    CK(cudaEventRecord(b.done, st), "event record");
    if (ctx->trace && !ctx->parse_log.empty() && ctx->parse_log.back().end) CK(cudaEventRecord(ctx->parse_log.back().end, st), "event record");
    b.used = true;
    for (int id : ctx->parse_pending) ctx->slots[id].vlc_wait = b.done;
    ctx->stats.h2d_bytes += (uint64_t)n * ctx->desc_stride;
    ctx->stats.vlc_launches += 1;
    ctx->parse_pending.clear();
    // the next batch writes into the next buffer; wait (rare: kParseBufs launches deep) until its previous use has been read
    ctx->parse_buf_rr = (ctx->parse_buf_rr + 1) % mp2v_recon::kParseBufs;
    mp2v_recon::parse_buf_t& nb = ctx->parse_buf[ctx->parse_buf_rr];
    if (nb.used) { CK(cudaEventSynchronize(nb.done), "event sync"); nb.used = false; }
    return MP2V_OK;
}

// launch one group: H2D of host-parsed records / wait for device parses, then the reconstruction kernel (ctx->mu held)
static int launch_group(mp2v_recon* ctx, const std::vector<int>& group) {
    // H2D of every host-parsed picture: params + macroblock records + the used part of the coefficient arena
    // (pictures parsed on the device already have them there: the launch waits for their parse instead)
    bool any_h2d = false;
    cudaEvent_t waited = nullptr;
    for (int id : group) {
        slot_t& s = ctx->slots[id];
        if (s.vlc) {
            if (s.vlc_wait != waited) { CK(cudaStreamWaitEvent(ctx->s_compute, s.vlc_wait, 0), "stream wait"); waited = s.vlc_wait; }
            continue;
        }
        const size_t bytes = ctx->coef_off + (size_t)s.pub.params->n_coef * sizeof(mp2v_coef_t);
        CK(cudaMemcpyAsync(s.d_arena, s.h_arena, bytes, cudaMemcpyHostToDevice, ctx->s_copy), "H2D picture records");
        ctx->stats.h2d_bytes += bytes;
        any_h2d = true;
    }
    if (any_h2d) {
        CK(cudaEventRecord(ctx->ev_h2d, ctx->s_copy), "event record");
        CK(cudaStreamWaitEvent(ctx->s_compute, ctx->ev_h2d, 0), "stream wait");
    }
    const int rc = launch_slots(ctx, group.data(), (int)group.size(), ctx->auto_dl);
    if (rc != MP2V_OK) return rc;
    for (int id : group) ctx->slots[id].state = SLOT_INFLIGHT;
    return MP2V_OK;
}

// Launch everything that is queued: the pending parses in one launch, then the pictures by DEPENDENCY LEVEL.
// A picture's level is one more than the deepest queued picture it reads from, or whose frame accesses it must
// not overtake (it overwrites a frame that one reads or writes); pictures of one level are independent and share
// a launch, whatever GOP chain they belong to -- the I pictures of every queued GOP, then their P pictures, ...
// This is the reference's add_dependency order (decoder.cpp:294-305) taken over the whole lot.
static int flush_locked(mp2v_recon* ctx) {
    if (ctx->pending.empty() && ctx->parse_pending.empty()) return MP2V_OK;
    CK(cudaSetDevice(ctx->cfg.device), "cudaSetDevice");
    int rc = launch_parse_batch(ctx);
    if (rc != MP2V_OK) return rc;
    const int n = (int)ctx->pending.size();
    std::vector<int> level((size_t)n, 0);
    int top = 0;
    for (int i = 0; i < n; i++) {
        const mp2v_pic_params_t& p = *ctx->slots[ctx->pending[i]].pub.params;
        int lv = 0;
        for (int j = 0; j < i; j++) {
            const mp2v_pic_params_t& q = *ctx->slots[ctx->pending[j]].pub.params;
            if (q.dst_frame == p.l0_frame || q.dst_frame == p.l1_frame || q.dst_frame == p.dst_frame ||
                q.l0_frame == p.dst_frame || q.l1_frame == p.dst_frame) lv = std::max(lv, level[j] + 1);
        }
        level[i] = lv;
        top = std::max(top, lv);
    }
    std::vector<int> group;
    for (int lv = 0; lv <= top && n; lv++) {
        group.clear();
        for (int i = 0; i < n; i++) {
            if (level[i] != lv) continue;
            group.push_back(ctx->pending[i]);
            if ((int)group.size() == ctx->max_batch) { rc = launch_group(ctx, group); if (rc != MP2V_OK) return rc; group.clear(); }
        }
        if (!group.empty()) { rc = launch_group(ctx, group); if (rc != MP2V_OK) return rc; }
    }
    ctx->pending.clear();
    ctx->queued = 0;
    if (ctx->batch_ramp < std::max(ctx->max_batch, 16)) ctx->batch_ramp *= 2;
    return MP2V_OK;
}

static bool frame_is_queued(const mp2v_recon* ctx, int f) {
    for (int id : ctx->pending) if (ctx->slots[id].pub.params->dst_frame == f) return true;
    return false;
}

// OUT + REF + COEF + META, and (optionally) the host-side validation of the records
static int account_and_validate(mp2v_recon* ctx, slot_t& s, bool validate, std::string* why) {
    auto bad = [&](int code, const char* msg) { *why = msg; return code; };
    const mp2v_pic_params_t& pp = *s.pub.params;
    const int nf = ctx->cfg.n_frames;
    if (pp.dst_frame < 0 || pp.dst_frame >= nf || pp.l0_frame >= nf || pp.l1_frame >= nf) return bad(MP2V_ERR_ARG, "frame id out of range");
    if (pp.n_coef > s.pub.coef_capacity) return bad(MP2V_ERR_RANGE, "coefficient arena overflow");
    const uint64_t mb_bytes = ctx->cfg.chroma_format == 1 ? 384 : ctx->cfg.chroma_format == 2 ? 512 : 768;
    uint64_t out_bytes = 0;
    for (int p = 0; p < 3; p++) out_bytes += (uint64_t)ctx->lay.width[p] * ctx->lay.height[p];
    uint64_t ref = 0, coded = 0;
    const uint32_t cbp_mask = (1u << ctx->nblk) - 1u;
    const int W = ctx->cfg.width, H = ctx->cfg.height;
    for (int m = 0; m < ctx->mb_count; m++) {
        const mp2v_mb_info_t& r = s.pub.mb[m];
        const uint32_t bits = r.bits;
        const int ndir = (bits & MP2V_MB_INTRA) ? 0 : ((bits & MP2V_MB_FWD) ? 1 : 0) + ((bits & MP2V_MB_BWD) ? 1 : 0);
        ref += (uint64_t)ndir * mb_bytes;
        coded += (uint64_t)__builtin_popcount(MP2V_MB_CBP(bits) & cbp_mask);
        if (!validate) continue;
        if (MP2V_MB_CBP(bits) & ~cbp_mask) return bad(MP2V_ERR_RANGE, "coded_block_pattern names a block this chroma format does not have");
        if ((uint64_t)MP2V_MB_COEF_OFF(r.coef_off) + MP2V_MB_NCOEF(bits) > pp.n_coef) return bad(MP2V_ERR_RANGE, "macroblock coefficient range outside the arena");
        if (!(bits & MP2V_MB_INTRA)) {
            if (ndir == 0) return bad(MP2V_ERR_RANGE, "non-intra macroblock without a prediction direction");
            const int mbx = m % ctx->mbw, mby = m / ctx->mbw;
            for (int d = 0; d < 2; d++) {
                if (!(bits & (d ? MP2V_MB_BWD : MP2V_MB_FWD))) continue;
                const int fr = d ? pp.l1_frame : pp.l0_frame;
                if (fr < 0) return bad(MP2V_ERR_STATE, "prediction from a missing reference frame");
                const int mvx = r.mv[d][0], mvy = r.mv[d][1];
                const int x0 = mbx * 16 + (mvx >> 1), y0 = mby * 16 + (mvy >> 1);
                // the reference does not clamp: a vector leaving the frame is rejected here
                if (x0 < 0 || y0 < 0 || x0 + 16 + (mvx & 1) > W || y0 + 16 + (mvy & 1) > H)
                    return bad(MP2V_ERR_RANGE, "motion vector points outside the reference frame");
            }
        }
    }
    s.alg_bytes = out_bytes + ref + 128u * coded + 16u * (uint64_t)ctx->mb_count;
    return MP2V_OK;
}

// ---------------------------------------------------------------------------------------------
// device-side slice parsing

// Fold the parse results that have arrived into the statistics and the sticky error (ctx->mu held).
static void fold_status(mp2v_recon* ctx, slot_t& s) {
    s.status_pending = false;
    uint64_t n_coef = 0, coded = 0, dirs = 0;
    for (int i = 0; i < s.n_slices; i++) {
        const vlc_slice_status_t& st = s.h_status[i];
        n_coef += st.n_coef; coded += st.coded_blocks; dirs += st.ref_dirs;
        if (st.error && ctx->vlc_error.empty())
            ctx->vlc_error = "picture " + std::to_string(s.picture_no) + " slice " + std::to_string(i) + ": " + slice_error_string((int)st.error);
    }
    const uint64_t mb_bytes = ctx->cfg.chroma_format == 1 ? 384 : ctx->cfg.chroma_format == 2 ? 512 : 768;
    uint64_t out_bytes = 0;
    for (int p = 0; p < 3; p++) out_bytes += (uint64_t)ctx->lay.width[p] * ctx->lay.height[p];
    ctx->stats.algorithmic_bytes += out_bytes + dirs * mb_bytes + 128u * coded + 16u * (uint64_t)ctx->mb_count;
    ctx->stats.vlc_coefs += n_coef;
}
// Parses finish (almost) in submission order: fold from the front of the queue and stop at the first
// one still running, so a call costs one event query, not one per slot.
static void harvest_all(mp2v_recon* ctx) {
    while (!ctx->status_fifo.empty()) {
        slot_t& s = ctx->slots[ctx->status_fifo.front()];
        if (!s.vlc_wait || cudaEventQuery(s.vlc_wait) != cudaSuccess) break;      // (nullptr: its batched parse has not been launched yet)
        fold_status(ctx, s);
        ctx->status_fifo.pop_front();
    }
}
#define CHECK_VLC_ERROR() do { if (!ctx->vlc_error.empty()) return ctx->fail(MP2V_ERR_RANGE, ctx->vlc_error); } while (0)

extern "C" MP2V_API int mp2v_recon_acquire_picture(mp2v_recon_t* ctx, mp2v_picture_t** out) {
    if (!ctx || !out) return MP2V_ERR_ARG;
    for (int attempt = 0; attempt < 4; attempt++) {
        cudaEvent_t wait_for = nullptr;
        int candidate = -1;
        {
            std::lock_guard<std::mutex> lk(ctx->mu);
            int best = -1;
            for (size_t i = 0; i < ctx->slots.size(); i++)
                if (ctx->slots[i].state == SLOT_FREE) { best = (int)i; break; }
            // the oldest in-flight slot: as good as a free one if its launch already finished (one event query, not one per slot)
            for (size_t i = 0; i < ctx->slots.size() && best < 0; i++) {
                slot_t& s = ctx->slots[i];
                if (s.state == SLOT_INFLIGHT && (candidate < 0 || s.seq < ctx->slots[candidate].seq)) candidate = (int)i;
            }
            if (best < 0 && candidate >= 0 && cudaEventQuery(ctx->slots[candidate].done) == cudaSuccess) best = candidate;
            if (best >= 0) {
                slot_t& s = ctx->slots[best];
                harvest_all(ctx);        // a finished launch implies every parse queued before it has finished
                CHECK_VLC_ERROR();
                s.state = SLOT_FILLING;
                s.prechecked = false;
                s.vlc = false;
                s.stream_pic = false;
                s.staged = false;
                memset(s.pub.params, 0, sizeof(mp2v_pic_params_t));
                s.pub.params->l0_frame = s.pub.params->l1_frame = -1;
                *out = &s.pub;
                return MP2V_OK;
            }
            // otherwise wait (outside the lock) for that oldest in-flight slot; launch queued work first
            if (candidate < 0) {
                if (ctx->queued == 0) return ctx->fail(MP2V_ERR_STATE, "no picture slot available (all slots are being filled or resident)");
                const int rc = flush_locked(ctx);
                if (rc != MP2V_OK) return rc;
                continue;
            }
            wait_for = ctx->slots[candidate].done;
        }
        const cudaError_t e = cudaEventSynchronize(wait_for);
        if (e != cudaSuccess) { std::lock_guard<std::mutex> lk(ctx->mu); return ctx->cuda_fail(e, "event sync"); }
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    return ctx->fail(MP2V_ERR_STATE, "no picture slot became available");
}

static slot_t* slot_of(mp2v_recon* ctx, mp2v_picture_t* pic) {
    if (!pic || pic->slot < 0 || pic->slot >= (int)ctx->slots.size() || &ctx->slots[pic->slot].pub != pic) return nullptr;
    return &ctx->slots[pic->slot];
}

extern "C" MP2V_API int mp2v_recon_release_picture(mp2v_recon_t* ctx, mp2v_picture_t* pic) {
    if (!ctx) return MP2V_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    slot_t* s = slot_of(ctx, pic);
    if (!s) return ctx->fail(MP2V_ERR_ARG, "not a picture of this context");
    if (s->state == SLOT_QUEUED) { const int rc = flush_locked(ctx); if (rc != MP2V_OK) return rc; }
    if (s->state == SLOT_INFLIGHT || s->state == SLOT_RESIDENT) {
        CK(cudaSetDevice(ctx->cfg.device), "cudaSetDevice");
        CK(cudaStreamSynchronize(ctx->s_compute), "stream sync");
    }
    s->state = SLOT_FREE;
    return MP2V_OK;
}

// a reference must have been written by a launched picture or be the destination of a queued one; ctx->mu held
static bool references_available(const mp2v_recon* ctx, const mp2v_pic_params_t& pp) {
    for (int d = 0; d < 2; d++) {
        const int fr = d ? pp.l1_frame : pp.l0_frame;
        if (fr < 0 || ctx->frame_written[fr] || frame_is_queued(ctx, fr)) continue;
        return false;
    }
    return true;
}

// Queue a slot whose records are (or will be) complete; ctx->mu held.  Everything queued is launched once a lot
// of pictures waits: `max_batch` of them -- 1, 2, 4, ... right after a sync, so that the first frames of a decode
// leave early -- or by flush / sync / a wait for one of their frames.
static int queue_slot(mp2v_recon* ctx, slot_t* s) {
    const mp2v_pic_params_t& pp = *s->pub.params;
    if (!references_available(ctx, pp)) return ctx->fail(MP2V_ERR_STATE, "reference frame has never been written");
    s->state = SLOT_QUEUED;
    s->seq = ++ctx->seq;
    s->picture_no = ctx->pictures_submitted++;
    ctx->pending.push_back(s->pub.slot);
    ctx->queued = (int)ctx->pending.size();
    // Device-parsed stream pictures are launched in larger lots (4, 8, ... kMaxStreamBatch): a slice parses at the speed of
    // ONE thread (about a millisecond for a dense 1080p row), so the parser's throughput is the number of pictures in flight.
    // MP2V_RECON_THROUGHPUT: full lots from the first picture on (nobody is waiting for the first frame).
    // Otherwise lots of 4, 8, 16, 32, 32, ...: the frame copies to the host (the slower stage) start early and never run dry
    // (measured: a last lot of 60 pictures left the copy engine idle for 2 ms of a 12 ms decode).
    const bool rate = (ctx->cfg.flags & MP2V_RECON_THROUGHPUT) != 0;
    const int ramp = rate ? kMaxBatch : ctx->batch_ramp;
    const int quota = s->stream_pic ? std::min(ctx->lot_cap ? ctx->lot_cap : (rate ? 64 : 32), 4 * ramp) : std::min(ctx->max_batch, ramp);
    // (a smaller first lot in throughput mode measured slower: 4.3 ms instead of 4.0 ms per 120-picture call)
    if (ctx->queued >= quota) return flush_locked(ctx);
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_recon_submit(mp2v_recon_t* ctx, mp2v_picture_t* pic) {
    if (!ctx) return MP2V_ERR_ARG;
    slot_t* s = slot_of(ctx, pic);
    if (!s || s->state != SLOT_FILLING) { std::lock_guard<std::mutex> lk(ctx->mu); return ctx->fail(MP2V_ERR_STATE, "submit: picture was not acquired"); }
    if (ctx->vlc) { std::lock_guard<std::mutex> lk(ctx->mu); return ctx->fail(MP2V_ERR_STATE, "submit: this context parses on the device (MP2V_RECON_DEVICE_VLC): use mp2v_recon_submit_slices"); }
    // the slot belongs to the caller until it is queued: validate its records without holding the lock
    std::string why;
    int rc = s->prechecked ? MP2V_OK : account_and_validate(ctx, *s, (ctx->cfg.flags & MP2V_RECON_VALIDATE) != 0, &why);
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (rc != MP2V_OK) return ctx->fail(rc, why);
    return queue_slot(ctx, s);
}

extern "C" MP2V_API int mp2v_recon_stage_slices(mp2v_recon_t* ctx, mp2v_picture_t* pic, const mp2v_pic_syntax_t* syntax,
                                                const mp2v_slice_ref_t* slices, int n_slices) {
    if (!ctx) return MP2V_ERR_ARG;
    auto fail = [&](int code, const char* what) { std::lock_guard<std::mutex> lk(ctx->mu); return ctx->fail(code, what); };
    slot_t* s = slot_of(ctx, pic);
    if (!s || s->state != SLOT_FILLING) return fail(MP2V_ERR_STATE, "stage_slices: picture was not acquired");
    if (!ctx->vlc) return fail(MP2V_ERR_STATE, "stage_slices: context was created without MP2V_RECON_DEVICE_VLC");
    if (!syntax || (n_slices > 0 && !slices) || n_slices < 0) return fail(MP2V_ERR_ARG, "stage_slices: bad arguments");
    const mp2v_pic_params_t& pp = *s->pub.params;
    const int nf = ctx->cfg.n_frames;
    if (pp.dst_frame < 0 || pp.dst_frame >= nf || pp.l0_frame >= nf || pp.l1_frame >= nf) return fail(MP2V_ERR_ARG, "frame id out of range");
    if (pp.picture_coding_type < 1 || pp.picture_coding_type > 3) return fail(MP2V_ERR_ARG, "picture_coding_type must be 1 (I), 2 (P) or 3 (B)");
    if ((pp.picture_coding_type >= 2 && pp.l0_frame < 0) || (pp.picture_coding_type == 3 && pp.l1_frame < 0))
        return fail(MP2V_ERR_STATE, "prediction from a missing reference frame");
    if (syntax->intra_dc_precision < 0 || syntax->intra_dc_precision > 3) return fail(MP2V_ERR_ARG, "intra_dc_precision out of range");
    if (n_slices > ctx->mbh) return fail(MP2V_ERR_RANGE, "the device parser takes at most one slice per macroblock row");
    // ---- stage header + slice table + bitstream (the slot is the caller's: no lock)
    if (!s->h_staged) {
        // first use of this slot by the staged front end: its pinned + device staging block, stream and event
        cudaError_t e = cudaSetDevice(ctx->cfg.device);
        if (e == cudaSuccess) e = cudaHostAlloc(&s->h_staged, ctx->staged_bytes, cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaMalloc(&s->d_staged, ctx->staged_bytes);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->s_vlc, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->vlc_done, cudaEventDisableTiming);
        if (e != cudaSuccess) { std::lock_guard<std::mutex> lk(ctx->mu); return ctx->cuda_fail(e, "bitstream staging buffers"); }
    }
    vlc_pic_header_t& hdr = *reinterpret_cast<vlc_pic_header_t*>(s->h_staged + kVlcParamsBytes);
    memset(&hdr, 0, sizeof(hdr));
    hdr.sx.picture_coding_type = pp.picture_coding_type;
    for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) hdr.sx.f_code[a][b] = syntax->f_code[a][b];
    hdr.sx.intra_dc_precision = syntax->intra_dc_precision;
    hdr.sx.q_scale_type = syntax->q_scale_type != 0;
    hdr.sx.intra_vlc_format = syntax->intra_vlc_format != 0;
    hdr.sx.chroma_format = ctx->cfg.chroma_format;
    hdr.sx.vertical_size = ctx->cfg.height;            // only compared with 2800 (slice_vertical_position_extension)
    hdr.sx.mbw = ctx->mbw; hdr.sx.mbh = ctx->mbh;
    hdr.sx.field_dct_syntax = syntax->field_dct_syntax != 0;
    hdr.n_slices = (uint32_t)n_slices;
    hdr.slice_region = ctx->slice_region;
    hdr.data_off = (uint32_t)((kVlcParamsBytes + sizeof(vlc_pic_header_t) + (size_t)ctx->mbh * sizeof(vlc_slice_t) + 15) & ~(size_t)15);
    vlc_slice_t* table = reinterpret_cast<vlc_slice_t*>(s->h_staged + kVlcParamsBytes + sizeof(vlc_pic_header_t));
    size_t staged_end = hdr.data_off;
    int rows_covered = 0;
    if (n_slices > 0) {
        // one copy of the byte range that covers every slice (slices of a picture are contiguous in a stream)
        const uint8_t* lo = slices[0].payload;
        const uint8_t* hi = slices[0].payload + slices[0].bytes;
        std::vector<uint8_t> row_seen((size_t)ctx->mbh, 0);
        for (int i = 0; i < n_slices; i++) {
            if (!slices[i].payload || slices[i].code < 1 || slices[i].code > 0xAF) return fail(MP2V_ERR_ARG, "stage_slices: bad slice reference");
            lo = std::min(lo, slices[i].payload);
            hi = std::max(hi, slices[i].payload + slices[i].bytes);
            int row = slices[i].code - 1;
            if (ctx->cfg.height > 2800 && slices[i].bytes > 0) row += (slices[i].payload[0] >> 5) << 7;
            if (row < 0 || row >= ctx->mbh) return fail(MP2V_ERR_RANGE, "slice row outside the picture");
            if (row_seen[row]) return fail(MP2V_ERR_RANGE, "the device parser takes at most one slice per macroblock row");
            row_seen[row] = 1;
            rows_covered++;
        }
        const size_t span = (size_t)(hi - lo);
        if (hdr.data_off + span + 16 > ctx->staged_bytes) return fail(MP2V_ERR_RANGE, "coded picture larger than bitstream_capacity");
        memcpy(s->h_staged + hdr.data_off, lo, span);
        memset(s->h_staged + hdr.data_off + span, 0, 16);
        for (int i = 0; i < n_slices; i++) { table[i].byte_off = (uint32_t)(slices[i].payload - lo); table[i].code = slices[i].code; }
        staged_end = hdr.data_off + span + 16;
    }
    s->pub.params->n_coef = 0;
    memcpy(s->h_staged, s->pub.params, sizeof(mp2v_pic_params_t));      // the parameters travel with the staged bitstream: one H2D per picture
    s->d_params = s->d_staged;
    s->n_slices = n_slices;
    s->staged_end = staged_end;
    s->rows_covered = rows_covered;
    s->staged = true;
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_recon_submit_staged(mp2v_recon_t* ctx, mp2v_picture_t* pic) {
    if (!ctx) return MP2V_ERR_ARG;
    slot_t* s = slot_of(ctx, pic);
    // ---- H2D + parse on the slot's own stream, then queue the reconstruction
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!s || s->state != SLOT_FILLING || !s->staged) return ctx->fail(MP2V_ERR_STATE, "submit_staged: picture was not staged with mp2v_recon_stage_slices");
    s->staged = false;
    const int n_slices = s->n_slices;
    const size_t staged_end = s->staged_end;
    const int rows_covered = s->rows_covered;
    CHECK_VLC_ERROR();
    // everything queue_slot can refuse is checked before any device work is issued for the picture
    if (!references_available(ctx, *s->pub.params)) { s->staged = true; return ctx->fail(MP2V_ERR_STATE, "reference frame has never been written"); }
    CK(cudaSetDevice(ctx->cfg.device), "cudaSetDevice");
    mp2v_mb_info_t* d_mb = reinterpret_cast<mp2v_mb_info_t*>(s->d_arena + kParamsBytes);
    s->trace_idx = -1;
    if (ctx->trace) {
        if (!ctx->trace_base) { CK(cudaEventCreate(&ctx->trace_base), "event"); CK(cudaEventRecord(ctx->trace_base, s->s_vlc), "event record"); }
        mp2v_recon::trace_rec_t tr{ctx->pictures_submitted, nullptr, nullptr, nullptr, nullptr};
        CK(cudaEventCreate(&tr.h2d), "event"); CK(cudaEventCreate(&tr.vlc), "event"); CK(cudaEventCreate(&tr.recon), "event"); CK(cudaEventCreate(&tr.d2h), "event");
        s->trace_idx = (int)ctx->trace_log.size();
        ctx->trace_log.push_back(tr);
    }
    CK(cudaMemcpyAsync(s->d_staged, s->h_staged, staged_end, cudaMemcpyHostToDevice, s->s_vlc), "H2D bitstream");
    if (s->trace_idx >= 0) CK(cudaEventRecord(ctx->trace_log[s->trace_idx].h2d, s->s_vlc), "event record");
    // the kernel blanks what a slice leaves uncoded in its own row; rows without any slice (never in valid streams) are blanked here
    if (rows_covered < ctx->mbh)
        CK(cudaMemcpyAsync(d_mb, ctx->d_blank_mb, (size_t)ctx->mb_count * sizeof(mp2v_mb_info_t), cudaMemcpyDeviceToDevice, s->s_vlc), "blank records");
    CK(launch_vlc(s->d_staged, ctx->d_tables, d_mb, reinterpret_cast<mp2v_coef_t*>(s->d_arena + ctx->coef_off), s->d_status, n_slices,
                  ctx->vlc_lanes, s->s_vlc), "slice parser kernel launch");
    if (s->trace_idx >= 0) CK(cudaEventRecord(ctx->trace_log[s->trace_idx].vlc, s->s_vlc), "event record");
    CK(cudaEventRecord(s->vlc_done, s->s_vlc), "event record");
    s->vlc_wait = s->vlc_done;
    ctx->stats.h2d_bytes += staged_end;
    ctx->stats.d2h_bytes += (uint64_t)n_slices * sizeof(vlc_slice_status_t);
    if (n_slices > 0) { ctx->stats.vlc_launches += 1; ctx->stats.vlc_slices += (uint64_t)n_slices; }
    s->vlc = true;
    s->status_pending = true;
    ctx->status_fifo.push_back(s->pub.slot);
    s->alg_bytes = 0;                                    // folded in from the parse status (fold_status)
    return queue_slot(ctx, s);
}

extern "C" MP2V_API int mp2v_recon_submit_slices(mp2v_recon_t* ctx, mp2v_picture_t* pic, const mp2v_pic_syntax_t* syntax,
                                                 const mp2v_slice_ref_t* slices, int n_slices) {
    const int rc = mp2v_recon_stage_slices(ctx, pic, syntax, slices, n_slices);
    return rc != MP2V_OK ? rc : mp2v_recon_submit_staged(ctx, pic);
}

// ---------------------------------------------------------------------------------------------
// stream-resident front end

// waits for the scan launched by stream_begin and hands out its list (ctx->mu held)
static int fetch_codes_locked(mp2v_recon* ctx, const uint32_t** codes, uint32_t* n_codes) {
    if (!ctx->scan_pending) return ctx->fail(MP2V_ERR_STATE, "stream_codes: no start-code scan is pending (mp2v_recon_stream_begin with scan)");
    ctx->scan_pending = false;
    CK(cudaSetDevice(ctx->cfg.device), "cudaSetDevice");
    CK(cudaStreamSynchronize(ctx->s_parse[0]), "stream sync");
    const uint32_t total = *ctx->h_total;
    if (total > ctx->codes_cap) return ctx->fail(MP2V_ERR_RANGE, "stream_begin: more start codes than the scan list holds");
    if (total) CK(cudaMemcpy(ctx->h_codes, ctx->d_codes, (size_t)total * sizeof(uint32_t), cudaMemcpyDeviceToHost), "D2H start codes");
    ctx->stats.d2h_bytes += (uint64_t)total * sizeof(uint32_t);
    *codes = ctx->h_codes;
    *n_codes = total;
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_recon_stream_begin(mp2v_recon_t* ctx, const uint8_t* data, size_t bytes, const mp2v_byte_range_t* ranges, int n_ranges,
                                                int scan, const uint32_t** codes, uint32_t* n_codes) {
    if (!ctx || (!data && bytes) || n_ranges < 0 || (n_ranges && !ranges) || scan < 0 || scan > 2 || (scan == 1 && (!codes || !n_codes))) return MP2V_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->vlc) return ctx->fail(MP2V_ERR_STATE, "stream_begin: context was created without MP2V_RECON_DEVICE_VLC");
    if (bytes > 0x7ffffff0ull) return ctx->fail(MP2V_ERR_ARG, "stream_begin: stream too large");
    for (int i = 0; i < n_ranges; i++)
        if (ranges[i].offset > bytes || ranges[i].bytes > bytes - ranges[i].offset) return ctx->fail(MP2V_ERR_ARG, "stream_begin: range outside the stream");
    if (scan && n_ranges && ranges[0].bytes && (ranges[0].offset & 15)) return ctx->fail(MP2V_ERR_ARG, "stream_begin: the scanned range must start on a 16-byte boundary");
    int rc = flush_locked(ctx);
    if (rc != MP2V_OK) return rc;
    CK(cudaSetDevice(ctx->cfg.device), "cudaSetDevice");
    // parses of the previous stream may still be reading it
    for (auto st : ctx->s_parse) if (st) CK(cudaStreamSynchronize(st), "stream sync");
    ctx->scan_pending = false;
    const size_t need = ((bytes + 4095) & ~(size_t)4095) + 4096;         // the scan reads whole 4 KiB chunks + look-ahead; parsers read a few bytes past a slice
    if (need > ctx->stream_cap) {
        if (ctx->d_stream) CK(cudaFree(ctx->d_stream), "cudaFree");
        ctx->d_stream = nullptr; ctx->stream_cap = 0;
        const size_t cap = need + need / 4;
        CK(cudaMalloc(&ctx->d_stream, cap), "cudaMalloc stream");
        ctx->stream_cap = cap;
    }
    cudaStream_t st = ctx->s_parse[0];
    if (ctx->trace && !ctx->trace_base) { CK(cudaEventCreate(&ctx->trace_base), "event"); CK(cudaEventRecord(ctx->trace_base, st), "event record"); }
    // what is scanned: the whole stream, or the first range (a device's part of a stream whose scan is shared among devices)
    const size_t scan_len = n_ranges ? ranges[0].bytes : bytes, scan_lo = (n_ranges && scan_len) ? ranges[0].offset : 0;
    if (n_ranges == 0) {
        if (bytes) CK(cudaMemcpyAsync(ctx->d_stream, data, bytes, cudaMemcpyHostToDevice, st), "H2D stream");
        ctx->stats.h2d_bytes += bytes;
    } else {
        for (int i = 0; i < n_ranges; i++) {
            // a start code that begins in the last two bytes of the scanned range ends behind it: those bytes come along
            const size_t n = (i == 0 && scan) ? std::min(bytes - ranges[i].offset, ranges[i].bytes + 32) : ranges[i].bytes;
            if (n) CK(cudaMemcpyAsync(ctx->d_stream + ranges[i].offset, data + ranges[i].offset, n, cudaMemcpyHostToDevice, st), "H2D stream");
            ctx->stats.h2d_bytes += n;
        }
    }
    CK(cudaMemsetAsync(ctx->d_stream + bytes, 0, need - bytes, st), "memset stream tail");
    ctx->stream_len = bytes;
    ctx->h_stream = data;
    if (scan) {
        const size_t nblk = vlc_scan_blocks(scan_len);
        if (nblk + 1 > ctx->counts_cap) {
            if (ctx->d_counts) CK(cudaFree(ctx->d_counts), "cudaFree");
            ctx->d_counts = nullptr; ctx->counts_cap = 0;
            CK(cudaMalloc(&ctx->d_counts, (nblk + 1 + 1024) * sizeof(uint32_t)), "cudaMalloc scan scratch");
            ctx->counts_cap = nblk + 1 + 1024;
        }
        // a code every 64 bytes on average is far beyond any real stream (a 1080p slice is kilobytes); more than that takes the host path
        const uint32_t cap = (uint32_t)std::max<size_t>(1u << 16, scan_len / 64);
        if (cap > ctx->codes_cap) {
            if (ctx->d_codes) CK(cudaFree(ctx->d_codes), "cudaFree");
            if (ctx->h_codes) CK(cudaFreeHost(ctx->h_codes), "cudaFreeHost");
            ctx->d_codes = nullptr; ctx->h_codes = nullptr; ctx->codes_cap = 0;
            CK(cudaMalloc(&ctx->d_codes, (size_t)cap * sizeof(uint32_t)), "cudaMalloc start codes");
            CK(cudaHostAlloc(&ctx->h_codes, (size_t)cap * sizeof(uint32_t), cudaHostAllocDefault), "cudaHostAlloc start codes");
            ctx->codes_cap = cap;
        }
        CK(launch_start_code_scan(ctx->d_stream + scan_lo, scan_len, (uint32_t)scan_lo, ctx->d_counts, ctx->d_codes, ctx->codes_cap, ctx->d_total, st), "start code scan");
        ctx->scan_pending = true;
    }
    CK(cudaEventRecord(ctx->ev_stream, st), "event record");
    if (ctx->trace) {
        if (!ctx->ev_stream_timed) CK(cudaEventCreate(&ctx->ev_stream_timed), "event");
        CK(cudaEventRecord(ctx->ev_stream_timed, st), "event record");
    }
    if (scan == 1) return fetch_codes_locked(ctx, codes, n_codes);
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_recon_stream_codes(mp2v_recon_t* ctx, const uint32_t** codes, uint32_t* n_codes) {
    if (!ctx || !codes || !n_codes) return MP2V_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return fetch_codes_locked(ctx, codes, n_codes);
}

extern "C" MP2V_API int mp2v_recon_stream_add(mp2v_recon_t* ctx, const mp2v_byte_range_t* ranges, int n_ranges) {
    if (!ctx || n_ranges < 0 || (n_ranges && !ranges)) return MP2V_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->vlc || !ctx->d_stream || !ctx->h_stream) return ctx->fail(MP2V_ERR_STATE, "stream_add: no resident stream (mp2v_recon_stream_begin)");
    for (int i = 0; i < n_ranges; i++)
        if (ranges[i].offset > ctx->stream_len || ranges[i].bytes > ctx->stream_len - ranges[i].offset) return ctx->fail(MP2V_ERR_ARG, "stream_add: range outside the stream");
    CK(cudaSetDevice(ctx->cfg.device), "cudaSetDevice");
    cudaStream_t st = ctx->s_parse[0];
    for (int i = 0; i < n_ranges; i++) {
        if (ranges[i].bytes) CK(cudaMemcpyAsync(ctx->d_stream + ranges[i].offset, ctx->h_stream + ranges[i].offset, ranges[i].bytes, cudaMemcpyHostToDevice, st), "H2D stream");
        ctx->stats.h2d_bytes += ranges[i].bytes;
    }
    CK(cudaEventRecord(ctx->ev_stream, st), "event record");      // parse launches from now on wait for these bytes too
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_recon_submit_stream_picture(mp2v_recon_t* ctx, mp2v_picture_t* pic, const mp2v_pic_syntax_t* syntax,
                                                         const uint32_t* slice_offsets, int n_slices) {
    if (!ctx) return MP2V_ERR_ARG;
    slot_t* s = slot_of(ctx, pic);
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!s || s->state != SLOT_FILLING) return ctx->fail(MP2V_ERR_STATE, "submit_stream_picture: picture was not acquired");
    if (!ctx->vlc || !ctx->d_stream) return ctx->fail(MP2V_ERR_STATE, "submit_stream_picture: no resident stream (mp2v_recon_stream_begin)");
    if (!syntax || (n_slices > 0 && !slice_offsets) || n_slices < 0) return ctx->fail(MP2V_ERR_ARG, "submit_stream_picture: bad arguments");
    const mp2v_pic_params_t& pp = *s->pub.params;
    const int nf = ctx->cfg.n_frames;
    if (pp.dst_frame < 0 || pp.dst_frame >= nf || pp.l0_frame >= nf || pp.l1_frame >= nf) return ctx->fail(MP2V_ERR_ARG, "frame id out of range");
    if (pp.picture_coding_type < 1 || pp.picture_coding_type > 3) return ctx->fail(MP2V_ERR_ARG, "picture_coding_type must be 1 (I), 2 (P) or 3 (B)");
    if ((pp.picture_coding_type >= 2 && pp.l0_frame < 0) || (pp.picture_coding_type == 3 && pp.l1_frame < 0))
        return ctx->fail(MP2V_ERR_STATE, "prediction from a missing reference frame");
    if (syntax->intra_dc_precision < 0 || syntax->intra_dc_precision > 3) return ctx->fail(MP2V_ERR_ARG, "intra_dc_precision out of range");
    if (n_slices > ctx->mbh) return ctx->fail(MP2V_ERR_RANGE, "the device parser takes at most one slice per macroblock row");
    CHECK_VLC_ERROR();
    if (!references_available(ctx, pp)) return ctx->fail(MP2V_ERR_STATE, "reference frame has never been written");
    if ((int)ctx->parse_pending.size() >= kMaxStreamBatch) { const int rc = flush_locked(ctx); if (rc != MP2V_OK) return rc; }
    // ---- the picture's descriptor, in the buffer of the next parse launch
    vlc_stream_pic_t& d = *reinterpret_cast<vlc_stream_pic_t*>(ctx->parse_buf[ctx->parse_buf_rr].h + ctx->parse_pending.size() * ctx->desc_stride);
    int rows_covered = 0;
    {
        uint64_t seen[16] = {};                               // mbh <= 1024 rows
        if (ctx->mbh > 1024) return ctx->fail(MP2V_ERR_RANGE, "picture too tall for the device parser");
        for (int i = 0; i < n_slices; i++) {
            const uint32_t off = slice_offsets[i];
            if ((size_t)off + 8 > ctx->stream_len) return ctx->fail(MP2V_ERR_ARG, "submit_stream_picture: slice offset outside the stream");
            const uint8_t* sc = ctx->h_stream + off;
            if (sc[0] != 0 || sc[1] != 0 || sc[2] != 1 || sc[3] < 1 || sc[3] > 0xAF) return ctx->fail(MP2V_ERR_ARG, "submit_stream_picture: not a slice start code");
            int row = sc[3] - 1;
            if (ctx->cfg.height > 2800) row += (sc[4] >> 5) << 7;
            if (row < 0 || row >= ctx->mbh) return ctx->fail(MP2V_ERR_RANGE, "slice row outside the picture");
            if (seen[row >> 6] >> (row & 63) & 1) return ctx->fail(MP2V_ERR_RANGE, "the device parser takes at most one slice per macroblock row");
            seen[row >> 6] |= 1ull << (row & 63);
            rows_covered++;
            d.slice_off[i] = off;
        }
    }
    d.params = pp;
    d.params.n_coef = 0;
    memset(&d.sx, 0, sizeof(d.sx));
    d.sx.picture_coding_type = pp.picture_coding_type;
    for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) d.sx.f_code[a][b] = syntax->f_code[a][b];
    d.sx.intra_dc_precision = syntax->intra_dc_precision;
    d.sx.q_scale_type = syntax->q_scale_type != 0;
    d.sx.intra_vlc_format = syntax->intra_vlc_format != 0;
    d.sx.chroma_format = ctx->cfg.chroma_format;
    d.sx.vertical_size = ctx->cfg.height;                 // only compared with 2800 (slice_vertical_position_extension)
    d.sx.mbw = ctx->mbw; d.sx.mbh = ctx->mbh;
    d.sx.field_dct_syntax = syntax->field_dct_syntax != 0;
    d.n_slices = (uint32_t)n_slices;
    d.slice_region = ctx->slice_region;
    d.params_out = reinterpret_cast<mp2v_pic_params_t*>(s->d_arena);
    s->d_params = s->d_arena;
    d.mb = reinterpret_cast<mp2v_mb_info_t*>(s->d_arena + kParamsBytes);
    d.coef = reinterpret_cast<mp2v_coef_t*>(s->d_arena + ctx->coef_off);
    d.status = s->d_status;
    s->n_slices = n_slices;
    s->need_blank = rows_covered < ctx->mbh;
    s->vlc = true;
    s->stream_pic = true;
    s->vlc_wait = nullptr;
    s->status_pending = true;
    s->trace_idx = -1;
    s->alg_bytes = 0;                                     // folded in from the parse status (fold_status)
    ctx->status_fifo.push_back(s->pub.slot);
    ctx->parse_pending.push_back(s->pub.slot);
    ctx->stats.d2h_bytes += (uint64_t)n_slices * sizeof(vlc_slice_status_t);
    ctx->stats.vlc_slices += (uint64_t)n_slices;
    return queue_slot(ctx, s);
}

extern "C" MP2V_API int mp2v_recon_precheck(mp2v_recon_t* ctx, mp2v_picture_t* pic) {
    if (!ctx) return MP2V_ERR_ARG;
    slot_t* s = slot_of(ctx, pic);
    if (!s || s->state != SLOT_FILLING) { std::lock_guard<std::mutex> lk(ctx->mu); return ctx->fail(MP2V_ERR_STATE, "precheck: picture was not acquired"); }
    std::string why;
    const int rc = account_and_validate(ctx, *s, (ctx->cfg.flags & MP2V_RECON_VALIDATE) != 0, &why);
    if (rc != MP2V_OK) { std::lock_guard<std::mutex> lk(ctx->mu); return ctx->fail(rc, why); }
    s->prechecked = true;
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_recon_numa_node(mp2v_recon_t* ctx) { return ctx ? ctx->numa_node : -1; }

extern "C" MP2V_API int mp2v_numa_parse_cpu_list(const char* list, int32_t* cpus, int cap) {
    if (!list) return 0;
    const std::vector<int> v = parse_cpu_list(list);
    for (size_t i = 0; i < v.size() && (int)i < cap && cpus; i++) cpus[i] = v[i];
    return (int)v.size();
}

extern "C" MP2V_API int mp2v_recon_flush(mp2v_recon_t* ctx) {
    if (!ctx) return MP2V_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return flush_locked(ctx);
}

extern "C" MP2V_API int mp2v_recon_sync(mp2v_recon_t* ctx) {
    if (!ctx) return MP2V_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int rc = flush_locked(ctx);
    if (rc != MP2V_OK) return rc;
    CK(cudaSetDevice(ctx->cfg.device), "cudaSetDevice");
    CK(cudaStreamSynchronize(ctx->s_copy), "stream sync");
    CK(cudaStreamSynchronize(ctx->s_compute), "stream sync");
    CK(cudaStreamSynchronize(ctx->s_d2h), "stream sync");      // queued frame copies (auto download, downloads of other threads)
    CK(cudaStreamSynchronize(ctx->s_d2h2), "stream sync");
    for (auto& b : ctx->parse_buf) b.used = false;
    std::fill(ctx->read_pending.begin(), ctx->read_pending.end(), 0);
    for (auto& s : ctx->slots) if (s.state == SLOT_INFLIGHT) s.state = SLOT_FREE;
    ctx->batch_ramp = 1;
    ctx->pictures_submitted = 0;                         // slice errors name pictures by their number since the last sync
    if (ctx->trace && !ctx->parse_log.empty()) {
        float up = -1;
        if (ctx->ev_stream_timed) cudaEventElapsedTime(&up, ctx->trace_base, ctx->ev_stream_timed);
        fprintf(stderr, "[mp2v trace] resident stream uploaded + scanned at %7.3f ms\n", up);
        int i = 0;
        for (auto& lt : ctx->parse_log) {
            float a = -1, b = -1;
            cudaEventElapsedTime(&a, ctx->trace_base, lt.begin); cudaEventElapsedTime(&b, ctx->trace_base, lt.end);
            fprintf(stderr, "[mp2v trace] parse  %3d  %3d pictures  kernel %7.3f .. %7.3f ms\n", i++, lt.n, a, b);
            cudaEventDestroy(lt.begin); cudaEventDestroy(lt.end);
        }
        ctx->parse_log.clear();
    }
    if (ctx->trace && !ctx->launch_log.empty()) {
        int i = 0;
        for (auto& lt : ctx->launch_log) {
            float a = -1, b = -1, c = -1;
            cudaEventElapsedTime(&a, ctx->trace_base, lt.begin); cudaEventElapsedTime(&b, ctx->trace_base, lt.end);
            if (cudaEventQuery(lt.copied) == cudaSuccess) cudaEventElapsedTime(&c, ctx->trace_base, lt.copied);
            fprintf(stderr, "[mp2v trace] launch %3d  %3d pictures  kernel %7.3f .. %7.3f  frames on the host %7.3f ms\n", i++, lt.n, a, b, c);
            cudaEventDestroy(lt.begin); cudaEventDestroy(lt.end); cudaEventDestroy(lt.copied);
        }
        cudaGetLastError();
        ctx->launch_log.clear();
        if (ctx->trace_log.empty()) { cudaEventDestroy(ctx->trace_base); ctx->trace_base = nullptr; }
    }
    if (ctx->trace && !ctx->trace_log.empty()) {
        cudaStreamSynchronize(ctx->s_d2h);
        for (auto& tr : ctx->trace_log) {
            float a = -1, b = -1, c = -1, d = -1;
            cudaEventElapsedTime(&a, ctx->trace_base, tr.h2d); cudaEventElapsedTime(&b, ctx->trace_base, tr.vlc);
            cudaEventElapsedTime(&c, ctx->trace_base, tr.recon);
            if (cudaEventQuery(tr.d2h) == cudaSuccess) cudaEventElapsedTime(&d, ctx->trace_base, tr.d2h);
            fprintf(stderr, "[mp2v trace] pic %3llu  h2d %7.3f  vlc %7.3f  recon %7.3f  d2h %7.3f ms\n", (unsigned long long)tr.picture_no, a, b, c, d);
            cudaEventDestroy(tr.h2d); cudaEventDestroy(tr.vlc); cudaEventDestroy(tr.recon); cudaEventDestroy(tr.d2h);
        }
        cudaGetLastError();   // an event never recorded (no download) reports an error that is not ours
        ctx->trace_log.clear();
        cudaEventDestroy(ctx->trace_base);
        ctx->trace_base = nullptr;
    }
    harvest_all(ctx);      // every parse precedes a reconstruction launch that has now finished
    if (!ctx->vlc_error.empty()) {
        const std::string why = ctx->vlc_error;
        ctx->vlc_error.clear();                          // reported once; the context stays usable
        return ctx->fail(MP2V_ERR_RANGE, why);
    }
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_recon_reset(mp2v_recon_t* ctx) {
    if (!ctx) return MP2V_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->cfg.device), "cudaSetDevice");
    // whatever was issued runs to completion (records of a failed picture are garbage-safe by construction)
    for (auto& s : ctx->slots) if (s.s_vlc) CK(cudaStreamSynchronize(s.s_vlc), "stream sync");
    CK(cudaStreamSynchronize(ctx->s_copy), "stream sync");
    CK(cudaStreamSynchronize(ctx->s_compute), "stream sync");
    CK(cudaStreamSynchronize(ctx->s_d2h), "stream sync");
    CK(cudaStreamSynchronize(ctx->s_d2h2), "stream sync");
    for (auto st : ctx->s_parse) if (st) CK(cudaStreamSynchronize(st), "stream sync");
    ctx->pending.clear();
    ctx->queued = 0;
    ctx->parse_pending.clear();
    for (auto& b : ctx->parse_buf) b.used = false;
    ctx->status_fifo.clear();
    for (auto& s : ctx->slots) { s.state = SLOT_FREE; s.staged = false; s.status_pending = false; s.prechecked = false; s.vlc = false; s.stream_pic = false; }
    std::fill(ctx->frame_written.begin(), ctx->frame_written.end(), 0);
    std::fill(ctx->mirror_valid.begin(), ctx->mirror_valid.end(), 0);
    std::fill(ctx->read_pending.begin(), ctx->read_pending.end(), 0);
    ctx->vlc_error.clear();
    ctx->err.clear();
    ctx->batch_ramp = 1;
    ctx->pictures_submitted = 0;
    return MP2V_OK;
}

// ---------------------------------------------------------------------------------------------
// device-resident mode

extern "C" MP2V_API int mp2v_recon_upload(mp2v_recon_t* ctx, mp2v_picture_t* pic) {
    if (!ctx) return MP2V_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    slot_t* s = slot_of(ctx, pic);
    if (!s || (s->state != SLOT_FILLING && s->state != SLOT_RESIDENT)) return ctx->fail(MP2V_ERR_STATE, "upload: picture was not acquired");
    if (ctx->vlc) return ctx->fail(MP2V_ERR_STATE, "upload: this context parses on the device (MP2V_RECON_DEVICE_VLC)");
    std::string why;
    const int rc = account_and_validate(ctx, *s, (ctx->cfg.flags & MP2V_RECON_VALIDATE) != 0, &why);
    if (rc != MP2V_OK) return ctx->fail(rc, why);
    CK(cudaSetDevice(ctx->cfg.device), "cudaSetDevice");
    const size_t bytes = ctx->coef_off + (size_t)s->pub.params->n_coef * sizeof(mp2v_coef_t);
    CK(cudaMemcpyAsync(s->d_arena, s->h_arena, bytes, cudaMemcpyHostToDevice, ctx->s_copy), "H2D picture records");
    CK(cudaStreamSynchronize(ctx->s_copy), "stream sync");
    ctx->stats.h2d_bytes += bytes;
    s->state = SLOT_RESIDENT;
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_recon_run_resident(mp2v_recon_t* ctx, mp2v_picture_t* const* pics, const int32_t* levels, int n) {
    if (!ctx || !pics || n < 0) return MP2V_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = flush_locked(ctx);
    if (rc != MP2V_OK) return rc;
    CK(cudaSetDevice(ctx->cfg.device), "cudaSetDevice");
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) {
        slot_t* s = slot_of(ctx, pics[i]);
        if (!s || s->state != SLOT_RESIDENT) return ctx->fail(MP2V_ERR_STATE, "run_resident: picture is not resident");
        order[i] = i;
    }
    if (levels) std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return levels[a] < levels[b]; });
    int ids[kMaxBatch];
    int cnt = 0;
    for (int k = 0; k < n; k++) {
        const int i = order[k];
        ids[cnt++] = pics[i]->slot;
        const bool last_of_level = (k + 1 == n) || !levels || levels[order[k + 1]] != levels[i];
        if (cnt == ctx->max_batch || last_of_level) {
            rc = launch_slots(ctx, ids, cnt);
            if (rc != MP2V_OK) return rc;
            cnt = 0;
        }
    }
    return MP2V_OK;
}

// ---------------------------------------------------------------------------------------------
// frames

// Enqueue the D2H copies of a frame behind its last writer (call with ctx->mu held); the caller waits
// for *done outside the lock so that submissions and launches keep flowing while the copy runs.
static int enqueue_frame_copy(mp2v_recon* ctx, int frame_id, uint8_t* const dst[3], const int32_t dst_stride[3], cudaEvent_t* done) {
    if (frame_id < 0 || frame_id >= ctx->cfg.n_frames) return ctx->fail(MP2V_ERR_ARG, "frame id out of range");
    if (frame_is_queued(ctx, frame_id)) { const int rc = flush_locked(ctx); if (rc != MP2V_OK) return rc; }
    if (!ctx->frame_written[frame_id]) return ctx->fail(MP2V_ERR_STATE, "frame has never been written");
    CK(cudaSetDevice(ctx->cfg.device), "cudaSetDevice");
    CK(cudaStreamWaitEvent(ctx->s_d2h, ctx->frame_launch[frame_id] >= 0 ? ctx->launch_ev[ctx->frame_launch[frame_id]] : ctx->frame_ev[frame_id], 0), "stream wait");
    // destination laid out exactly like the device frame (the pinned mirrors are): one copy for all planes
    bool same = true;
    for (int p = 0; p < 3; p++)
        same = same && dst_stride[p] == ctx->lay.stride[p] && dst[p] == dst[0] + ctx->lay.plane_offset[p];
    if (same) {
        CK(cudaMemcpyAsync(dst[0], ctx->frame_ptr(frame_id, 0), ctx->lay.bytes, cudaMemcpyDeviceToHost, ctx->s_d2h), "D2H frame");
        ctx->stats.d2h_bytes += ctx->lay.bytes;
    } else {
        for (int p = 0; p < 3; p++) {
            CK(cudaMemcpy2DAsync(dst[p], (size_t)dst_stride[p], ctx->frame_ptr(frame_id, p), (size_t)ctx->lay.stride[p],
                                 (size_t)ctx->lay.width[p], (size_t)ctx->lay.height[p], cudaMemcpyDeviceToHost, ctx->s_d2h), "D2H frame");
            ctx->stats.d2h_bytes += (uint64_t)ctx->lay.width[p] * ctx->lay.height[p];
        }
    }
    *done = ctx->get_event();
    CK(cudaEventRecord(*done, ctx->s_d2h), "event record");
    CK(cudaEventRecord(ctx->read_ev[frame_id], ctx->s_d2h), "event record");
    ctx->read_pending[frame_id] |= 1;
    return MP2V_OK;
}

static int wait_frame_copy(mp2v_recon* ctx, cudaEvent_t done, bool pooled = true) {
    const cudaError_t e = cudaEventSynchronize(done);
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (pooled) ctx->ev_pool.push_back(done);
    if (e != cudaSuccess) return ctx->cuda_fail(e, "D2H frame");
    harvest_all(ctx);      // the frame's picture has been parsed by now: a slice error surfaces with its frame
    CHECK_VLC_ERROR();
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_recon_download_frame(mp2v_recon_t* ctx, int frame_id, uint8_t* const dst[3], const int32_t dst_stride[3]) {
    if (!ctx || !dst || !dst_stride) return MP2V_ERR_ARG;
    cudaEvent_t done = nullptr;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        const int rc = enqueue_frame_copy(ctx, frame_id, dst, dst_stride, &done);
        if (rc != MP2V_OK) return rc;
    }
    return wait_frame_copy(ctx, done);
}

extern "C" MP2V_API int mp2v_recon_map_frame(mp2v_recon_t* ctx, int frame_id, uint8_t* planes[3], int32_t strides[3]) {
    if (!ctx || !planes || !strides) return MP2V_ERR_ARG;
    cudaEvent_t done = nullptr;
    bool pooled = true;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (frame_id < 0 || frame_id >= ctx->cfg.n_frames) return ctx->fail(MP2V_ERR_ARG, "frame id out of range");
        if (!ctx->h_frames[frame_id]) {      // (contexts with MP2V_RECON_AUTO_DOWNLOAD own all mirrors from the start)
            CK(cudaSetDevice(ctx->cfg.device), "cudaSetDevice");
            CK(cudaHostAlloc(&ctx->h_frames[frame_id], ctx->lay.bytes, cudaHostAllocDefault), "cudaHostAlloc frame mirror");
        }
        for (int p = 0; p < 3; p++) { planes[p] = ctx->h_frames[frame_id] + ctx->lay.plane_offset[p]; strides[p] = ctx->lay.stride[p]; }
        if (frame_is_queued(ctx, frame_id)) { const int rc = flush_locked(ctx); if (rc != MP2V_OK) return rc; }
        if (ctx->auto_dl && ctx->mirror_valid[frame_id]) {
            done = ctx->mirror_ev[ctx->mirror_ev_of[frame_id]];   // the copy was queued with the picture's launch
            pooled = false;
        } else {
            const int rc = enqueue_frame_copy(ctx, frame_id, planes, strides, &done);
            if (rc != MP2V_OK) return rc;
        }
    }
    return wait_frame_copy(ctx, done, pooled);
}

extern "C" MP2V_API int mp2v_recon_upload_frame(mp2v_recon_t* ctx, int frame_id, const uint8_t* const src[3], const int32_t src_stride[3]) {
    if (!ctx || !src || !src_stride) return MP2V_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (frame_id < 0 || frame_id >= ctx->cfg.n_frames) return ctx->fail(MP2V_ERR_ARG, "frame id out of range");
    int rc = flush_locked(ctx);
    if (rc != MP2V_OK) return rc;
    CK(cudaSetDevice(ctx->cfg.device), "cudaSetDevice");
    CK(cudaStreamSynchronize(ctx->s_compute), "stream sync");
    for (int p = 0; p < 3; p++)
        CK(cudaMemcpy2D(ctx->frame_ptr(frame_id, p), (size_t)ctx->lay.stride[p], src[p], (size_t)src_stride[p],
                        (size_t)ctx->lay.width[p], (size_t)ctx->lay.height[p], cudaMemcpyHostToDevice), "H2D frame");
    ctx->frame_written[frame_id] = 1;
    ctx->mirror_valid[frame_id] = 0;
    ctx->frame_launch[frame_id] = -1;
    CK(cudaEventRecord(ctx->frame_ev[frame_id], ctx->s_compute), "event record");
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_recon_frame_device_ptrs(mp2v_recon_t* ctx, int frame_id, void* planes[3], int32_t strides[3]) {
    if (!ctx || !planes || !strides) return MP2V_ERR_ARG;
    if (frame_id < 0 || frame_id >= ctx->cfg.n_frames) { std::lock_guard<std::mutex> lk(ctx->mu); return ctx->fail(MP2V_ERR_ARG, "frame id out of range"); }
    for (int p = 0; p < 3; p++) { planes[p] = ctx->frame_ptr(frame_id, p); strides[p] = ctx->lay.stride[p]; }
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_recon_convert_frames(mp2v_recon_t* ctx, int format, const int32_t* frame_ids, void* const* dst_device, int n, int32_t dst_pitch) {
    if (!ctx || !frame_ids || !dst_device || n < 0) return MP2V_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int cf = ctx->cfg.chroma_format;
    if (format == MP2V_OUT_NV12 || format == MP2V_OUT_P010) { if (cf != 1) return ctx->fail(MP2V_ERR_ARG, "NV12 / P010 are 4:2:0 formats"); }
    else if (format == MP2V_OUT_UYVY) { if (cf != 2) return ctx->fail(MP2V_ERR_ARG, "UYVY is a 4:2:2 format"); }
    else return ctx->fail(MP2V_ERR_ARG, "unknown output format");
    const int row_bytes = ctx->cfg.width * (format == MP2V_OUT_NV12 ? 1 : 2);
    if (dst_pitch < row_bytes || (dst_pitch & 15)) return ctx->fail(MP2V_ERR_ARG, "destination must be 16-byte aligned with a pitch that is a multiple of 16 and >= the row bytes");
    for (int i = 0; i < n; i++) {
        if (frame_ids[i] < 0 || frame_ids[i] >= ctx->cfg.n_frames) return ctx->fail(MP2V_ERR_ARG, "frame id out of range");
        if (!dst_device[i] || ((uintptr_t)dst_device[i] & 15)) return ctx->fail(MP2V_ERR_ARG, "destination must be 16-byte aligned with a pitch that is a multiple of 16 and >= the row bytes");
        if (!ctx->frame_written[frame_ids[i]] && !frame_is_queued(ctx, frame_ids[i])) return ctx->fail(MP2V_ERR_STATE, "frame has never been written");
    }
    const int rc = flush_locked(ctx);
    if (rc != MP2V_OK) return rc;
    CK(cudaSetDevice(ctx->cfg.device), "cudaSetDevice");
    // on the compute stream: ordered behind the launches that write the frames and ahead of any that overwrites them
    for (int first = 0; first < n; first += kMaxConvertBatch) {
        convert_batch_t b{};
        b.n_frames = std::min(n - first, (int)kMaxConvertBatch);
        b.width = ctx->cfg.width; b.height = ctx->cfg.height;
        b.stride_y = ctx->lay.stride[0]; b.stride_c = ctx->lay.stride[1]; b.dst_pitch = dst_pitch;
        for (int i = 0; i < b.n_frames; i++)
            b.frame[i] = {ctx->frame_ptr(frame_ids[first + i], 0), ctx->frame_ptr(frame_ids[first + i], 1), ctx->frame_ptr(frame_ids[first + i], 2),
                          static_cast<uint8_t*>(dst_device[first + i])};
        CK(launch_convert(format, b, ctx->s_compute), "output conversion launch");
        ctx->stats.launches += 1;
    }
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_recon_convert_frames_nv12(mp2v_recon_t* ctx, const int32_t* frame_ids, void* const* dst_device, int n, int32_t dst_pitch) {
    return mp2v_recon_convert_frames(ctx, MP2V_OUT_NV12, frame_ids, dst_device, n, dst_pitch);
}

extern "C" MP2V_API int mp2v_recon_convert_frame_nv12(mp2v_recon_t* ctx, int frame_id, void* dst_device, int32_t dst_pitch) {
    const int32_t id = frame_id;
    void* const dst = dst_device;
    return mp2v_recon_convert_frames(ctx, MP2V_OUT_NV12, &id, &dst, 1, dst_pitch);
}

// Wait until a frame's reconstruction has finished (device consumers that read the planes with their own streams).
extern "C" MP2V_API int mp2v_recon_wait_frame(mp2v_recon_t* ctx, int frame_id) {
    if (!ctx) return MP2V_ERR_ARG;
    cudaEvent_t ev = nullptr;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (frame_id < 0 || frame_id >= ctx->cfg.n_frames) return ctx->fail(MP2V_ERR_ARG, "frame id out of range");
        if (frame_is_queued(ctx, frame_id)) { const int rc = flush_locked(ctx); if (rc != MP2V_OK) return rc; }
        if (!ctx->frame_written[frame_id]) return ctx->fail(MP2V_ERR_STATE, "frame has never been written");
        ev = ctx->frame_launch[frame_id] >= 0 ? ctx->launch_ev[ctx->frame_launch[frame_id]] : ctx->frame_ev[frame_id];
    }
    const cudaError_t e = cudaEventSynchronize(ev);
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (e != cudaSuccess) return ctx->cuda_fail(e, "event sync");
    harvest_all(ctx);      // the frame's picture has been parsed by now: a slice error surfaces with its frame
    CHECK_VLC_ERROR();
    return MP2V_OK;
}

// ---------------------------------------------------------------------------------------------
// statistics

extern "C" MP2V_API int mp2v_recon_set_timing(mp2v_recon_t* ctx, int enable) {
    if (!ctx) return MP2V_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->timing = enable != 0;
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_recon_get_stats(mp2v_recon_t* ctx, mp2v_recon_stats_t* out, int reset) {
    if (!ctx || !out) return MP2V_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->cfg.device), "cudaSetDevice");
    if (!ctx->timed.empty()) {
        CK(cudaStreamSynchronize(ctx->s_compute), "stream sync");
        for (auto& pr : ctx->timed) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, pr.first, pr.second), "event elapsed");
            ctx->stats.kernel_ms += ms;
            ctx->timing_pool.push_back(pr.first); ctx->timing_pool.push_back(pr.second);
        }
        ctx->timed.clear();
    }
    if (ctx->timing && ctx->d_counters) {
        unsigned long long c[3] = {0, 0, 0};
        CK(cudaStreamSynchronize(ctx->s_compute), "stream sync");
        CK(cudaMemcpy(c, ctx->d_counters, sizeof(c), cudaMemcpyDeviceToHost), "D2H counters");
        ctx->stats.idct_batches = c[0]; ctx->stats.idct_exact_pass2 = c[1]; ctx->stats.idct_exact_pass1 = c[2];
        if (reset) CK(cudaMemset(ctx->d_counters, 0, sizeof(c)), "cudaMemset counters");
    }
    *out = ctx->stats;
    if (reset) ctx->stats = mp2v_recon_stats_t{};
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_recon_timer_start(mp2v_recon_t* ctx) {
    if (!ctx) return MP2V_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int rc = flush_locked(ctx);
    if (rc != MP2V_OK) return rc;
    CK(cudaSetDevice(ctx->cfg.device), "cudaSetDevice");
    CK(cudaEventRecord(ctx->ev_t0, ctx->s_compute), "event record");
    return MP2V_OK;
}

extern "C" MP2V_API int mp2v_recon_timer_stop(mp2v_recon_t* ctx, double* elapsed_ms) {
    if (!ctx || !elapsed_ms) return MP2V_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int rc = flush_locked(ctx);
    if (rc != MP2V_OK) return rc;
    CK(cudaSetDevice(ctx->cfg.device), "cudaSetDevice");
    CK(cudaEventRecord(ctx->ev_t1, ctx->s_compute), "event record");
    CK(cudaEventSynchronize(ctx->ev_t1), "event sync");
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1), "event elapsed");
    *elapsed_ms = ms;
    return MP2V_OK;
}
