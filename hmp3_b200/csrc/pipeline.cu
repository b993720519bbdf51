// Host side of the device pipeline: batch plan (buffers sized for a batch shape), chunked kernel
// launches, and the C ABI of include/hmp3_b200.h.  No CPU fallback exists: without a CUDA device
// every entry point fails.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <array>
#include <chrono>
#include <limits>
#include <vector>

#include "../../include/hmp3_b200.h"
#include "../../include/hmp3_b200_debug.h"
#include "enc_init.h"
#include "analysis.h"
#include "batch_types.h"
#include "rate_driver.h"
#include "resample.h"

using namespace hmp3;

namespace {

thread_local std::string g_err;
thread_local int g_create_status = 0;  // why the last hmp3_batch_create* of this thread failed (HMP3_ERR_*)
void set_err(const std::string &s) { g_err = s; }

#define CK(call)                                                                               \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            set_err(std::string(#call) + ": " + cudaGetErrorString(e_));                       \
            return HMP3_ERR_CUDA;                                                              \
        }                                                                                      \
    } while (0)

bool same_control(const hmp3_control &a, const hmp3_control &b) { return memcmp(&a, &b, sizeof(a)) == 0; }

inline unsigned blocks_for(long long items, int bs) { return (unsigned)((items + bs - 1) / bs); }

// Which kernel runs the serial stage.  Phase-scheduled (kernels_rate_ph.cu) when the batch gives every SM enough
// streams to schedule by phase, the one-warp-per-stream kernel with the nested drivers otherwise (same bytes).
// Measured (tools/gpu_call_r3f.sh, x realtime nested / phased): 1250 streams 65.9k / 60.6k, 2500: 82.7k / 78.0k,
// 4736: equal, 9472: 87.5k / 116.7k.  HMP3_RATE_MODE=nested | phased overrides.
bool rate_mode_phased(int nstreams, int sm_count) {
    const char *e = getenv("HMP3_RATE_MODE");
    if (e && strcmp(e, "nested") == 0) return false;
    if (e && strcmp(e, "phased") == 0) return true;
    return nstreams > 24 * (sm_count > 0 ? sm_count : 148);
}

// The serial stage's hot per-stream state (RateState, ~7 KB per stream) is read and written by every granule of its
// stream, while ~14 KB of granule inputs and records per stream stream through L2 between two visits: left alone, the
// state is evicted and comes back from HBM every granule (profiles/r2o: L2 hit rate 60 %, long_scoreboard the top
// stall).  An access-policy window on the serial stage's CUDA stream marks the state array as persisting; the L2
// set-aside is HMP3_RATE_L2_PERSIST_MB (default 0 = off: see below).
void keep_rate_state_in_l2(cudaStream_t stream, void *base, size_t bytes, int device) {
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device);
    size_t want = 0;  // measured (profiles/r2p): DRAM writes of the serial stage halve, its time does not change -> off
    if (const char *e = getenv("HMP3_RATE_L2_PERSIST_MB")) want = (size_t)atoll(e) << 20;
    if (want == 0 || max_persist <= 0 || max_window <= 0) return;
    if (want > (size_t)max_persist) want = (size_t)max_persist;
    size_t cur = 0;
    cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
    if (cur < want && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    cudaStreamAttrValue a;
    memset(&a, 0, sizeof(a));
    a.accessPolicyWindow.base_ptr = base;
    a.accessPolicyWindow.num_bytes = bytes < (size_t)max_window ? bytes : (size_t)max_window;
    a.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)a.accessPolicyWindow.num_bytes);
    a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &a) != cudaSuccess) cudaGetLastError();
}

enum { PH_POLY = 0, PH_ATTACK, PH_SWITCH, PH_HYBRID, PH_PSY, PH_MS, PH_PSY2, PH_PREP, PH_RATE, PH_PACK, PH_ASSEMBLE, PH_COUNT };
const char *kPhaseNames[PH_COUNT] = {"polyphase", "attack", "switch_scan", "hybrid_mdct", "psy_stage1", "ms_scan",
                                     "psy_stage2", "prepare", "rate_loop", "pack", "assemble"};

}  // namespace

constexpr int kCycleLaunches = 64;
constexpr int kMaxSets = 3;

struct hmp3_batch {
    int device = 0;
    int n = 0;
    int NG = 0;
    int max_gran = 0;
    int max_frames = 0;
    std::vector<hmp3_control> controls;  // distinct
    std::vector<EncTables> tabs_h;
    std::vector<StreamDev> st_h;
    std::vector<StreamOut> so_h;
    std::vector<StreamResult> res_h;
    std::vector<long long> out_off_h;
    EncTables *d_tabs = nullptr;
    StreamDev *d_st = nullptr;
    StreamOut *d_so = nullptr;
    int16_t *d_pcm = nullptr;
    long long pcm_elems = 0;
    SwitchState *d_sw = nullptr;
    SwitchState *d_sw_init = nullptr;
    RateState *d_rs = nullptr;
    void *d_rs_cold = nullptr;   // RateCold[n]: short-block / CBitAllo1 state, kept out of the hot array
    unsigned char *d_main = nullptr;
    FrameRec *d_frames = nullptr;
    StreamResult *d_res = nullptr;
    long long *d_out_off = nullptr;
    unsigned char *d_out = nullptr;
    long long out_cap = 0;
    ChunkBufs cb2[kMaxSets]{};          // rotating chunk work areas: Phase A runs up to nbuf-1 chunks ahead of the serial stage
    ChunkBufs &cb = cb2[0];
    cudaStream_t stream = nullptr;      // serial stage, finish, copies
    bool one_stream = false;
    cudaStream_t stream_a = nullptr;    // Phase A
    cudaStream_t stream_p = nullptr;    // packing pass
    cudaStream_t stream_a0 = nullptr, stream_p0 = nullptr;  // the plan's own Phase A / packing streams (see hmp3_batch_set_serialize)
    cudaEvent_t ev_a[kMaxSets] = {nullptr, nullptr, nullptr}, ev_r[kMaxSets] = {nullptr, nullptr, nullptr},
                ev_p[kMaxSets] = {nullptr, nullptr, nullptr},
                ev_start = nullptr;
    const void *const *h_src = nullptr;  // callers' pinned PCM pointers of a staged run
    bool staged = false;                // this run copies PCM chunk by chunk (copy engine) ahead of each chunk's Phase A
    cudaStream_t stream_c = nullptr;    // H2D staging copies
    cudaEvent_t ev_c[kMaxSets] = {nullptr, nullptr, nullptr};
    float *d_pcmf = nullptr;            // float PCM: float inputs and the DC-filtered copies of the streams with -S1
    std::vector<int> fmt;               // per stream: 0 = int16 input, 1 = float32 input (scaled to +-32768)
    float *d_dc = nullptr;              // [n][2] filter state
    bool any_allo0 = false, any_allo1 = false;  // which serial-stage kernels the plan's streams need (CBitAllo3 / CBitAllo1)
    int poly_mode = 0;                  // 0 = exact polyphase; 3 / 1 = tensor-core contraction, 3xTF32 / TF32 (HMP3_POLY_MODE)
    float *d_polyw = nullptr;           // its coefficient blocks
    bool any_filter = false;
    int *d_msmem = nullptr;             // [n] M/S hysteresis memory (scan carry)
    PsyState *d_psy = nullptr;          // [n][2] psychoacoustic stage-2 carry
    int *d_flags = nullptr;             // [n] packing/accounting mismatch flags (must stay 0)
    // incremental output of the host entry (pinned out buffers): frames are assembled chunk by chunk into fixed
    // per-stream regions and copied out (DMA) while later chunks are still being encoded
    bool direct = false;
    std::vector<void *> cp_dst, cp_src;  // staging copies of one chunk (batched submission)
    std::vector<size_t> cp_len;
    bool no_batch_copy = false;
    unsigned char *d_out_inc = nullptr; // [out_cap] fixed regions: stream s at StreamDev::out_off
    int *d_done_lo = nullptr;           // [n] frames assembled so far
    long long *d_bytes_done = nullptr;  // [n] output bytes complete after the chunk
    long long *h_prog = nullptr;        // pinned [chunks][n] copies of d_bytes_done
    int h_prog_chunks = 0;
    std::vector<cudaEvent_t> ev_o;      // per chunk: its h_prog row has landed
    cudaStream_t stream_o = nullptr;    // D2H copies of finished output (copy engine)
    int chunks_run = 0;
    std::vector<std::array<float, 3>> timeline;
    int sm_count = 0;                   // multiprocessors of the plan's device
    int tap_stream = -1;                // hmp3_debug_set_rate_tap: stream whose pack records are copied out, -1 = off
    unsigned char *tap_out = nullptr;
    long long tap_cap = 0;
    long long *d_cycles = nullptr;      // [kCycleLaunches][n] serial-stage clocks per launch (diagnostics, on request)
    int cycle_launches = 0;
    std::vector<int> flags_h;
    int nbuf = 2;
    int launches = 0;
    bool results_valid = false;
    std::vector<int> status;
    // optional per-kernel timing
    bool timing = false;
    std::vector<cudaEvent_t> ev;
    std::vector<int> ev_phase;
    size_t ev_used = 0;
    float phase_ms[PH_COUNT] = {0};
    int phase_launches[PH_COUNT] = {0};
    cudaEvent_t ev_run0 = nullptr, ev_run1 = nullptr;  // device time of the last run
    float last_run_ms = 0;
    // pinned staging for the host-buffer entry
    unsigned char *h_stage = nullptr;
    long long h_stage_bytes = 0;

    ~hmp3_batch() {
        cudaSetDevice(device);
        cudaFree(d_tabs);
        cudaFree(d_st);
        cudaFree(d_so);
        cudaFree(d_pcm);
        cudaFree(d_sw);
        cudaFree(d_sw_init);
        cudaFree(d_rs);
        cudaFree(d_rs_cold);
        cudaFree(d_main);
        cudaFree(d_frames);
        cudaFree(d_res);
        cudaFree(d_out_off);
        cudaFree(d_out);
        for (int k = 0; k < kMaxSets; k++) {
            cudaFree(cb2[k].P);
            cudaFree(cb2[k].E);
            cudaFree(cb2[k].gi);
            cudaFree(cb2[k].xr);
            cudaFree(cb2[k].raw);
            cudaFree(cb2[k].ms_raw);
            cudaFree(cb2[k].ms);
            cudaFree(cb2[k].sm);
            cudaFree(cb2[k].prep);
            cudaFree(cb2[k].pack);
            cudaFree(cb2[k].fr0);
            cudaFree(cb2[k].fr1);
            cudaFree(cb2[k].fd1);
            if (ev_a[k]) cudaEventDestroy(ev_a[k]);
            if (ev_r[k]) cudaEventDestroy(ev_r[k]);
            if (ev_p[k]) cudaEventDestroy(ev_p[k]);
        }
        if (ev_start) cudaEventDestroy(ev_start);
        cudaFree(d_flags);
        cudaFree(d_out_inc);
        cudaFree(d_done_lo);
        cudaFree(d_bytes_done);
        if (h_prog) cudaFreeHost(h_prog);
        for (auto e : ev_o) cudaEventDestroy(e);
        if (stream_o) cudaStreamDestroy(stream_o);
        cudaFree(d_cycles);
        cudaFree(d_msmem);
        cudaFree(d_pcmf);
        cudaFree(d_dc);
        cudaFree(d_polyw);
        cudaFree(d_psy);
        if (stream_c) cudaStreamDestroy(stream_c);
        for (int k = 0; k < kMaxSets; k++)
            if (ev_c[k]) cudaEventDestroy(ev_c[k]);
        if (stream_a0) cudaStreamDestroy(stream_a0);
        if (stream_p0) cudaStreamDestroy(stream_p0);
        for (auto e : ev) cudaEventDestroy(e);
        if (ev_run0) cudaEventDestroy(ev_run0);
        if (ev_run1) cudaEventDestroy(ev_run1);
        if (h_stage) cudaFreeHost(h_stage);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

// number of encode calls the CLI makes for a clip (tomp3.cpp:923-931: four zero frames appended at EOF,
// whole frames only)
// Encode calls of the CLI's main loop for a stream of `nsamples` (test/tomp3.cpp:908-942): at end of file it appends
// 4 x bytes_in_init zero bytes and keeps calling while bytes_in_init bytes are buffered; bytes_in_init is what
// Csrc::sr_convert_init returns, 1152 * source/target + ntaps (= 1 without down-sampling) sample frames
// (srcc.cpp:185-187, 769-773) = 1153, and every call consumes 1152: floor((n + 3 * 1153) / 1152) + 1 calls.
long long calls_for(long long nsamples) { return (nsamples + 3 * 1153 + 1152) / 1152; }
const int kFlushCalls = 12;  // upper bound of the tail-flush calls we provision analysis data for

int max_main_frame_bytes(const EncConfig &C) {
    return C.vbr_flag ? C.vbr_main_framebytes[C.ivbr_max] : C.main_framebytes + 1;
}

// Number of chunk buffer sets of a batch plan.  Two: Phase A of chunk c+1 overlaps the serial stage of c.  Three
// (HMP3_CHUNK_SETS=3) lets Phase A run a chunk further ahead; measured no faster (Phase A only gets SM slots in
// the tail of a serial-stage launch either way), so it is not the default.
int chunk_sets() {
    const char *e = getenv("HMP3_CHUNK_SETS");
    const int s = e ? atoi(e) : 2;
    return s < 2 ? 2 : (s > kMaxSets ? kMaxSets : s);
}

int plan_create(hmp3_batch *b, const hmp3_control *controls, const long long *nsamples, int n, int device,
                int chunk_granules, bool analysis_only, bool streaming = false, const int *formats = nullptr) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) {
        set_err("no usable CUDA device (this library has no CPU path)");
        return HMP3_ERR_NO_DEVICE;
    }
    CK(cudaSetDevice(device));
    for (int i = 0; i < n; i++)
        if (nsamples[i] < 0) {
            set_err("negative sample count");
            return HMP3_ERR_ARG;
        }
    b->device = device;
    b->n = n;
    b->status.assign(n, HMP3_OK);
    b->fmt.assign(n, 0);
    if (formats)
        for (int i = 0; i < n; i++) b->fmt[i] = formats[i] ? 1 : 0;
    bool any_float = false;
    b->st_h.resize(n);
    b->so_h.resize(n);
    long long pcm_off = 0, pcmf_off = 0, main_off = 0, frames_off = 0, out_cap = 0;
    int max_gran = 0, max_frames = 0;
    for (int i = 0; i < n; i++) {
        int cfg = -1;
        for (size_t c = 0; c < b->controls.size(); c++)
            if (same_control(b->controls[c], controls[i])) cfg = (int)c;
        if (cfg < 0) {
            EncTables *T = new EncTables;
            int unsup = 0;
            int r = build_tables(&controls[i], T, &unsup);
            if (!r) {
                b->status[i] = unsup ? HMP3_ERR_UNSUPPORTED : HMP3_ERR_BAD_CONTROL;
            } else {
                cfg = (int)b->controls.size();
                b->controls.push_back(controls[i]);
                b->tabs_h.push_back(*T);
            }
            delete T;
        }
        StreamDev &sd = b->st_h[i];
        memset(&sd, 0, sizeof(sd));
        sd.cfg = cfg < 0 ? 0 : cfg;
        sd.nch = cfg < 0 ? 0 : b->tabs_h[cfg].cfg.nchan;
        sd.nsamples = nsamples[i];
        sd.pcm_off = pcm_off;
        long long calls = calls_for(nsamples[i]);
        sd.ngran_real = cfg < 0 ? 0 : (int)(2 * calls);
        sd.ngran = cfg < 0 ? 0 : (int)(2 * (calls + kFlushCalls));
        if (streaming) sd.ngran_real = sd.ngran;  // a handle never decides by itself that the stream has ended
        pcm_off += nsamples[i] * (sd.nch ? sd.nch : 1);
        pcm_off = (pcm_off + 7) & ~7LL;
        sd.pcmf_off = -1;
        sd.pcmf_len = 0;
        sd.rawf_off = -1;
        if (cfg >= 0 && b->fmt[i]) {
            sd.rawf_off = pcmf_off;
            pcmf_off += ((nsamples[i] * sd.nch + 3) & ~3LL);
            any_float = true;
        }
        if (cfg >= 0 && b->tabs_h[cfg].cfg.filter_select) {
            sd.pcmf_off = pcmf_off;
            sd.pcmf_len = 576LL * sd.ngran;
            pcmf_off += sd.pcmf_len * sd.nch;
            b->any_filter = true;
        }
        max_gran = std::max(max_gran, sd.ngran);
        StreamOut &so = b->so_h[i];
        so.main_off = main_off;
        so.frames_off = frames_off;
        so.frames_cap = 0;
        if (cfg >= 0) {
            const EncConfig &C = b->tabs_h[cfg].cfg;
            const int nfr = sd.ngran / C.granules_per_frame + 2;
            so.frames_cap = nfr;
            main_off += ((long long)nfr * max_main_frame_bytes(C) + 4096 + 15) & ~15LL;
            frames_off += nfr;
            sd.out_off = out_cap;
            sd.out_cap = (long long)nfr * (4 + C.side_bytes + max_main_frame_bytes(C));
            out_cap += (sd.out_cap + 15) & ~15LL;
            max_frames = std::max(max_frames, nfr);
        }
    }
    if (b->tabs_h.empty()) {
        set_err("no stream has a valid control block");
        return HMP3_ERR_BAD_CONTROL;
    }
    for (int i = 0; i < n; i++)
        if (b->status[i] == HMP3_OK) {
            if (b->tabs_h[b->st_h[i].cfg].cfg.allocator == 1) b->any_allo1 = true;
            else b->any_allo0 = true;
        }
    b->pcm_elems = pcm_off;
    b->max_gran = max_gran;
    b->max_frames = max_frames;
    b->NG = chunk_granules;
    b->out_cap = out_cap;
    CK(cudaStreamCreate(&b->stream));
    CK(cudaMalloc(&b->d_tabs, sizeof(EncTables) * b->tabs_h.size()));
    CK(cudaMemcpy(b->d_tabs, b->tabs_h.data(), sizeof(EncTables) * b->tabs_h.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&b->d_st, sizeof(StreamDev) * n));
    CK(cudaMemcpy(b->d_st, b->st_h.data(), sizeof(StreamDev) * n, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&b->d_pcm, sizeof(int16_t) * std::max<long long>(b->pcm_elems, 8)));
    CK(cudaMemset(b->d_pcm, 0, sizeof(int16_t) * std::max<long long>(b->pcm_elems, 8)));
    if (b->any_filter || any_float) {
        CK(cudaMalloc(&b->d_pcmf, sizeof(float) * std::max<long long>(pcmf_off, 1)));
        CK(cudaMemset(b->d_pcmf, 0, sizeof(float) * std::max<long long>(pcmf_off, 1)));
        CK(cudaMalloc(&b->d_dc, sizeof(float) * 2 * n));
    }
    if (const char *pm = getenv("HMP3_POLY_MODE")) {  // experiments: the filterbank as a tensor-core contraction
        b->poly_mode = !strcmp(pm, "3xtf32") ? 3 : (!strcmp(pm, "tf32") ? 1 : 0);
        if (b->any_filter || any_float) b->poly_mode = 0;  // that kernel reads int16 PCM only
        if (b->poly_mode) {
            std::vector<float> w(polymm_matrix_floats());
            build_polymm_matrix(&b->tabs_h[0], w.data());
            CK(cudaMalloc(&b->d_polyw, sizeof(float) * w.size()));
            CK(cudaMemcpy(b->d_polyw, w.data(), sizeof(float) * w.size(), cudaMemcpyHostToDevice));
        }
    }
    CK(cudaMalloc(&b->d_msmem, sizeof(int) * n));
    CK(cudaMalloc(&b->d_psy, sizeof_psy_state() * n * 2));
    CK(cudaMalloc(&b->d_sw, sizeof(SwitchState) * n));
    CK(cudaMalloc(&b->d_sw_init, sizeof(SwitchState) * n));
    {
        std::vector<SwitchState> sw(n);
        for (auto &s : sw) switch_state_init(&s);
        CK(cudaMemcpy(b->d_sw_init, sw.data(), sizeof(SwitchState) * n, cudaMemcpyHostToDevice));
    }
    const long long NG = b->NG, G = NG + 3;
    b->nbuf = analysis_only ? 1 : (streaming ? 1 : chunk_sets());
    if (getenv("HMP3_SERIALIZE")) {  // diagnostics: every kernel on one stream, so per-kernel times are uncontended
        b->stream_a = b->stream_p = b->stream;
        b->one_stream = true;
    } else {  // Phase A and the packing pass get the free SM slots before the serial stage's next launch does
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const bool prio = !(getenv("HMP3_NO_STREAM_PRIO"));
        CK(cudaStreamCreateWithPriority(&b->stream_a0, cudaStreamDefault, prio ? hi : lo));
        CK(cudaStreamCreateWithPriority(&b->stream_p0, cudaStreamDefault, prio ? hi : lo));
        b->stream_a = b->stream_a0;
        b->stream_p = b->stream_p0;
    }
    CK(cudaEventCreateWithFlags(&b->ev_start, cudaEventDisableTiming));
    for (int k = 0; k < b->nbuf; k++) {
        ChunkBufs &cb = b->cb2[k];
        cb.NG = (int)NG;
        CK(cudaMalloc(&cb.P, sizeof(float) * n * G * 2 * 576));
        CK(cudaMalloc(&cb.E, sizeof(int) * n * G * 2 * 9));
        CK(cudaMalloc(&cb.gi, sizeof(GranuleInfo) * n * NG));
        CK(cudaMalloc(&cb.xr, sizeof(float) * n * NG * 2 * 576));
        CK(cudaMalloc(&cb.raw, sizeof(PsyRaw) * n * NG * 2));
        CK(cudaMalloc(&cb.ms_raw, sizeof(int) * n * NG));
        CK(cudaMalloc(&cb.ms, n * NG));
        CK(cudaMalloc(&cb.sm, sizeof(float) * 2 * 72 * n * NG));
        CK(cudaMalloc(&cb.prep, sizeof_prep_granule() * n * NG));
        if (!analysis_only) {
            CK(cudaMalloc(&cb.pack, sizeof_pack_gc() * n * NG * 2));
            CK(cudaMalloc(&cb.fr0, sizeof(int) * n));
            CK(cudaMalloc(&cb.fr1, sizeof(int) * n));
            CK(cudaMalloc(&cb.fd1, sizeof(int) * n));
        }
        CK(cudaEventCreateWithFlags(&b->ev_a[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&b->ev_r[k], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&b->ev_p[k], cudaEventDisableTiming));
    }
    if (!analysis_only) {
        CK(cudaMalloc(&b->d_so, sizeof(StreamOut) * n));
        CK(cudaMemcpy(b->d_so, b->so_h.data(), sizeof(StreamOut) * n, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&b->d_rs, sizeof_rate_state() * n));
        CK(cudaMalloc(&b->d_rs_cold, sizeof_rate_cold() * n));
        keep_rate_state_in_l2(b->stream, b->d_rs, sizeof_rate_state() * (size_t)n, device);
        cudaDeviceGetAttribute(&b->sm_count, cudaDevAttrMultiProcessorCount, device);
        CK(cudaMalloc(&b->d_main, std::max<long long>(main_off, 16)));
        CK(cudaMalloc(&b->d_frames, sizeof_frame_rec() * std::max<long long>(frames_off, 1)));
        CK(cudaMalloc(&b->d_res, sizeof(StreamResult) * n));
        CK(cudaMalloc(&b->d_flags, sizeof(int) * n));
        b->flags_h.assign(n, 0);
        CK(cudaMalloc(&b->d_out_off, sizeof(long long) * (n + 1)));
        CK(cudaMalloc(&b->d_out, std::max<long long>(out_cap, 16)));
        b->res_h.resize(n);
        b->out_off_h.resize(n + 1);
    }
    return HMP3_OK;
}

// device address / element size of stream i's input PCM
char *pcm_dev_ptr(hmp3_batch *b, int i, long long sample, size_t *elem) {
    const StreamDev &sd = b->st_h[i];
    if (b->fmt[i]) {
        *elem = sizeof(float);
        return (char *)(b->d_pcmf + sd.rawf_off + sample * sd.nch);
    }
    *elem = sizeof(int16_t);
    return (char *)(b->d_pcm + sd.pcm_off + sample * sd.nch);
}
int plan_reset_state(hmp3_batch *b) {
    launch_prepare_init(b->d_msmem, b->d_psy, b->n, b->stream);
    if (b->any_filter) CK(cudaMemsetAsync(b->d_dc, 0, sizeof(float) * 2 * b->n, b->stream));
    CK(cudaMemcpyAsync(b->d_sw, b->d_sw_init, sizeof(SwitchState) * b->n, cudaMemcpyDeviceToDevice, b->stream));
    return HMP3_OK;
}

// Per-kernel timing (when enabled): events are recorded in (begin, end) pairs on the launching stream;
// mark(phase >= 0) opens a pair, mark(-1) closes it.
void mark(hmp3_batch *b, int phase, cudaStream_t st) {
    if (!b->timing) return;
    if (b->ev_used == b->ev.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        b->ev.push_back(e);
        b->ev_phase.push_back(-1);
    }
    b->ev_phase[b->ev_used] = phase;
    cudaEventRecord(b->ev[b->ev_used++], st);
}

// Phase A for the chunk starting at encode granule K0, into chunk buffer set `k`, on stream `st`.
int launch_analysis(hmp3_batch *b, int K0, int k, cudaStream_t st, bool with_prepare = true, int ng = 0) {
    const int n = b->n;
    ChunkBufs cb = b->cb2[k];
    if (ng > 0) cb.NG = ng;
    if (b->any_filter)
        launch_dc_filter(b->d_tabs, b->d_st, b->d_pcm, b->d_pcmf, b->d_dc, 576LL * K0, 576LL * (K0 + cb.NG), n, st);
    mark(b, PH_POLY, st);
    if (b->poly_mode) launch_polyphase_mm(b->d_tabs, b->d_st, b->d_pcm, b->d_polyw, cb, K0, n, b->poly_mode, st);
    else launch_polyphase(b->d_tabs, b->d_st, b->d_pcm, b->d_pcmf, cb, K0, n, st);
    mark(b, -1, st);
    mark(b, PH_ATTACK, st);
    launch_attack(b->d_tabs, b->d_st, cb, K0, n, st);
    mark(b, -1, st);
    mark(b, PH_SWITCH, st);
    launch_switch_scan(b->d_tabs, b->d_st, b->d_sw, cb, K0, n, st);
    mark(b, -1, st);
    mark(b, PH_HYBRID, st);
    launch_hybrid(b->d_tabs, b->d_st, cb, K0, n, st);
    mark(b, -1, st);
    mark(b, PH_PSY, st);
    launch_psy_stage1(b->d_tabs, b->d_st, cb, K0, n, st);
    mark(b, -1, st);
    if (with_prepare) {
        mark(b, PH_MS, st);
        launch_ms_scan(b->d_tabs, b->d_st, b->d_msmem, cb, K0, n, st);
        mark(b, -1, st);
        mark(b, PH_PSY2, st);
        launch_psy_stage2(b->d_tabs, b->d_st, b->d_psy, cb, K0, n, st);
        mark(b, -1, st);
        mark(b, PH_PREP, st);
        launch_prepare(b->d_tabs, b->d_st, cb, K0, n, st);
        mark(b, -1, st);
        b->launches += 3;
    }
    b->launches += 5;
    CK(cudaGetLastError());
    return HMP3_OK;
}

int run_plan_impl(hmp3_batch *b);
// Errors after work has been queued: nothing may still be reading the caller's PCM or writing the caller's output
// buffers when the call returns, so the device is drained first.
int drain_on_error(hmp3_batch *b, int r) {
    if (r != HMP3_OK && b) {
        const std::string keep = g_err;
        cudaSetDevice(b->device);
        cudaDeviceSynchronize();
        cudaGetLastError();
        g_err = keep;
    }
    return r;
}
int run_plan(hmp3_batch *b) { return drain_on_error(b, run_plan_impl(b)); }
int run_plan_impl(hmp3_batch *b) {
    const int n = b->n;
    CK(cudaSetDevice(b->device));
    b->launches = 0;
    b->ev_used = 0;
    b->results_valid = false;
    if (!b->ev_run0) {
        CK(cudaEventCreate(&b->ev_run0));
        CK(cudaEventCreate(&b->ev_run1));
    }
    CK(cudaEventRecord(b->ev_run0, b->stream));
    int r = plan_reset_state(b);
    if (r != HMP3_OK) return r;
    launch_rate_init(b->d_tabs, b->d_st, b->d_rs, b->d_rs_cold, n, b->stream);
    b->launches++;
    // Phase A runs on its own stream one chunk ahead of the serial stage, the packing pass on a third stream
    // one chunk behind it (two chunk buffer sets):
    //   analysis(c) -> ev_a -> serial(c) -> ev_r -> pack(c) -> ev_p ;  with S buffer sets analysis(c+S) waits ev_r(c) and
    //   serial(c+S) waits ev_p(c): S = 3 lets Phase A finish a whole chunk ahead, so the serial stage never waits for it
    CK(cudaMemsetAsync(b->d_flags, 0, sizeof(int) * n, b->stream));
    if (b->direct) CK(cudaMemsetAsync(b->d_done_lo, 0, sizeof(int) * n, b->stream));
    CK(cudaEventRecord(b->ev_start, b->stream));
    CK(cudaStreamWaitEvent(b->stream_a, b->ev_start, 0));
    CK(cudaStreamWaitEvent(b->stream_p, b->ev_start, 0));
    // the first chunk is short so that the serial stage starts early (its Phase A cannot overlap anything); a short
    // chunk is just a narrower view of the same buffers (every kernel indexes with the view's NG)
    int c = 0;
    const bool trace = getenv("HMP3_TRACE_HOST") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto ms_now = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    for (int K0 = 0; K0 < b->max_gran; c++) {
        const double t_c0 = trace ? ms_now() : 0.0;
        const int nb = b->nbuf;
        const int k = c % nb;
        const int ng_c = (c == 0 && b->NG > 32) ? 32 : b->NG;
        ChunkBufs view = b->cb2[k];
        view.NG = ng_c;
        const int K0_this = K0;
        K0 += ng_c;
        if (c >= nb) CK(cudaStreamWaitEvent(b->stream_a, b->ev_r[k], 0));
        if (b->staged) {
            // the samples this chunk's polyphase needs first (up to the end of granule K0+NG-1), one DMA copy per
            // stream on the copy stream: the copy engine needs no SM resources, so the transfer overlaps the
            // serial stage of the previous chunk whatever its occupancy
            if (!b->stream_c) {
                CK(cudaStreamCreate(&b->stream_c));
                for (int q = 0; q < kMaxSets; q++) CK(cudaEventCreateWithFlags(&b->ev_c[q], cudaEventDisableTiming));
            }
            if (c == 0) CK(cudaStreamWaitEvent(b->stream_c, b->ev_start, 0));
            const long long lo = c == 0 ? 0 : 576LL * K0_this, hi = 576LL * (K0_this + ng_c);
            // one batched submission for the whole chunk (cudaMemcpyBatchAsync, CUDA 12.8+): thousands of separate
            // cudaMemcpyAsync calls cost ~15 us of host time each, which delays the first chunks
            b->cp_dst.clear();
            b->cp_src.clear();
            b->cp_len.clear();
            for (int i = 0; i < n; i++) {
                const StreamDev &sd = b->st_h[i];
                if (b->status[i] != HMP3_OK) continue;
                const long long a = std::min<long long>(lo, sd.nsamples), e = std::min<long long>(hi, sd.nsamples);
                if (e <= a) continue;
                size_t el;
                char *dst = pcm_dev_ptr(b, i, a, &el);
                b->cp_dst.push_back(dst);
                b->cp_src.push_back((void *)((const char *)b->h_src[i] + el * a * sd.nch));
                b->cp_len.push_back(el * (e - a) * sd.nch);
            }
            bool batched = false;
            if (!b->cp_len.empty() && !b->no_batch_copy) {
                cudaMemcpyAttributes at;
                memset(&at, 0, sizeof(at));
                at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
                at.flags = cudaMemcpyFlagPreferOverlapWithCompute;
                size_t idx0 = 0, fail = 0;
                if (cudaMemcpyBatchAsync(b->cp_dst.data(), b->cp_src.data(), b->cp_len.data(), b->cp_len.size(), &at, &idx0,
                                         1, &fail, b->stream_c) == cudaSuccess)
                    batched = true;
                else {
                    cudaGetLastError();
                    b->no_batch_copy = true;  // older driver: fall back for good
                }
            }
            if (!batched)
                for (size_t q = 0; q < b->cp_len.size(); q++)
                    CK(cudaMemcpyAsync(b->cp_dst[q], b->cp_src[q], b->cp_len[q], cudaMemcpyHostToDevice, b->stream_c));
            CK(cudaEventRecord(b->ev_c[k], b->stream_c));
            CK(cudaStreamWaitEvent(b->stream_a, b->ev_c[k], 0));
            if (trace) fprintf(stderr, "[run_plan] chunk %2d (%3d granules): copies enqueued %8.2f -> %8.2f ms\n", c, ng_c, t_c0, ms_now());
        }
        r = launch_analysis(b, K0_this, k, b->stream_a, true, ng_c);
        if (r != HMP3_OK) return r;
        CK(cudaEventRecord(b->ev_a[k], b->stream_a));
        CK(cudaStreamWaitEvent(b->stream, b->ev_a[k], 0));
        if (c >= nb) CK(cudaStreamWaitEvent(b->stream, b->ev_p[k], 0));
        mark(b, PH_RATE, b->stream);
        if (b->any_allo0) {
            if (rate_mode_phased(n, b->sm_count) && !b->d_cycles)
                launch_rate_ph(b->d_tabs, b->d_st, b->d_so, b->d_rs, view, b->d_frames, K0_this, n, b->stream);
            else
                launch_rate(b->d_tabs, b->d_st, b->d_so, b->d_rs, view, b->d_main, b->d_frames, K0_this, n, b->stream,
                            (b->d_cycles && c < kCycleLaunches) ? b->d_cycles + (long long)c * n : nullptr);
        }
        if (b->any_allo1) {
            launch_rate_a1(b->d_tabs, b->d_st, b->d_so, b->d_rs, view, b->d_main, b->d_frames, K0_this, n, b->stream);
            b->launches++;
        }
        if (b->d_cycles && c < kCycleLaunches) b->cycle_launches = c + 1;
        mark(b, -1, b->stream);
        if (b->tap_stream >= 0 && b->tap_out && 2LL * K0_this < b->tap_cap) {  // diagnostics: the chunk's records of one stream
            const size_t rec = sizeof_pack_gc();
            const long long nrec = std::min<long long>(2LL * ng_c, b->tap_cap - 2LL * K0_this);
            CK(cudaMemcpyAsync(b->tap_out + rec * 2 * K0_this,
                               (const unsigned char *)view.pack + rec * 2 * ((long long)b->tap_stream * ng_c), rec * nrec,
                               cudaMemcpyDeviceToHost, b->stream));
        }
        CK(cudaEventRecord(b->ev_r[k], b->stream));
        CK(cudaStreamWaitEvent(b->stream_p, b->ev_r[k], 0));
        mark(b, PH_PACK, b->stream_p);
        launch_pack(b->d_tabs, b->d_st, b->d_so, view, b->d_main, b->d_frames, b->d_flags, K0_this, n, b->stream_p);
        mark(b, -1, b->stream_p);
        if (b->direct && c < b->h_prog_chunks) {
            launch_assemble_inc(b->d_tabs, b->d_st, b->d_so, view, b->d_done_lo, b->d_main, b->d_frames, b->d_out_inc,
                                b->d_bytes_done, n, b->stream_p);
            CK(cudaMemcpyAsync(b->h_prog + (size_t)c * n, b->d_bytes_done, sizeof(long long) * n, cudaMemcpyDeviceToHost,
                               b->stream_p));
            CK(cudaEventRecord(b->ev_o[c], b->stream_p));
            b->launches += 2;
        }
        CK(cudaEventRecord(b->ev_p[k], b->stream_p));
        b->launches += 2;
    }
    b->chunks_run = c;
    for (int k = 0; k < b->nbuf && k < c; k++) CK(cudaStreamWaitEvent(b->stream, b->ev_p[k], 0));
    cudaEvent_t ev_asm = nullptr;
    if (b->timing) {
        mark(b, PH_ASSEMBLE, b->stream);  // re-recorded by launch_finish right before the assembly kernel
        ev_asm = b->ev[b->ev_used - 1];
    }
    launch_finish(b->d_tabs, b->d_st, b->d_so, b->d_rs, b->d_frames, b->d_res, b->d_out_off, b->d_main, b->d_out,
                  b->max_frames, n, b->stream, ev_asm);
    mark(b, -1, b->stream);
    b->launches += 3;
    CK(cudaGetLastError());
    CK(cudaEventRecord(b->ev_run1, b->stream));
    CK(cudaMemcpyAsync(b->res_h.data(), b->d_res, sizeof(StreamResult) * n, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaMemcpyAsync(b->out_off_h.data(), b->d_out_off, sizeof(long long) * (n + 1), cudaMemcpyDeviceToHost,
                       b->stream));
    CK(cudaMemcpyAsync(b->flags_h.data(), b->d_flags, sizeof(int) * n, cudaMemcpyDeviceToHost, b->stream));
    return HMP3_OK;
}

int sync_plan(hmp3_batch *b) {
    CK(cudaSetDevice(b->device));
    CK(cudaStreamSynchronize(b->stream));
    b->results_valid = true;
    cudaEventElapsedTime(&b->last_run_ms, b->ev_run0, b->ev_run1);
    if (b->timing) {
        for (int p = 0; p < PH_COUNT; p++) {
            b->phase_ms[p] = 0;
            b->phase_launches[p] = 0;
        }
        for (size_t i = 0; i + 1 < b->ev_used; i += 2) {  // (begin, end) pairs
            const int p = b->ev_phase[i];
            if (p < 0) continue;
            float ms = 0;
            cudaEventElapsedTime(&ms, b->ev[i], b->ev[i + 1]);
            b->phase_ms[p] += ms;
            b->phase_launches[p]++;
        }
        b->timeline.clear();  // (phase, begin, end) in ms since the run began
        for (size_t i = 0; i + 1 < b->ev_used; i += 2) {
            if (b->ev_phase[i] < 0) continue;
            float t0 = 0, t1 = 0;
            cudaEventElapsedTime(&t0, b->ev_run0, b->ev[i]);
            cudaEventElapsedTime(&t1, b->ev_run0, b->ev[i + 1]);
            b->timeline.push_back({(float)b->ev_phase[i], t0, t1});
        }
    }
    return HMP3_OK;
}

}  // namespace

extern "C" {

const char *hmp3_get_last_error(void) { return g_err.c_str(); }

int hmp3_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

void hmp3_control_defaults(hmp3_control *ec) { control_defaults(ec); }

int hmp3_resolve_control(const hmp3_control *ec, hmp3_resolved *out) {
    EncTables *T = new EncTables;
    int unsup = 0;
    const int r = build_tables(ec, T, &unsup);
    if (!r) {
        delete T;
        set_err(unsup ? "configuration outside the built path" : "control block rejected");
        return unsup ? HMP3_ERR_UNSUPPORTED : HMP3_ERR_BAD_CONTROL;
    }
    const EncConfig &C = T->cfg;
    memset(out, 0, sizeof(*out));
    out->nchan = C.nchan;
    out->h_id = C.h_id;
    out->sr_index = C.sr_index;
    out->nband = C.nband;
    out->band_limit = C.band_limit;
    out->nsb = C.nsb;
    out->nsb_limit = C.nsb_limit;
    out->nsb_limit_ms0 = C.nsb_hybrid;
    out->nsb_limit_ms1 = C.nsb_limit_ms1;
    out->ave_target_bits = C.ave_target_bits;
    out->framebytes = C.framebytes;
    out->main_framebytes = C.main_framebytes;
    out->side_bytes = C.side_bytes;
    out->remainder = C.pad_remainder;
    out->divisor = C.pad_divisor;
    out->ms_flag = C.ms_flag;
    out->is_flag = C.is_flag;
    out->frame_driver = C.frame_driver;
    out->granule_driver = 0;
    out->ivbr_min = C.ivbr_min;
    out->ivbr_max = C.ivbr_max;
    out->vbr_pool_target = C.vbr_pool_target;
    out->short_block_threshold = C.short_block_threshold;
    out->h_mode = C.h_mode;
    out->br_index = C.br_index;
    out->totbitrate = C.totbitrate;
    out->samprate = C.samprate;
    out->band_limit_stereo = C.band_limit_stereo;
    out->sf_bit_max = C.sf_bit_max;
    out->nsf_stereo = C.nsf_stereo;
    for (int i = 0; i < 4; i++) out->head[i] = C.head[i];
    out->hf_flag = C.hf_flag;
    out->filter_select = C.filter_select;
    out->bytes_in = r;
    delete T;
    return HMP3_OK;
}

int64_t hmp3_batch_out_bound(const hmp3_control *control, int64_t num_samples) {
    EncTables *T = new EncTables;
    if (!build_tables(control, T, nullptr)) {
        delete T;
        return 0;
    }
    const EncConfig &C = T->cfg;
    const long long nfr = 2 * (calls_for(num_samples) + kFlushCalls) / C.granules_per_frame + 2;
    const int64_t r = nfr * (4 + C.side_bytes + max_main_frame_bytes(C));
    delete T;
    return r;
}

hmp3_batch *hmp3_batch_create(const hmp3_control *controls, const int64_t *num_samples, int n, int device) {
    return hmp3_batch_create_ex(controls, num_samples, nullptr, n, device);
}

hmp3_batch *hmp3_batch_create_ex(const hmp3_control *controls, const int64_t *num_samples, const int32_t *pcm_formats,
                                 int n, int device) {
    if (n <= 0 || !controls || !num_samples) {
        set_err("bad arguments");
        return nullptr;
    }
    hmp3_batch *b = new hmp3_batch;
    std::vector<long long> ns(num_samples, num_samples + n);
    // chunk length: the longer the chunk, the less the serial-stage kernel idles at its tail (it ends when the
    // slowest stream of the chunk does); bounded by the memory the two chunk buffer sets may take
    int ng = 256;
    if (getenv("HMP3_CHUNK_GRANULES")) ng = atoi(getenv("HMP3_CHUNK_GRANULES"));
    else {
        size_t free_b = 0, total_b = 0;
        if (cudaSetDevice(device) == cudaSuccess && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const double per_granule = chunk_sets() * 19500.0 * n;  // bytes of all buffer sets per granule of every stream
            while (ng > 32 && per_granule * ng > 0.45 * (double)free_b) ng >>= 1;
        }
    }
    if (ng < 2) ng = 2;
    ng &= ~1;
    int r = plan_create(b, controls, ns.data(), n, device, ng, false, false, pcm_formats);
    g_create_status = r;
    if (r != HMP3_OK) {
        delete b;
        return nullptr;
    }
    return b;
}

void hmp3_batch_destroy(hmp3_batch *b) { delete b; }

int16_t *hmp3_batch_device_pcm(hmp3_batch *b) { return b ? b->d_pcm : nullptr; }
int64_t hmp3_batch_pcm_offset(const hmp3_batch *b, int i) { return (b && i >= 0 && i < b->n) ? b->st_h[i].pcm_off : -1; }

int hmp3_batch_set_tail(hmp3_batch *b, int i, float value) {
    if (!b || i < 0 || i >= b->n || b->fmt[i] != HMP3_PCM_F32) {
        set_err("hmp3_batch_set_tail: bad stream index or not a float stream");
        return HMP3_ERR_ARG;
    }
    CK(cudaSetDevice(b->device));
    b->st_h[i].tail = value;
    CK(cudaMemcpy(b->d_st + i, &b->st_h[i], sizeof(StreamDev), cudaMemcpyHostToDevice));
    return HMP3_OK;
}
uint8_t *hmp3_batch_device_out(hmp3_batch *b) { return b ? b->d_out : nullptr; }
int64_t hmp3_batch_out_capacity(const hmp3_batch *b) { return b ? b->out_cap : 0; }

namespace {
int upload_any(hmp3_batch *b, int i, const void *pcm, int64_t num_samples, int want_fmt) {
    if (!b || !pcm || i < 0 || i >= b->n || num_samples != b->st_h[i].nsamples || b->fmt[i] != want_fmt) {
        set_err("upload: stream index, length or sample format does not match the plan");
        return HMP3_ERR_ARG;
    }
    CK(cudaSetDevice(b->device));
    size_t el;
    char *dst = pcm_dev_ptr(b, i, 0, &el);
    CK(cudaMemcpyAsync(dst, pcm, el * num_samples * b->st_h[i].nch, cudaMemcpyHostToDevice, b->stream));
    return HMP3_OK;
}
}  // namespace

int hmp3_batch_upload(hmp3_batch *b, int i, const int16_t *pcm, int64_t num_samples) {
    return upload_any(b, i, pcm, num_samples, 0);
}
int hmp3_batch_upload_f32(hmp3_batch *b, int i, const float *pcm, int64_t num_samples) {
    return upload_any(b, i, pcm, num_samples, 1);
}

int hmp3_batch_wait_uploads(hmp3_batch *b) {
    if (!b) return HMP3_ERR_ARG;
    CK(cudaSetDevice(b->device));
    CK(cudaStreamSynchronize(b->stream));
    return HMP3_OK;
}

int hmp3_batch_set_timing(hmp3_batch *b, int on) {
    if (!b) return HMP3_ERR_ARG;
    b->timing = on != 0;
    return HMP3_OK;
}

int hmp3_batch_set_serialize(hmp3_batch *b, int on) {
    if (!b) return HMP3_ERR_ARG;
    if (!b->stream_a0) return HMP3_OK;  // created with HMP3_SERIALIZE: always on
    b->one_stream = on != 0;
    b->stream_a = on ? b->stream : b->stream_a0;
    b->stream_p = on ? b->stream : b->stream_p0;
    return HMP3_OK;
}

int hmp3_batch_run(hmp3_batch *b, int async) {
    if (!b) return HMP3_ERR_ARG;
    int r = run_plan(b);
    if (r != HMP3_OK) return r;
    return async ? HMP3_OK : sync_plan(b);
}

int hmp3_batch_sync(hmp3_batch *b) { return b ? sync_plan(b) : HMP3_ERR_ARG; }

int hmp3_batch_results(hmp3_batch *b, int64_t *out_bytes, int32_t *out_frames, int64_t *out_offsets,
                       int32_t *status) {
    if (!b || !b->results_valid) {
        set_err("results: no completed run");
        return HMP3_ERR_ARG;
    }
    for (int i = 0; i < b->n; i++) {
        int st = b->status[i];
        if (st == HMP3_OK && !b->res_h[i].finished) st = HMP3_ERR_OUT_SPACE;
        if (st == HMP3_OK && b->flags_h[i]) st = HMP3_ERR_INTERNAL;
        if (out_bytes) out_bytes[i] = b->status[i] == HMP3_OK ? b->res_h[i].out_bytes : 0;
        if (out_frames) out_frames[i] = b->status[i] == HMP3_OK ? b->res_h[i].frames : 0;
        if (out_offsets) out_offsets[i] = b->out_off_h[i];
        if (status) status[i] = st;
    }
    return HMP3_OK;
}

int hmp3_batch_download(hmp3_batch *b, int i, uint8_t *out, int64_t cap) {
    if (!b || !out || !b->results_valid || i < 0 || i >= b->n) {
        set_err("download: no completed run or bad index");
        return HMP3_ERR_ARG;
    }
    const long long nb = b->res_h[i].out_bytes;
    if (nb > cap) {
        set_err("download: output buffer too small");
        return HMP3_ERR_OUT_SPACE;
    }
    CK(cudaSetDevice(b->device));
    CK(cudaMemcpy(out, b->d_out + b->out_off_h[i], nb, cudaMemcpyDeviceToHost));
    return HMP3_OK;
}

int hmp3_batch_download_all(hmp3_batch *b, uint8_t *out, int64_t cap, int64_t *total) {
    if (!b->results_valid) {
        set_err("download: no completed run");
        return HMP3_ERR_ARG;
    }
    const long long nb = b->out_off_h[b->n];
    if (total) *total = nb;
    if (nb > cap) {
        set_err("download: output buffer too small");
        return HMP3_ERR_OUT_SPACE;
    }
    CK(cudaSetDevice(b->device));
    CK(cudaMemcpyAsync(out, b->d_out, nb, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return HMP3_OK;
}

// Per-call output log of one stream, rebuilt from its frame records (one D2H copy of the records).
int hmp3_batch_call_log(hmp3_batch *b, int i, int32_t *frames_after_call, int64_t *bytes_after_call, int cap) {
    if (!b || !b->results_valid || i < 0 || i >= b->n || b->status[i] != HMP3_OK) {
        set_err("call_log: no completed run or bad index");
        return HMP3_ERR_ARG;
    }
    const int nrec = b->res_h[i].frames_recorded;
    if (nrec <= 0) return 0;
    std::vector<unsigned char> raw((size_t)nrec * sizeof_frame_rec());
    CK(cudaSetDevice(b->device));
    CK(cudaMemcpy(raw.data(), (const unsigned char *)b->d_frames + b->so_h[i].frames_off * sizeof_frame_rec(), raw.size(),
                  cudaMemcpyDeviceToHost));
    const FrameRec *fr = (const FrameRec *)raw.data();
    const EncConfig &C = b->tabs_h[b->st_h[i].cfg].cfg;
    const int per_call = (C.h_id == 1) ? 1 : 2;  // frames recorded per encode call
    std::vector<long long> end_off(nrec + 1, 0);
    for (int f = 0; f < nrec; f++) end_off[f + 1] = (long long)fr[f].out_off + 4 + C.side_bytes + fr[f].mf_bytes;
    const int ncalls = nrec / per_call;
    for (int c = 0; c < ncalls && c < cap; c++) {
        const int done = fr[(c + 1) * per_call - 1].done_after;
        if (frames_after_call) frames_after_call[c] = done;
        if (bytes_after_call) bytes_after_call[c] = end_off[done];
    }
    return ncalls;
}

int hmp3_batch_last_launches(const hmp3_batch *b) { return b ? b->launches : 0; }
int hmp3_batch_chunk_granules(const hmp3_batch *b) { return b ? b->NG : 0; }
float hmp3_batch_last_run_ms(const hmp3_batch *b) { return b ? b->last_run_ms : 0.0f; }

int hmp3_batch_phase_ms(const hmp3_batch *b, const char **names, float *ms, int *launches, int cap) {
    int k = 0;
    for (int p = 0; p < PH_COUNT && k < cap; p++, k++) {
        if (names) names[k] = kPhaseNames[p];
        if (ms) ms[k] = b->phase_ms[p];
        if (launches) launches[k] = b->phase_launches[p];
    }
    return k;
}

// Host buffers in, host buffers out: upload every stream's PCM, run, copy each stream's frames back.
namespace {
bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}
}  // namespace

static int encode_host_impl(hmp3_batch *b, const void *const *pcm, uint8_t *const *out, const int64_t *out_cap,
                            int64_t *out_bytes, int32_t *out_frames, int32_t *status);
int hmp3_batch_encode_host(hmp3_batch *b, const void *const *pcm, uint8_t *const *out, const int64_t *out_cap,
                           int64_t *out_bytes, int32_t *out_frames, int32_t *status) {
    if (!b || !pcm || !out || !out_cap) {
        set_err("bad arguments");
        return HMP3_ERR_ARG;
    }
    return drain_on_error(b, encode_host_impl(b, pcm, out, out_cap, out_bytes, out_frames, status));
}
static int encode_host_impl(hmp3_batch *b, const void *const *pcm, uint8_t *const *out, const int64_t *out_cap,
                           int64_t *out_bytes, int32_t *out_frames, int32_t *status) {
    CK(cudaSetDevice(b->device));
    const bool trace = getenv("HMP3_TRACE_HOST") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto stamp = [&](const char *what) {
        if (trace)
            fprintf(stderr, "[encode_host] %-28s %8.2f ms\n", what,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    // Pinned input: each chunk's samples are copied (DMA) right before that chunk's Phase A, overlapping the serial
    // stage of the previous chunk.  Pageable input: one queued copy per stream before the first kernel.
    bool pinned_in = true, pinned_out = true;
    for (int i = 0; i < b->n && (pinned_in || pinned_out); i++) {
        if (b->status[i] != HMP3_OK) continue;
        if (pinned_in && b->st_h[i].nsamples > 0 && !is_pinned(pcm[i])) pinned_in = false;
        if (pinned_out && !is_pinned(out[i])) pinned_out = false;
    }
    if (!pinned_in) {
        for (int i = 0; i < b->n; i++) {
            if (b->status[i] != HMP3_OK) continue;
            int r = upload_any(b, i, pcm[i], b->st_h[i].nsamples, b->fmt[i]);
            if (r != HMP3_OK) return r;
        }
    }
    stamp("pinned checks / uploads");
    // Pinned output with room for every stream's bound: finished frames leave the device chunk by chunk
    bool direct = pinned_out && !getenv("HMP3_NO_INCREMENTAL_OUT");
    for (int i = 0; i < b->n && direct; i++)
        if (b->status[i] == HMP3_OK && out_cap[i] < b->st_h[i].out_cap) direct = false;
    std::vector<long long> copied;
    if (direct) {
        const int chunks = (b->max_gran + 31) / 32 + 2;  // upper bound of the chunk count
        if (!b->d_out_inc) {
            CK(cudaMalloc(&b->d_out_inc, std::max<long long>(b->out_cap, 16)));
            CK(cudaMalloc(&b->d_done_lo, sizeof(int) * b->n));
            CK(cudaMalloc(&b->d_bytes_done, sizeof(long long) * b->n));
            CK(cudaStreamCreate(&b->stream_o));
        }
        if (b->h_prog_chunks < chunks) {
            if (b->h_prog) cudaFreeHost(b->h_prog);
            b->h_prog = nullptr;
            b->h_prog_chunks = 0;
            CK(cudaMallocHost(&b->h_prog, sizeof(long long) * (size_t)chunks * b->n));
            b->h_prog_chunks = chunks;
            while ((int)b->ev_o.size() < chunks) {
                cudaEvent_t e;
                CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                b->ev_o.push_back(e);
            }
        }
        copied.assign(b->n, 0);
    }
    stamp("output set-up");
    b->direct = direct;
    b->staged = pinned_in;
    b->h_src = pcm;
    int r = run_plan(b);
    stamp("run enqueued");
    b->staged = false;
    b->h_src = nullptr;
    b->direct = false;
    if (r != HMP3_OK) return r;
    if (direct) {  // follow the run: as each chunk's progress row lands, queue the copies of the bytes it completed
        const int rows = std::min(b->chunks_run, b->h_prog_chunks);
        for (int c = 0; c < rows; c++) {
            CK(cudaEventSynchronize(b->ev_o[c]));
            const long long *row = b->h_prog + (size_t)c * b->n;
            b->cp_dst.clear();
            b->cp_src.clear();
            b->cp_len.clear();
            for (int i = 0; i < b->n; i++) {
                if (b->status[i] != HMP3_OK) continue;
                const long long hi = std::min<long long>(row[i], out_cap[i]);
                if (hi > copied[i]) {
                    b->cp_dst.push_back(out[i] + copied[i]);
                    b->cp_src.push_back(b->d_out_inc + b->st_h[i].out_off + copied[i]);
                    b->cp_len.push_back((size_t)(hi - copied[i]));
                    copied[i] = hi;
                }
            }
            bool batched = false;
            if (!b->cp_len.empty() && !b->no_batch_copy) {
                cudaMemcpyAttributes at;
                memset(&at, 0, sizeof(at));
                at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
                at.flags = cudaMemcpyFlagPreferOverlapWithCompute;
                size_t idx0 = 0, fail = 0;
                if (cudaMemcpyBatchAsync(b->cp_dst.data(), b->cp_src.data(), b->cp_len.data(), b->cp_len.size(), &at, &idx0,
                                         1, &fail, b->stream_o) == cudaSuccess)
                    batched = true;
                else {
                    cudaGetLastError();
                    b->no_batch_copy = true;
                }
            }
            if (!batched)
                for (size_t q = 0; q < b->cp_len.size(); q++)
                    CK(cudaMemcpyAsync(b->cp_dst[q], b->cp_src[q], b->cp_len[q], cudaMemcpyDeviceToHost, b->stream_o));
        }
    }
    r = sync_plan(b);
    if (r != HMP3_OK) return r;
    stamp("run finished");
    const long long total = b->out_off_h[b->n];
    if (!pinned_out) {  // one bulk D2H into pinned staging, then scatter to the callers' pageable buffers
        if (b->h_stage_bytes < total) {
            if (b->h_stage) cudaFreeHost(b->h_stage);
            b->h_stage = nullptr;
            b->h_stage_bytes = 0;
            CK(cudaMallocHost(&b->h_stage, total + (total >> 3) + 4096));
            b->h_stage_bytes = total + (total >> 3) + 4096;
        }
        CK(cudaMemcpyAsync(b->h_stage, b->d_out, total, cudaMemcpyDeviceToHost, b->stream));
        CK(cudaStreamSynchronize(b->stream));
    }
    for (int i = 0; i < b->n; i++) {
        int st = b->status[i];
        long long nb = st == HMP3_OK ? b->res_h[i].out_bytes : 0;
        if (st == HMP3_OK && !b->res_h[i].finished) st = HMP3_ERR_OUT_SPACE;
        if (st == HMP3_OK && b->flags_h[i]) st = HMP3_ERR_INTERNAL;
        if (st == HMP3_OK && nb > out_cap[i]) {
            st = HMP3_ERR_OUT_SPACE;
            nb = 0;
        }
        if (nb) {
            if (pinned_out) {  // straight into the caller's pinned buffer (whatever the incremental copies left)
                const long long have = direct ? std::min(copied[i], nb) : 0;
                if (nb > have)
                    CK(cudaMemcpyAsync(out[i] + have, b->d_out + b->out_off_h[i] + have, nb - have, cudaMemcpyDeviceToHost,
                                       b->stream));
            }
            else memcpy(out[i], b->h_stage + b->out_off_h[i], nb);
        }
        if (out_bytes) out_bytes[i] = nb;
        if (out_frames) out_frames[i] = st == HMP3_OK ? b->res_h[i].frames : 0;
        if (status) status[i] = st;
    }
    if (pinned_out) CK(cudaStreamSynchronize(b->stream));
    if (direct) CK(cudaStreamSynchronize(b->stream_o));
    stamp("results / residual copies");
    return HMP3_OK;
}

int hmp3_encode_batch(hmp3_stream_desc *streams, int n, int device) {
    if (n <= 0 || !streams) {
        set_err("bad arguments");
        return HMP3_ERR_ARG;
    }
    std::vector<hmp3_control> ctl(n);
    std::vector<int64_t> ns(n), cap(n), nb(n);
    std::vector<const void *> pcm(n);
    std::vector<int32_t> fmts(n);
    std::vector<uint8_t *> out(n);
    std::vector<int32_t> nf(n), st(n);
    for (int i = 0; i < n; i++) {
        if (!streams[i].control) {
            set_err("stream without a control block");
            return HMP3_ERR_ARG;
        }
        ctl[i] = *streams[i].control;
        ns[i] = streams[i].num_samples;
        pcm[i] = streams[i].pcm;
        fmts[i] = streams[i].pcm_format;
        out[i] = streams[i].out;
        cap[i] = streams[i].out_capacity;
    }
    hmp3_batch *b = hmp3_batch_create_ex(ctl.data(), ns.data(), fmts.data(), n, device);
    if (!b) {  // the reason plan_create gave: no device, bad control, allocation failure ...
        const int ndev = hmp3_device_count();
        if (ndev <= device) return HMP3_ERR_NO_DEVICE;
        return g_create_status != HMP3_OK ? g_create_status : HMP3_ERR_BAD_CONTROL;
    }
    int r = hmp3_batch_encode_host(b, pcm.data(), out.data(), cap.data(), nb.data(), nf.data(), st.data());
    if (r == HMP3_OK)
        for (int i = 0; i < n; i++) {
            streams[i].out_bytes = nb[i];
            streams[i].out_frames = nf[i];
            streams[i].status = st[i];
        }
    hmp3_batch_destroy(b);
    return r;
}


// ---------------------------------------------------------------------------------------------
// (2) CMp3Enc mirrors: one stream per handle, two granules per call through the same kernels.
// ---------------------------------------------------------------------------------------------
}  // extern "C"

struct hmp3_encoder {
    int device = 0;
    hmp3_batch *b = nullptr;
    int nch = 0;
    int calls = 0;              // encode calls made inside the current window of the device buffers
    int window_calls = 0;       // calls the window holds; when it is full the handle rebases (encoder_rebase)
    int frames_out = 0;         // CMp3Enc::tot_frames_out / tot_bytes_out: the whole stream
    long long bytes_out = 0;
    int win_frames_out = 0;     // the same counters relative to the current window (what the device state counts)
    long long win_bytes_out = 0;
    int win_recorded = 0;       // frames the serial stage has recorded in the window
    int ave_bytes = 0;          // CMp3Enc::ave_tot_bytes_out: running average of the bytes per call, x256
    bool float_in = false;      // the plan takes float PCM (every input format except 16-bit integer)
    int src_bits = 16, src_float = 0;
    int src_chan = 0;           // channels of the caller's PCM (2 with nch == 1: down-mix to mono, Csrc kfilter 2)
    bool up2 = false;           // 1:2 up-conversion (Csrc case 1)
    bool resample = false;      // general conversion (Csrc cases 2-4, resample.h)
    hmp3::Resampler rs;
    int frames_in = 1152;       // sample frames of the caller's PCM consumed per call
    std::vector<float> conv;    // the call's samples after sr_convert's type conversion
    int capacity_seconds = 20;
    std::vector<float> stage;
    ~hmp3_encoder() { delete b; }
};

namespace {
int encoder_init(hmp3_encoder *e, const hmp3_control *ec, bool float_in) {
    delete e->b;
    e->b = nullptr;
    e->calls = e->frames_out = e->win_frames_out = e->win_recorded = 0;
    e->bytes_out = e->win_bytes_out = 0;
    e->ave_bytes = 0;
    e->float_in = float_in;
    EncTables *T = new EncTables;
    int unsup = 0;
    const int bytes_in = build_tables(ec, T, &unsup);
    const int samprate = T->cfg.samprate;
    delete T;
    if (!bytes_in) {
        set_err(unsup ? "configuration outside the built path" : "control block rejected");
        return 0;
    }
    e->b = new hmp3_batch;
    long long ns = (long long)e->capacity_seconds * samprate;
    ns -= ns % 1152;
    if (ns < 16 * 1152) ns = 16 * 1152;
    const int fmt = float_in ? 1 : 0;
    if (plan_create(e->b, ec, &ns, 1, e->device, 2, false, true, &fmt) != HMP3_OK) {
        delete e->b;
        e->b = nullptr;
        return 0;
    }
    hmp3_batch *b = e->b;
    e->nch = b->st_h[0].nch;
    e->window_calls = (int)(ns / 1152);
    e->stage.assign((size_t)1152 * e->nch, 0.0f);
    cudaSetDevice(e->device);
    if (plan_reset_state(b) != HMP3_OK) return 0;
    cudaMemsetAsync(b->d_flags, 0, sizeof(int), b->stream);
    launch_rate_init(b->d_tabs, b->d_st, b->d_rs, b->d_rs_cold, 1, b->stream);
    if (cudaStreamSynchronize(b->stream) != cudaSuccess) {
        set_err("device initialisation failed");
        return 0;
    }
    return bytes_in;
}

// The window of the handle's device buffers is full: keep what the next calls still need -- the last kKeepCalls
// calls of PCM (polyphase history and the three-granule look-back of the hybrid transform), the frames whose
// main-data slot is not complete yet and their main data -- move it to the start of the buffers and shift every
// absolute index (granule, frame, main-data and output offsets) by the same amounts.  Nothing else of the state is
// positional, so the stream continues exactly as if the buffers were unbounded.
constexpr int kKeepCalls = 4;
int encoder_rebase(hmp3_encoder *e) {
    hmp3_batch *b = e->b;
    const StreamDev &sd = b->st_h[0];
    const long long s0 = (long long)(e->calls - kKeepCalls) * 1152, ns = (long long)kKeepCalls * 1152;
    size_t el;
    char *base = pcm_dev_ptr(b, 0, 0, &el);
    CK(cudaMemcpyAsync(base, base + el * s0 * sd.nch, el * ns * sd.nch, cudaMemcpyDeviceToDevice, b->stream));
    if (sd.pcmf_off >= 0)  // the DC-filtered copy (-S1) is indexed by sample as well
        CK(cudaMemcpyAsync(b->d_pcmf + sd.pcmf_off, b->d_pcmf + sd.pcmf_off + s0 * sd.nch, sizeof(float) * ns * sd.nch,
                           cudaMemcpyDeviceToDevice, b->stream));
    launch_handle_rebase(b->d_rs, b->d_frames, b->d_main, 2 * (e->calls - kKeepCalls), b->d_res, b->stream);
    CK(cudaMemcpyAsync(b->res_h.data(), b->d_res, sizeof(StreamResult), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    e->calls = kKeepCalls;
    e->win_frames_out = 0;
    e->win_bytes_out = 0;
    e->win_recorded = b->res_h[0].frames_recorded;
    return HMP3_OK;
}

// one encode call: 1152 samples per channel in, whatever frames became complete out; with `packet` also the frame(s)
// this call produced as self-contained packets (CMp3Enc::L3_audio_encode_*Packet, mp3enc.cpp:2868-3440)
hmp3_in_out encoder_step(hmp3_encoder *e, const void *pcm, unsigned char *bs_out, bool want_bs = true,
                         unsigned char *packet = nullptr, int *nbytes_out = nullptr) {
    hmp3_in_out io = {0, 0};
    hmp3_batch *b = e->b;
    if (!b) {
        set_err("encoder not initialised");
        return io;
    }
    cudaSetDevice(e->device);
    if (e->calls >= e->window_calls && encoder_rebase(e) != HMP3_OK) return io;
    const int K0 = 2 * e->calls;
    size_t el;
    char *dst = pcm_dev_ptr(b, 0, (long long)e->calls * 1152, &el);
    const size_t nb_in = el * 1152 * e->nch;
    if (cudaMemcpyAsync(dst, pcm, nb_in, cudaMemcpyHostToDevice, b->stream) != cudaSuccess) {
        set_err("H2D copy failed");
        return io;
    }
    if (launch_analysis(b, K0, 0, b->stream) != HMP3_OK) return io;
    if (b->any_allo1) launch_rate_a1(b->d_tabs, b->d_st, b->d_so, b->d_rs, b->cb2[0], b->d_main, b->d_frames, K0, 1, b->stream);
    else launch_rate(b->d_tabs, b->d_st, b->d_so, b->d_rs, b->cb2[0], b->d_main, b->d_frames, K0, 1, b->stream);
    launch_pack(b->d_tabs, b->d_st, b->d_so, b->cb2[0], b->d_main, b->d_frames, b->d_flags, K0, 1, b->stream);
    launch_finish(b->d_tabs, b->d_st, b->d_so, b->d_rs, b->d_frames, b->d_res, b->d_out_off, b->d_main, b->d_out, 48, 1,
                  b->stream, nullptr, e->win_frames_out, e->win_bytes_out);
    cudaMemcpyAsync(b->res_h.data(), b->d_res, sizeof(StreamResult), cudaMemcpyDeviceToHost, b->stream);
    if (cudaStreamSynchronize(b->stream) != cudaSuccess) {
        set_err(std::string("device error: ") + cudaGetErrorString(cudaGetLastError()));
        return io;
    }
    const StreamResult &r = b->res_h[0];
    const EncConfig &C = b->tabs_h[0].cfg;
    const long long nb = r.out_bytes - e->win_bytes_out;
    if (nb > 0 && want_bs) cudaMemcpy(bs_out, b->d_out, nb, cudaMemcpyDeviceToHost);
    if (nbytes_out) nbytes_out[0] = nbytes_out[1] = 0;
    if (packet && nbytes_out) {
        // the frames recorded by this call: header (CBR form: L3_pack_head, also in VBR mode) | side information with
        // main_data_begin still zero | the frame's own scale-factor and Huffman bytes
        const int nf = r.frames_recorded - e->win_recorded;
        std::vector<FrameRec> fr(nf > 0 ? nf : 0);
        if (nf > 0)
            cudaMemcpy(fr.data(), b->d_frames + (r.frames_recorded - nf), sizeof(FrameRec) * nf, cudaMemcpyDeviceToHost);
        for (int k = 0; k < nf && k < 2; k++) {
            unsigned char *p = packet;
            memcpy(p, fr[k].head, 4);
            if (C.vbr_flag) p[2] = C.head[2];
            memcpy(p + 4, fr[k].side, C.side_bytes);
            p[4] = 0;
            if (C.h_id == 1) p[5] &= 0x7F;
            const int bytes = (fr[k].data_bits + 7) >> 3;
            if (bytes > 0) cudaMemcpy(p + 4 + C.side_bytes, b->d_main + fr[k].data_start, bytes, cudaMemcpyDeviceToHost);
            nbytes_out[k] = 4 + C.side_bytes + bytes;
            packet += nbytes_out[k];
        }
    }
    const long long counted = want_bs ? nb : 0;  // with bs_out == NULL the reference counts frames but no bytes
    e->frames_out += r.frames - e->win_frames_out;
    e->bytes_out += counted;
    e->win_frames_out = r.frames;
    e->win_bytes_out = r.out_bytes;
    e->win_recorded = r.frames_recorded;
    // CMp3Enc::ave_tot_bytes_out (mp3enc.cpp:2209, 2316: >> 7 per MPEG-1 call; :2476, 2590: >> 6 per MPEG-2 call)
    e->ave_bytes += (int)(((int)counted << 8) - e->ave_bytes) >> (C.h_id == 1 ? 7 : 6);
    e->calls++;
    io.in_bytes = (int)nb_in;
    io.out_bytes = (int)counted;
    return io;
}
}  // namespace

extern "C" {

hmp3_encoder *hmp3_encoder_new(int device) {
    if (hmp3_device_count() <= device) {
        set_err("no usable CUDA device (this library has no CPU path)");
        return nullptr;
    }
    hmp3_encoder *e = new hmp3_encoder;
    e->device = device;
    return e;
}
void hmp3_encoder_delete(hmp3_encoder *e) { delete e; }
int hmp3_encoder_set_capacity_seconds(hmp3_encoder *e, int seconds) {
    if (!e || seconds < 1) return HMP3_ERR_ARG;
    e->capacity_seconds = seconds;
    return HMP3_OK;
}

int hmp3_MP3_audio_encode_init(hmp3_encoder *e, const hmp3_control *ec, int source_bits, int source_is_float,
                               int mpeg_select, int mono_convert) {
    if (!e || !ec) return 0;
    // the encode rate as CMp3Enc::MP3_audio_encode_init picks it from the source rate and mpeg_select
    // (mp3enc.cpp:2628-2651, 2683-2748): track the input / an MPEG-1 rate / an MPEG-2 rate / the rate given
    auto nearest = [](const int *t, int n, int x) {
        int best = t[0], d0 = abs(t[0] - x);
        for (int i = 0; i < n; i++)
            if (abs(t[i] - x) < d0) {
                d0 = abs(t[i] - x);
                best = t[i];
            }
        return best;
    };
    static const int rates[6] = {22050, 24000, 16000, 44100, 48000, 32000};
    const int source = ec->samprate;
    if (source < 4000 || source > 48000) return 0;
    if (mpeg_select < 0) mpeg_select = 0;
    int target = 0, t2;
    switch (mpeg_select) {
    case 0:
        if (source < 16000 && (t2 = nearest(rates, 3, 2 * source)) == 2 * source) target = t2;
        else target = nearest(rates, 6, source);
        break;
    case 1:
        if (source < 16000 && (t2 = nearest(rates + 3, 3, 4 * source)) == 4 * source) target = t2;
        else if (source < 32000 && (t2 = nearest(rates + 3, 3, 2 * source)) == 2 * source) target = t2;
        else target = nearest(rates + 3, 3, source);
        break;
    case 2:
        if (source < 16000 && (t2 = nearest(rates, 3, 2 * source)) == 2 * source) target = t2;
        else if (source > 24000 && 2 * (t2 = nearest(rates, 3, source / 2)) == source) target = t2;
        else target = nearest(rates, 3, source);
        break;
    default:
        target = nearest(rates, 6, mpeg_select);
        if (target != mpeg_select) return 0;
    }
    const bool native = target == source;
    const bool up2 = target == 2 * source;
    const bool fmt_ok = source_is_float ? (source_bits == 32)
                                        : (source_bits == 8 || source_bits == 16 || source_bits == 24 || source_bits == 32);
    // Csrc::sr_convert_init's own limits (srcc.cpp:741-756)
    if (!fmt_ok || source < 8000 || target < 5000 || target > 50400) {
        set_err("MP3_audio_encode_init: sample format or rate outside what the converter takes");
        return 0;
    }
    e->resample = false;
    if (!native && !up2) {  // general conversion (Csrc cases 2-4)
        if (e->rs.init(source, target) <= 0 || e->rs.ncase < 2) {
            set_err("MP3_audio_encode_init: no conversion from this source rate to the encode rate");
            return 0;
        }
        e->resample = true;
    }
    // channels as CMp3Enc::MP3_audio_encode_init derives them (mp3enc.cpp:2689-2696): the source has two unless the
    // mode is mono; mono_convert encodes a two-channel source as mono (Csrc kfilter 2)
    hmp3_control ec2 = *ec;
    e->src_chan = ec->mode == 3 ? 1 : 2;
    const bool downmix = mono_convert && e->src_chan == 2;
    if (downmix) ec2.mode = 3;
    e->up2 = up2;
    ec2.samprate = target;
    if (source < target) {  // band limit of an up-converted signal (mp3enc.cpp:2765-2787; cutoff: srcc.cpp:780-781)
        const int cutoff = (int)(0.90f * source / 2);
        int nsb = (64 * cutoff + target / 2) / target;
        if (nsb > 30) nsb = 30;
        if (ec2.nsb_limit <= 0) ec2.nsb_limit = 30;
        if (ec2.nsb_limit > nsb) ec2.nsb_limit = nsb;
    }
    e->src_bits = source_bits;
    e->src_float = source_is_float;
    const int bytes_in = encoder_init(e, &ec2, up2 || e->resample || downmix || !(source_bits == 16 && !source_is_float));
    if (!bytes_in) return 0;
    // what Csrc::sr_convert_init returns: the sample frames that must be buffered for a call (srcc.cpp:185-187, 769-773):
    // the 1152 (576 when up-converting) it consumes plus one
    if (e->resample) {  // the converter's own figure: the source frames a call may look at
        e->frames_in = e->rs.minbuf;
        return e->frames_in * e->src_chan * (source_bits / 8);
    }
    e->frames_in = up2 ? 576 : 1152;
    return (e->frames_in + 1) * e->src_chan * (source_bits / 8);
}
}  // extern "C"
namespace {
hmp3_in_out mp3_encode_call(hmp3_encoder *e, const unsigned char *pcm, unsigned char *bs_out, bool want_bs,
                            unsigned char *packet, int *nbytes_out) {
    hmp3_in_out io = {0, 0};
    if (!e || !e->b) {
        set_err("encoder not initialised");
        return io;
    }
    if (!e->float_in) {
        io = encoder_step(e, pcm, bs_out, want_bs, packet, nbytes_out);
        return io;
    }
    // sample conversion of Csrc::sr_convert (hmp3/src/srcc.cpp:804-834): everything becomes float on a +-32768 scale.
    // Up-conversion looks one sample frame past the 576 it consumes (the caller buffers frames_in + 1, see init).
    const bool downmix = e->src_chan == 2 && e->nch == 1;
    const hmp3::Resampler::Layout lay = e->src_chan == 1 ? hmp3::Resampler::MONO
                                        : (downmix ? hmp3::Resampler::TO_MONO : hmp3::Resampler::DUAL);
    // the frames this call reads: the general converter looks as far as the reference's does (Resampler::reach)
    const int nfr = e->resample ? e->rs.reach(lay) : e->frames_in + (e->up2 ? 1 : 0);
    const int n = nfr * e->src_chan;
    if ((int)e->conv.size() < n) e->conv.resize(n);
    float *d = e->conv.data();
    if (e->src_float) {
        const float *s = (const float *)pcm;
        for (int i = 0; i < n; i++) d[i] = (float)(s[i]) * 32768.0f;
    } else if (e->src_bits == 32) {
        const int *s = (const int *)pcm;
        for (int i = 0; i < n; i++) d[i] = (float)(s[i] / 65536.0f);
    } else if (e->src_bits == 24) {
        const unsigned char *s = pcm;
        for (int i = 0; i < n; i++, s += 3) {
            const int v = (int)(((unsigned)s[2] << 24) | ((unsigned)s[1] << 16) | ((unsigned)s[0] << 8)) >> 8;
            d[i] = (float)((float)v / 256.0f);
        }
    } else if (e->src_bits == 16) {
        const short *s = (const short *)pcm;
        for (int i = 0; i < n; i++) d[i] = (float)s[i];
    } else {  // 8-bit unsigned
        for (int i = 0; i < n; i++) d[i] = (((float)pcm[i]) - 128.0f) * (256.0f);
    }
    if ((int)e->stage.size() < 1152 * e->nch) e->stage.resize((size_t)1152 * e->nch);
    float *y = e->stage.data();
    if (e->resample) {  // Csrc cases 2-4: 1152 frames at the encode rate from as many source frames as that takes
        const int used = e->rs.run(lay, d, y);
        io = encoder_step(e, y, bs_out, want_bs, packet, nbytes_out);
        if (io.in_bytes) io.in_bytes = used * e->src_chan * (e->src_bits / 8);
        return io;
    }
    if (!e->up2) {
        if (downmix)  // src_filter_to_mono_case0 (hmp3/src/srccf.cpp:458-468)
            for (int i = 0; i < 1152; i++) y[i] = (float)((d[2 * i] + d[2 * i + 1]) * 0.5);
        else
            for (int i = 0; i < n; i++) y[i] = d[i];
    } else if (e->src_chan == 1) {  // src_filter_mono_case1 (srccf.cpp:80-100): integer samples, truncated
        for (int i = 0; i < 576; i++) {
            const int a = (int)d[i], b2 = (int)d[i + 1];
            y[2 * i] = (float)a;
            y[2 * i + 1] = (float)((a + b2) >> 1);
        }
    } else if (downmix) {  // src_filter_to_mono_case1 (srccf.cpp:472-492)
        for (int i = 0; i < 576; i++) {
            const float a = d[2 * i] + d[2 * i + 1], b2 = d[2 * i + 2] + d[2 * i + 3];
            y[2 * i] = (float)(a * 0.5);
            y[2 * i + 1] = (float)((a + b2) * 0.25);
        }
    } else {  // src_filter_dual_case1 (srccf.cpp:258-276)
        for (int i = 0; i < 576; i++)
            for (int c = 0; c < 2; c++) {
                y[2 * (2 * i) + c] = d[2 * i + c];
                y[2 * (2 * i + 1) + c] = (float)((d[2 * i + c] + d[2 * i + 2 + c]) * 0.5);
            }
    }
    io = encoder_step(e, y, bs_out, want_bs, packet, nbytes_out);
    if (io.in_bytes) io.in_bytes = e->frames_in * e->src_chan * (e->src_bits / 8);
    return io;
}
}  // namespace
extern "C" {
hmp3_in_out hmp3_MP3_audio_encode(hmp3_encoder *e, const unsigned char *pcm, unsigned char *bs_out) {
    return mp3_encode_call(e, pcm, bs_out, true, nullptr, nullptr);
}
hmp3_in_out hmp3_MP3_audio_encode_Packet(hmp3_encoder *e, const unsigned char *pcm, unsigned char *bs_out,
                                         unsigned char *packet, int nbytes_out[2]) {
    return mp3_encode_call(e, pcm, bs_out, bs_out != nullptr, packet, nbytes_out);
}
hmp3_in_out hmp3_L3_audio_encode_Packet(hmp3_encoder *e, const float *pcm, unsigned char *bs_out, unsigned char *packet,
                                        int nbytes_out[2]) {
    hmp3_in_out io = {0, 0};
    if (!e || !e->b || !e->float_in) {
        set_err("encoder not initialised for float input");
        return io;
    }
    return encoder_step(e, pcm, bs_out, bs_out != nullptr, packet, nbytes_out);
}
int hmp3_L3_audio_encode_init(hmp3_encoder *e, const hmp3_control *ec) {
    if (!e || !ec) return 0;
    e->src_bits = 32;
    e->src_float = 1;
    return encoder_init(e, ec, true);
}
hmp3_in_out hmp3_L3_audio_encode(hmp3_encoder *e, const float *pcm, unsigned char *bs_out) {
    hmp3_in_out io = {0, 0};
    if (!e || !e->b || !e->float_in) {
        set_err("encoder not initialised for float input");
        return io;
    }
    return encoder_step(e, pcm, bs_out);  // PCM already on the +-32768 scale (hmp3/src/pub/mp3enc.h:88-98)
}
void hmp3_L3_audio_encode_info_ec(hmp3_encoder *e, hmp3_control *ec) {
    if (e && e->b && ec) memcpy(ec, e->b->tabs_h[0].cfg.info_ec, sizeof(*ec));
}
void hmp3_L3_audio_encode_info_head(hmp3_encoder *e, hmp3_mpeg_head *h) {
    if (!e || !e->b || !h) return;
    const EncConfig &C = e->b->tabs_h[0].cfg;
    memset(h, 0, sizeof(*h));
    h->sync = 1;  // setup_header (setup.c:191-289)
    h->id = C.h_id;
    h->option = 1;
    h->prot = 1;
    h->br_index = C.br_index;
    h->sr_index = C.sr_index;
    h->mode = C.h_mode;
    h->mode_ext = (C.head[3] >> 4) & 3;
    h->cr = (C.head[3] >> 3) & 1;
    h->original = (C.head[3] >> 2) & 1;
}
void hmp3_L3_audio_encode_info_string(hmp3_encoder *e, char *s) {  // mpeg_info_string (setup.c:336-366)
    if (!e || !e->b || !s) return;
    const EncConfig &C = e->b->tabs_h[0].cfg;
    const char *mode_msg[] = {"mode 0 STEREO", "mode 1 STEREO", "DUAL", "MONO"};
    s += sprintf(s, "Layer %s ", "III");
    s += sprintf(s, "  %s ", mode_msg[C.h_mode & 3]);
    if (C.h_mode == 1 && C.info_nsbstereo < 32) s += sprintf(s, " IS-%d ", C.info_nsbstereo);
    s += sprintf(s, "  %ldHz ", (long)C.samprate);
    if (!C.vbr_flag) s += sprintf(s, "  %dkbps ", C.totbitrate);
    else {
        s += sprintf(s, " VBR-%d", C.vbr_mnr);
        if (C.vbr_delta_mnr) s += sprintf(s, "(%d)", C.vbr_delta_mnr);
    }
    if (C.hf_flag) {
        s += sprintf(s, "  hf");
        if (C.hf_flag & 2) s += sprintf(s, "2");
    }
}
unsigned int hmp3_L3_audio_encode_get_frames(hmp3_encoder *e) { return e ? (unsigned)e->frames_out : 0; }
float hmp3_L3_audio_encode_get_bitrate_float(hmp3_encoder *e) {  // mp3enc.cpp:3450-3463
    if (!e || !e->b || e->frames_out <= 0) return 0.0f;
    const EncConfig &C = e->b->tabs_h[0].cfg;
    const float samples = (C.h_id == 1) ? 1152.0f : 576.0f;
    return ((0.001f * 8.0f) * e->bytes_out * C.samprate / (samples * e->frames_out));
}
float hmp3_L3_audio_encode_get_bitrate2_float(hmp3_encoder *e) {  // mp3enc.cpp:3466-3480 (recent average)
    if (!e || !e->b || e->frames_out <= 0) return 0.0f;
    const EncConfig &C = e->b->tabs_h[0].cfg;
    return (float)((0.001f * 8.0f / (1152.0 * 256.0)) * e->ave_bytes * C.samprate);
}
hmp3_int_pair hmp3_L3_audio_encode_get_frames_bytes(hmp3_encoder *e) {  // mp3enc.cpp:3513-3521
    hmp3_int_pair x = {0, 0};
    if (e) {
        x.a = e->frames_out;
        x.b = (int)e->bytes_out;
    }
    return x;
}
int hmp3_L3_audio_encode_get_bitrate(hmp3_encoder *e) { return (int)(hmp3_L3_audio_encode_get_bitrate_float(e) + 0.5f); }
int hmp3_control_apply_option(hmp3_control *ec, const char *opt) { return control_apply_option(ec, opt); }

// Debug / parity entry: Phase A of ONE stream on the device, stage outputs copied back to the host.
// Same argument meaning as the host simulator's sim_analysis (tests/hostsim/hostsim.cpp).
int hmp3_debug_fp32_peak(int device, float *ffma_tflops, float *nonfused_tflops) {
    if (hmp3_device_count() <= device) return HMP3_ERR_NO_DEVICE;
    return fp32_peak(device, ffma_tflops, nonfused_tflops) == 0 ? HMP3_OK : HMP3_ERR_CUDA;
}

int hmp3_debug_timeline(const hmp3_batch *b, float *rows, int cap) {
    int k = 0;
    for (; k < (int)b->timeline.size() && k < cap; k++)
        for (int j = 0; j < 3; j++) rows[3 * k + j] = b->timeline[k][j];
    return k;
}

int hmp3_debug_resample(int source, int target, int layout, const float *x, int ncalls, float *y, int *used) {
    // the handle's sample-rate converter on its own (host code, no device): tests compare it with the reference's Csrc
    hmp3::Resampler r;
    const int minbuf = r.init(source, target);
    if (minbuf <= 0 || r.ncase < 2) return minbuf <= 0 ? 0 : -r.ncase;
    const int fw = layout == 0 ? 1 : 2, ow = layout == 1 ? 2 : 1;
    long off = 0;
    std::vector<float> stage;
    for (int k = 0; k < ncalls; k++) {
        // staged the way the handle stages a call: reach() frames and not one more (what lies behind them is poisoned)
        const int nfr = r.reach((hmp3::Resampler::Layout)layout);
        stage.assign((size_t)(nfr + 64) * fw, std::numeric_limits<float>::quiet_NaN());
        memcpy(stage.data(), x + off * fw, sizeof(float) * nfr * fw);
        used[k] = r.run((hmp3::Resampler::Layout)layout, stage.data(), y + (long)k * 1152 * ow);
        off += used[k];
    }
    return minbuf;
}

int hmp3_debug_set_rate_tap(hmp3_batch *b, int stream, void *records, long long cap_records) {
    if (!b || (records && (stream < 0 || stream >= b->n || cap_records <= 0))) return HMP3_ERR_ARG;
    b->tap_stream = records ? stream : -1;
    b->tap_out = (unsigned char *)records;
    b->tap_cap = records ? cap_records : 0;
    return HMP3_OK;
}
int hmp3_debug_rate_tap_record_bytes(void) { return (int)sizeof_pack_gc(); }

int hmp3_debug_rate_cycles(hmp3_batch *b, long long *cycles, int max_launches) {
    // first call (cycles == NULL or nothing recorded yet): switch recording on for the following runs
    if (!b) return HMP3_ERR_ARG;
    CK(cudaSetDevice(b->device));
    if (!b->d_cycles) {
        CK(cudaMalloc(&b->d_cycles, sizeof(long long) * kCycleLaunches * b->n));
        CK(cudaMemset(b->d_cycles, 0, sizeof(long long) * kCycleLaunches * b->n));
        return 0;
    }
    const int m = std::min(max_launches, b->cycle_launches);
    if (cycles && m > 0)
        CK(cudaMemcpy(cycles, b->d_cycles, sizeof(long long) * m * b->n, cudaMemcpyDeviceToHost));
    return m;
}

int hmp3_debug_analysis(const hmp3_control *ec, const int16_t *pcm, long long nsamples, int ngran, int device,
                        float *sbt_out, int *ginfo, float *xr_out, float *raw_out, int *ms_raw, int *att) {
    hmp3_batch b;
    int r = plan_create(&b, ec, &nsamples, 1, device, 32, true);
    if (r != HMP3_OK) return r;
    if (b.status[0] != HMP3_OK) return b.status[0];
    const int nch = b.st_h[0].nch;
    b.st_h[0].ngran = ngran;
    CK(cudaMemcpy(b.d_st, b.st_h.data(), sizeof(StreamDev), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b.d_pcm, pcm, sizeof(int16_t) * nsamples * nch, cudaMemcpyHostToDevice));
    r = plan_reset_state(&b);
    if (r != HMP3_OK) return r;
    const int NG = b.NG, G = NG + 3;
    std::vector<float> P((size_t)G * 2 * 576), X((size_t)NG * 2 * 576);
    std::vector<int> E((size_t)G * 2 * 9), M(NG);
    std::vector<GranuleInfo> GI(NG);
    std::vector<PsyRaw> RW((size_t)NG * 2);
    for (int K0 = 0; K0 < ngran; K0 += NG) {
        r = launch_analysis(&b, K0, 0, b.stream, false);  // the taps want xr before the in-place prologue
        if (r != HMP3_OK) return r;
        CK(cudaStreamSynchronize(b.stream));
        CK(cudaMemcpy(P.data(), b.cb.P, P.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(E.data(), b.cb.E, E.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(X.data(), b.cb.xr, X.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(M.data(), b.cb.ms_raw, M.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(GI.data(), b.cb.gi, GI.size() * sizeof(GranuleInfo), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(RW.data(), b.cb.raw, RW.size() * sizeof(PsyRaw), cudaMemcpyDeviceToHost));
        for (int q = 0; q < NG && K0 + q < ngran; q++) {
            int K = K0 + q;
            if (ginfo) memcpy(ginfo + 4 * K, &GI[q], 16);
            if (ms_raw) ms_raw[K] = M[q];
            for (int c = 0; c < nch; c++) {
                // P[K] lives at slot q+3 of this chunk only if K <= K0+NG-2; the last granule of a chunk is
                // produced by the next chunk (slot 2).  Report P[K-1] instead (slot q+2) for K >= 1.
                if (sbt_out && K >= 1)
                    memcpy(sbt_out + ((size_t)(K - 1) * nch + c) * 576, &P[((size_t)(q + 2) * 2 + c) * 576], 576 * 4);
                if (att && K >= 1) memcpy(att + ((size_t)(K - 1) * nch + c) * 9, &E[((size_t)(q + 2) * 2 + c) * 9], 36);
                if (xr_out) memcpy(xr_out + ((size_t)K * nch + c) * 576, &X[((size_t)q * 2 + c) * 576], 576 * 4);
                if (raw_out) memcpy(raw_out + ((size_t)K * nch + c) * 92, &RW[(size_t)q * 2 + c], sizeof(PsyRaw));
            }
        }
    }
    return HMP3_OK;
}

}  // extern "C"
