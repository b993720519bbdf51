// Host side of the device pipeline: batch plan (buffers sized for a batch shape), chunked kernel
// launches, and the C ABI of include/hmp3_b200.h.  No CPU fallback exists: without a CUDA device
// every entry point fails.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/hmp3_b200.h"
#include "../../include/hmp3_b200_debug.h"
#include "enc_init.h"
#include "kernels_analysis.cuh"

using namespace hmp3;

namespace {

thread_local std::string g_err;
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

}  // namespace

struct hmp3_batch {
    int device = 0;
    int n = 0;
    int NG = 0;
    int max_gran = 0;
    std::vector<hmp3_control> controls;  // distinct
    std::vector<EncTables> tabs_h;
    std::vector<StreamDev> st_h;
    EncTables *d_tabs = nullptr;
    StreamDev *d_st = nullptr;
    int16_t *d_pcm = nullptr;
    long long pcm_elems = 0;
    SwitchState *d_sw = nullptr;
    ChunkBufs cb{};
    cudaStream_t stream = nullptr;
    int launches = 0;
    std::vector<int> status;

    ~hmp3_batch() {
        cudaSetDevice(device);
        cudaFree(d_tabs);
        cudaFree(d_st);
        cudaFree(d_pcm);
        cudaFree(d_sw);
        cudaFree(cb.P);
        cudaFree(cb.E);
        cudaFree(cb.gi);
        cudaFree(cb.xr);
        cudaFree(cb.raw);
        cudaFree(cb.ms_raw);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

// number of encode calls the CLI makes for a clip (tomp3.cpp:923-931: four zero frames appended at EOF,
// whole frames only)
long long calls_for(long long nsamples) { return (nsamples + 4 * 1152) / 1152; }
const int kFlushCalls = 12;  // upper bound of the tail-flush calls we provision analysis data for

int plan_create(hmp3_batch *b, const hmp3_control *controls, const long long *nsamples, int n, int device,
                int chunk_granules) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) {
        set_err("no usable CUDA device (this library has no CPU path)");
        return HMP3_ERR_NO_DEVICE;
    }
    CK(cudaSetDevice(device));
    b->device = device;
    b->n = n;
    b->status.assign(n, HMP3_OK);
    b->st_h.resize(n);
    long long pcm_off = 0;
    int max_gran = 0;
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
                delete T;
            } else {
                cfg = (int)b->controls.size();
                b->controls.push_back(controls[i]);
                b->tabs_h.push_back(*T);
                delete T;
            }
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
        pcm_off += nsamples[i] * (sd.nch ? sd.nch : 1);
        pcm_off = (pcm_off + 7) & ~7LL;
        max_gran = std::max(max_gran, sd.ngran);
    }
    if (b->tabs_h.empty()) {
        set_err("no stream has a valid control block");
        return HMP3_ERR_BAD_CONTROL;
    }
    b->pcm_elems = pcm_off;
    b->max_gran = max_gran;
    b->NG = chunk_granules;
    CK(cudaStreamCreate(&b->stream));
    CK(cudaMalloc(&b->d_tabs, sizeof(EncTables) * b->tabs_h.size()));
    CK(cudaMemcpy(b->d_tabs, b->tabs_h.data(), sizeof(EncTables) * b->tabs_h.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&b->d_st, sizeof(StreamDev) * n));
    CK(cudaMemcpy(b->d_st, b->st_h.data(), sizeof(StreamDev) * n, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&b->d_pcm, sizeof(int16_t) * std::max<long long>(b->pcm_elems, 8)));
    CK(cudaMalloc(&b->d_sw, sizeof(SwitchState) * n));
    const long long NG = b->NG, G = NG + 3;
    b->cb.NG = (int)NG;
    CK(cudaMalloc(&b->cb.P, sizeof(float) * n * G * 2 * 576));
    CK(cudaMalloc(&b->cb.E, sizeof(int) * n * G * 2 * 9));
    CK(cudaMalloc(&b->cb.gi, sizeof(GranuleInfo) * n * NG));
    CK(cudaMalloc(&b->cb.xr, sizeof(float) * n * NG * 2 * 576));
    CK(cudaMalloc(&b->cb.raw, sizeof(PsyRaw) * n * NG * 2));
    CK(cudaMalloc(&b->cb.ms_raw, sizeof(int) * n * NG));
    return HMP3_OK;
}

int plan_reset_state(hmp3_batch *b) {
    std::vector<SwitchState> sw(b->n);
    for (auto &s : sw) switch_state_init(&s);
    CK(cudaMemcpyAsync(b->d_sw, sw.data(), sizeof(SwitchState) * b->n, cudaMemcpyHostToDevice, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return HMP3_OK;
}

// Phase A for the chunk starting at encode granule K0.
int launch_analysis(hmp3_batch *b, int K0) {
    const int n = b->n;
    const long long NG = b->NG, G = NG + 3;
    k_polyphase<<<blocks_for((long long)n * G * 2 * 18, 128), 128, 0, b->stream>>>(b->d_tabs, b->d_st, b->d_pcm, b->cb,
                                                                                     K0, n);
    k_attack<<<blocks_for((long long)n * G * 2 * 9, 128), 128, 0, b->stream>>>(b->d_tabs, b->d_st, b->cb, K0, n);
    k_switch_scan<<<blocks_for(n, 32), 32, 0, b->stream>>>(b->d_tabs, b->d_st, b->d_sw, b->cb, K0, n);
    k_hybrid<<<blocks_for((long long)n * NG * 2 * 32, 128), 128, 0, b->stream>>>(b->d_tabs, b->d_st, b->cb, K0, n);
    k_psy_stage1<<<blocks_for((long long)n * NG * 3, 64), 64, 0, b->stream>>>(b->d_tabs, b->d_st, b->cb, K0, n);
    b->launches += 5;
    CK(cudaGetLastError());
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

// Debug / parity entry: Phase A of ONE stream on the device, stage outputs copied back to the host.
// Same argument meaning as the host simulator's sim_analysis (tests/hostsim/hostsim.cpp).
int hmp3_debug_analysis(const hmp3_control *ec, const int16_t *pcm, long long nsamples, int ngran, int device,
                        float *sbt_out, int *ginfo, float *xr_out, float *raw_out, int *ms_raw, int *att) {
    hmp3_batch b;
    int r = plan_create(&b, ec, &nsamples, 1, device, 32);
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
        r = launch_analysis(&b, K0);
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
