// Phase A CUDA kernels (sm_100a): one thread (or warp) per independent work item of analysis.h.
// All are launched over a *chunk* of encode granules [K0, K0+NG) of every stream of the batch.
#pragma once
#include <cuda_runtime.h>
#include "analysis.h"
#include "prepare.h"
#include "batch_types.h"

namespace hmp3 {

// ---- K1: polyphase analysis, one thread per (stream, polyphase granule, channel, time slot)
__global__ void __launch_bounds__(128) k_polyphase(const EncTables *tabs, const StreamDev *st, const int16_t *pcm,
                                                   ChunkBufs cb, int K0, int nstreams) {
    const int G = cb.NG + 3;
    long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)nstreams * G * 2 * 18;
    if (id >= total) return;
    int t = (int)(id % 18);
    long long r = id / 18;
    int ch = (int)(r & 1);
    r >>= 1;
    int jj = (int)(r % G);
    int s = (int)(r / G);
    const StreamDev sd = st[s];
    if (ch >= sd.nch) return;
    long long j = (long long)K0 - 3 + jj;
    if (j >= sd.ngran) return;
    const EncTables *T = tabs + sd.cfg;
    float *out = cb.P + (((long long)s * G + jj) * 2 + ch) * 576;
    polyphase_item(T, pcm + sd.pcm_off, (long)sd.nsamples, sd.nch, ch, (long)j, t, out);
}

// ---- K2: attack energies, one thread per (stream, polyphase granule, channel, slot pair)
__global__ void k_attack(const EncTables *tabs, const StreamDev *st, ChunkBufs cb, int K0, int nstreams) {
    const int G = cb.NG + 3;
    long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)nstreams * G * 2 * 9;
    if (id >= total) return;
    int k = (int)(id % 9);
    long long r = id / 9;
    int ch = (int)(r & 1);
    r >>= 1;
    int jj = (int)(r % G);
    int s = (int)(r / G);
    const StreamDev sd = st[s];
    if (ch >= sd.nch) return;
    long long j = (long long)K0 - 3 + jj;
    if (j >= sd.ngran) return;
    const EncTables *T = tabs + sd.cfg;
    const float *p = cb.P + (((long long)s * G + jj) * 2 + ch) * 576;
    cb.E[(((long long)s * G + jj) * 2 + ch) * 9 + k] = attack_energy(T, p, k, T->cfg.h_id == 0);
}

// ---- K3: block-type decision scan, one thread per stream, sequential over the chunk
__global__ void k_switch_scan(const EncTables *tabs, const StreamDev *st, SwitchState *sw, ChunkBufs cb, int K0,
                              int nstreams) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nstreams) return;
    const StreamDev sd = st[s];
    const EncTables *T = tabs + sd.cfg;
    const int G = cb.NG + 3;
    SwitchState state = sw[s];
    for (int q = 0; q < cb.NG; q++) {
        int K = K0 + q;
        if (K >= sd.ngran) break;
        int jj = q + 2;  // P[K-1]
        const int *e0 = cb.E + (((long long)s * G + jj) * 2 + 0) * 9;
        const int *e1 = cb.E + (((long long)s * G + jj) * 2 + (sd.nch - 1)) * 9;
        cb.gi[(long long)s * cb.NG + q] = switch_step(T, &state, e0, e1);
    }
    sw[s] = state;
}

// ---- K4: hybrid window + MDCT + alias reduction, one warp per (stream, granule, channel), lane = sub-band
__global__ void __launch_bounds__(128) k_hybrid(const EncTables *tabs, const StreamDev *st, ChunkBufs cb, int K0,
                                                int nstreams) {
    const int G = cb.NG + 3;
    long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    long long total = (long long)nstreams * cb.NG * 2;
    if (wid >= total) return;
    int ch = (int)(wid & 1);
    long long r = wid >> 1;
    int q = (int)(r % cb.NG);
    int s = (int)(r / cb.NG);
    const StreamDev sd = st[s];
    if (ch >= sd.nch || K0 + q >= sd.ngran) return;
    const EncTables *T = tabs + sd.cfg;
    const int bt = cb.gi[(long long)s * cb.NG + q].block_type;
    const float *prev = cb.P + (((long long)s * G + q) * 2 + ch) * 576;      // P[K-3]
    const float *cur = cb.P + (((long long)s * G + q + 1) * 2 + ch) * 576;   // P[K-2]
    float *xr = cb.xr + (((long long)s * cb.NG + q) * 2 + ch) * 576;
    hybrid_item(T, prev, cur, bt, lane, xr);
    __syncwarp();
    if (bt != 2) alias_item(T, lane, xr);
}

// ---- K5: psychoacoustic stage 1 (one thread per granule-channel) and M/S measure (one per granule)
__global__ void __launch_bounds__(64) k_psy_stage1(const EncTables *tabs, const StreamDev *st, ChunkBufs cb, int K0,
                                                   int nstreams) {
    long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)nstreams * cb.NG * 3;
    if (id >= total) return;
    int job = (int)(id % 3);
    long long r = id / 3;
    int q = (int)(r % cb.NG);
    int s = (int)(r / cb.NG);
    const StreamDev sd = st[s];
    if (K0 + q >= sd.ngran) return;
    const EncTables *T = tabs + sd.cfg;
    const int bt = cb.gi[(long long)s * cb.NG + q].block_type;
    const float *x0 = cb.xr + (((long long)s * cb.NG + q) * 2) * 576;
    if (job < 2) {
        if (job >= sd.nch) return;
        PsyRaw *R = cb.raw + ((long long)s * cb.NG + q) * 2 + job;
        if (bt != 2) psy_long_stage1(T, x0 + 576 * job, R);
        else psy_short_stage1(T, x0 + 576 * job, R);
    } else {
        int m = 0;
        if (sd.nch == 2) m = (bt != 2) ? ms_measure_long(T, x0, x0 + 576) : ms_measure_short(T, x0, x0 + 576);
        cb.ms_raw[(long long)s * cb.NG + q] = m;
    }
}

// ---- K5b: M/S decision scan (hysteresis memory), one thread per stream, sequential over the chunk.
// Frames are pairs of granules (MPEG-1: the decision is per frame, mp3enc.cpp:1537-1546; MPEG-2: per granule).
__global__ void k_ms_scan(const EncTables *tabs, const StreamDev *st, int *msmem, ChunkBufs cb, int K0, int nstreams) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nstreams) return;
    const StreamDev sd = st[s];
    const EncTables *T = tabs + sd.cfg;
    const bool stereo_ms = (sd.nch == 2 && T->cfg.ms_flag);
    const bool m1 = T->cfg.h_id == 1;
    int mem = msmem[s];
    const GranuleInfo *gi = cb.gi + (long long)s * cb.NG;
    const int *raw = cb.ms_raw + (long long)s * cb.NG;
    signed char *out = cb.ms + (long long)s * cb.NG;
    for (int q = 0; q + 1 < cb.NG && K0 + q + 1 < sd.ngran; q += 2) {
        int f0 = 0, f1 = 0;
        if (stereo_ms) {
            const int a = ms_scan_step(&mem, gi[q].block_type, raw[q]);
            if (m1) {
                const int b = ms_scan_step(&mem, gi[q + 1].block_type, raw[q + 1]);
                f0 = f1 = ((a + b) >= 0);
            } else {
                f0 = (a >= 0);
                const int b = ms_scan_step(&mem, gi[q + 1].block_type, raw[q + 1]);
                f1 = (b >= 0);
            }
        }
        out[q] = (signed char)f0;
        out[q + 1] = (signed char)f1;
    }
    msmem[s] = mem;
}

// ---- K5c: psychoacoustic stage 2 (pre-echo memory), one thread per (stream, channel), sequential over the chunk
__global__ void k_psy_stage2(const EncTables *tabs, const StreamDev *st, PsyState *psy, ChunkBufs cb, int K0,
                             int nstreams) {
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = id >> 1, ch = id & 1;
    if (s >= nstreams) return;
    const StreamDev sd = st[s];
    if (ch >= sd.nch) return;
    const EncTables *T = tabs + sd.cfg;
    PsyState *P = psy + (long long)s * 2 + ch;
    for (int q = 0; q < cb.NG && K0 + q < sd.ngran; q++) {
        const GranuleInfo g = cb.gi[(long long)s * cb.NG + q];
        const PsyRaw *R = cb.raw + ((long long)s * cb.NG + q) * 2 + ch;
        if (g.block_type != 2) psy_long_stage2(T, R, P->echo, g.block_type, P->sm);
        else psy_short_stage2(T, R, P->echo, g.block_type_prev, P->sm);
        SigMask *o = cb.sm + (((long long)s * cb.NG + q) * 2 + ch) * 36;
        for (int i = 0; i < 36; i++) o[i] = P->sm[i];
    }
}

// ---- K5d: prepare pass, one warp per (stream, granule): state-free part of the rate-loop prologue (long blocks)
__global__ void __launch_bounds__(128) k_prepare(const EncTables *tabs, const StreamDev *st, ChunkBufs cb, int K0,
                                                 int nstreams) {
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int q = (int)(wid % cb.NG), s = (int)(wid / cb.NG);
    if (s >= nstreams) return;
    const StreamDev sd = st[s];
    if (K0 + q >= sd.ngran) return;
    const long long o = (long long)s * cb.NG + q;
    if (cb.gi[o].block_type == 2) return;
    const EncTables *T = tabs + sd.cfg;
    // the flag the allocator is called with (mp3enc.cpp:1556 / :1880: MPEG-2 mono passes the configured ms_flag)
    const int ms = (T->cfg.h_id == 0 && sd.nch != 2) ? T->cfg.ms_flag : (int)cb.ms[o];
    long_prepare(T, ms, cb.xr + o * 2 * 576, cb.prep + o);
}

__global__ void k_prepare_init(int *msmem, PsyState *psy, int nstreams) {
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= 2 * nstreams) return;
    psy_state_init(psy + id);
    if ((id & 1) == 0) msmem[id >> 1] = 0;
}

}  // namespace hmp3
