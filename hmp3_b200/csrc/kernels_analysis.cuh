// Phase A CUDA kernels (sm_100a): one thread (or warp) per independent work item of analysis.h.
// All are launched over a *chunk* of encode granules [K0, K0+NG) of every stream of the batch.
#pragma once
#include <cuda_runtime.h>
#include "analysis.h"
#include "prepare.h"
#include "batch_types.h"

namespace hmp3 {

// Bulk asynchronous global -> shared copies by the TMA engine (cp.async.bulk, 1-D form: a granule row is 2304
// contiguous bytes).  The staging loops of these kernels used to load a value and store it to shared memory at
// once: one line in flight per warp, and 55-75 % of the kernel's stall samples on that store (profiles/r2z).  Now
// one lane of the warp hands the whole tile to the copy engine (one or two instructions, no registers), the bytes land
// in shared memory past L1, and the warp waits on its own mbarrier for the transaction count.
struct WarpTma {
    unsigned bar;  // shared-memory address of the warp's mbarrier
};
__device__ __forceinline__ unsigned smem_u32_(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ WarpTma tma_begin(unsigned long long *bar, int lane, unsigned bytes) {
    WarpTma t;
    t.bar = smem_u32_(bar);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(t.bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(t.bar), "r"(bytes) : "memory");
    }
    __syncwarp();
    return t;
}
// bytes: a multiple of 16; both addresses 16-byte aligned.  Issued by lane 0 only.
__device__ __forceinline__ void tma_row(const WarpTma &t, void *smem, const void *gmem, unsigned bytes, int lane) {
    if (lane == 0)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32_(smem)),
                     "l"(gmem), "r"(bytes), "r"(t.bar)
                     : "memory");
}
__device__ __forceinline__ void tma_wait(const WarpTma &t) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@!p bra W_%=;\n\t}" ::"r"(t.bar)
        : "memory");
}

// The polyphase tables do not depend on the configuration: they live in constant memory, and with the window fold
// fully unrolled every coefficient is an instruction operand (no load) and every sample a shared-memory load at an
// immediate offset.  Filled once per device by launch_polyphase (fixed_polyphase_tables, enc_init.cpp).
__constant__ float c_polyA[32][8];
__constant__ float c_polyB[32][8];
__constant__ float c_dct32[31];
// kPolyIa / kPolyIb of tables_data.h as functions of k (checked against the tables when the constants are filled)
__host__ __device__ constexpr int poly_ia(int k) { return k <= 16 ? 16 + k : 80 - k; }
__host__ __device__ constexpr int poly_ib(int k) { return (k == 0 || k == 16) ? 0 : (k < 16 ? 16 - k : 16 + k); }
// polyphase_slot (dsp_core.h) for the staged PCM row: rb = the row at this slot's origin (33 * slot with the row
// padding), sample i (0 = newest) at rb[(511 - i) + ((511 - i) >> 5)].  Same sums, same order.
__device__ __forceinline__ void polyphase_slot_c(const float *rb, float *out, int stride) {
    float a[32], b[32];
#define HMP3_PCM(i) rb[(511 - (i)) + ((511 - (i)) >> 5)]
    {
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; j++) s += c_polyA[0][j] * HMP3_PCM(poly_ia(0) + 64 * j);
        b[0] = s;
    }
#pragma unroll
    for (int k = 1; k < 32; k++) {
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            s1 += c_polyA[k][j] * HMP3_PCM(poly_ia(k) + 64 * j);
            s2 += c_polyB[k][j] * HMP3_PCM(poly_ib(k) + 64 * j);
        }
        b[k] = s1 + s2;
    }
#undef HMP3_PCM
    const float *c = c_dct32;
    dct32_split<32, 1>(b, a);
    dct32_split<16, 2>(a, b);
    dct32_split<8, 4>(b, a);
    dct32_split<4, 8>(a, b);
    dct32_merge<2, 16>(b, a, c + 16 + 8 + 4 + 2);
    dct32_merge<4, 8>(a, b, c + 16 + 8 + 4);
    dct32_merge<8, 4>(b, a, c + 16 + 8);
    dct32_merge<16, 2>(a, b, c + 16);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const float tw = c[k] * b[k + 16];
        out[stride * k] = b[k] + tw;
        out[stride * (31 - k)] = b[k] - tw;
    }
}

// ---- K1: polyphase analysis.  One block = kPolyRun consecutive polyphase granules of one stream, both channels:
// the PCM span they need (576 * run + 480 samples per channel) is staged once in shared memory as float
// (coalesced 32-bit reads of the interleaved int16 input, zero outside the clip), then one thread per
// (channel, time slot) folds its 512-sample window and runs the 32-point fast DCT in the reference's operation
// order.  Shared-memory rows are padded (index + index / 32) so that the slots of a warp, whose windows start
// 32 samples apart, hit 32 different banks.
constexpr int kPolyRun = 7;                                 // granules per block: 126 slots per channel
constexpr int kPolySpan = 576 * kPolyRun + 480;             // samples per channel staged
constexpr int kPolyRow = kPolySpan + (kPolySpan >> 5) + 1;  // padded row length
// ---- K0: optional DC-blocking input filter (filter2.c:112-147, -S1): t = x - d; d += alpha * t over the whole
// stream including the zero tail; one thread per (stream, channel), sequential (a 1-pole recursion), carried state d.
__global__ void k_dc_filter(const EncTables *tabs, const StreamDev *st, const int16_t *pcm, float *pcmf, float *dc,
                            long long lo, long long hi, int nstreams) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = id >> 1, ch = id & 1;
    if (s >= nstreams) return;
    const StreamDev sd = st[s];
    if (ch >= sd.nch || sd.pcmf_off < 0) return;
    const float alpha = tabs[sd.cfg].cfg.dc_alpha;
    const int16_t *src = pcm + sd.pcm_off;
    const float *fsrc = pcmf + (sd.rawf_off >= 0 ? sd.rawf_off : 0);
    const bool fin = sd.rawf_off >= 0;
    float *dst = pcmf + sd.pcmf_off;
    float d = dc[2 * s + ch];
    const long long e = hi < sd.pcmf_len ? hi : sd.pcmf_len;
    for (long long n = lo; n < e; n++) {
        const float x = n < sd.nsamples ? (fin ? fsrc[n * sd.nch + ch] : (float)src[n * sd.nch + ch]) : sd.tail;
        const float t = (x - d);
        d = d + alpha * t;
        dst[n * sd.nch + ch] = t;
    }
    dc[2 * s + ch] = d;
}

__global__ void __launch_bounds__(256) k_polyphase(const EncTables *tabs, const StreamDev *st, const int16_t *pcm,
                                                   const float *pcmf, ChunkBufs cb, int K0, int nstreams) {
    __shared__ float s_pcm[2][kPolyRow];
    const int G = cb.NG + 3;
    const int s = blockIdx.y;
    if (s >= nstreams) return;
    const StreamDev sd = st[s];
    const int jj0 = blockIdx.x * kPolyRun;          // first polyphase granule of this block within the chunk
    const long long j0 = (long long)K0 - 3 + jj0;   // ... and its absolute index (may be negative: zero history)
    if (jj0 >= G || j0 >= sd.ngran) return;
    const EncTables *T = tabs + sd.cfg;
    const int nch = sd.nch;
    // stage samples n0 .. n0 + kPolySpan - 1 of every channel
    const long long n0 = 576 * j0 - 480;
    const int16_t *src = pcm + sd.pcm_off;
    if (sd.pcmf_off >= 0 || sd.rawf_off >= 0) {  // float samples: the DC-filtered copy, or float input as is
        const float *fsrc = pcmf + (sd.pcmf_off >= 0 ? sd.pcmf_off : sd.rawf_off);
        const long long flen = sd.pcmf_off >= 0 ? sd.pcmf_len : sd.nsamples;
        for (int p = threadIdx.x; p < kPolySpan; p += 256) {
            const long long n = n0 + p;
            const bool in = (n >= 0 && n < flen);
            const float out = n < 0 ? 0.0f : sd.tail;  // before the stream: silence; after it: the flush value
            const int q = p + (p >> 5);
            s_pcm[0][q] = in ? fsrc[n * nch] : out;
            if (nch == 2) s_pcm[1][q] = in ? fsrc[n * nch + 1] : out;
        }
    } else if (nch == 2) {
        const unsigned *src2 = (const unsigned *)src;  // pcm_off is even-aligned: one 32-bit word = (left, right)
        for (int p0 = threadIdx.x; p0 < kPolySpan; p0 += 6 * 256) {  // six loads in flight, then the six stores
            unsigned w[6];
#pragma unroll
            for (int u = 0; u < 6; u++) {
                const int p = p0 + 256 * u;
                const long long n = n0 + p;
                w[u] = (p < kPolySpan && n >= 0 && n < sd.nsamples) ? src2[n] : 0u;
            }
#pragma unroll
            for (int u = 0; u < 6; u++) {
                const int p = p0 + 256 * u;
                if (p < kPolySpan) {
                    const int q = p + (p >> 5);
                    s_pcm[0][q] = (float)(short)(w[u] & 0xffffu);
                    s_pcm[1][q] = (float)(short)(w[u] >> 16);
                }
            }
        }
    } else {
        for (int p = threadIdx.x; p < kPolySpan; p += 256) {
            const long long n = n0 + p;
            s_pcm[0][p + (p >> 5)] = (n >= 0 && n < sd.nsamples) ? (float)src[n] : 0.0f;
        }
    }
    __syncthreads();
    const int ch = threadIdx.x >> 7, slot = threadIdx.x & 127;  // slot = 18 * (granule within the run) + t
    if (ch >= nch || slot >= 18 * kPolyRun) return;
    const int jr = slot / 18, t = slot - 18 * jr;
    const int jj = jj0 + jr;
    if (jj >= G || j0 + jr >= sd.ngran) return;
    // this slot's newest sample sits at position p = 32 * slot + 511 of the staged span, i.e. at padded index
    // p + (p >> 5) = 33 * slot + 511 + 15
    float col[32];
    polyphase_slot_c(s_pcm[ch] + 33 * slot, col, 1);
    float *out = cb.P + (((long long)s * G + jj) * 2 + ch) * 576;
    const int nsb = T->cfg.nsb_hybrid;
#pragma unroll
    for (int sb = 0; sb < 32; sb++) out[18 * sb + t] = freq_inverted(sb, t, nsb) ? -col[sb] : col[sb];
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

// ---- K4: hybrid window + MDCT + alias reduction, one warp per (stream, granule, channel), lane = sub-band.
// The two polyphase granules are brought into shared memory with coalesced loads, the transform and the alias
// butterflies run there, and the 576 lines go back with coalesced stores (the per-lane access pattern -- 18
// consecutive values per sub-band -- would otherwise touch 32 different cache lines per instruction).
__global__ void __launch_bounds__(128) k_hybrid(const EncTables *tabs, const StreamDev *st, ChunkBufs cb, int K0,
                                                int nstreams) {
    __shared__ __align__(16) float s_buf[4][3][576];
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
    float(*sb)[576] = s_buf[threadIdx.x >> 5];
    {
        __shared__ unsigned long long s_bar[4];
        const WarpTma t = tma_begin(&s_bar[threadIdx.x >> 5], lane, 2 * 2304);
        tma_row(t, sb[0], prev, 2304, lane);
        tma_row(t, sb[1], cur, 2304, lane);
        tma_wait(t);
    }
    hybrid_item(T, sb[0], sb[1], bt, lane, sb[2]);
    __syncwarp();
    if (bt != 2) alias_item(T, lane, sb[2]);
    __syncwarp();
    for (int k = lane; k < 576; k += 32) xr[k] = sb[2][k];
}

// ---- K5: psychoacoustic stage 1 and the M/S measure, one warp per (stream, granule).  The granule's spectra are
// staged in shared memory; long blocks run one partition (or scale-factor band) per lane -- each lane does the
// reference's sequential sums for its partition, the integer statistics are warp reductions -- and the rare
// short blocks run the plain sequential routine on one lane.
__device__ __forceinline__ void psy_long_stage1_warp(const EncTables *T, const float *xr, PsyRaw *R, float *xtab,
                                                     int *mbetab, int *snrv, int lane) {
    const int nmap = T->psy_emap_n_l;
    const int npart = T->psy_npart_l;
    const int npart2 = (npart + 1) & ~1;
    const float *w = T->w_spd_l;
    for (int j = lane; j < 44; j += 32) {  // partition energies (emap.c:96-121)
        float s = 0.0f;
        if (j < nmap) {
            const int n = T->psy_nsum_l[j];
            const float *x = xr + T->psy_start_l[j];
            for (int q = 0; q < n; q++) s += x[q] * x[q];
        }
        float e = 0.0f, xt = 0.0f;
        int mbe = 0;
        if (j < npart2) {
            e = w[j] + s;
            mbe = mb_log(T, e);
            xt = mb_exp(T, (int)(0.30f * mbe));
        }
        R->e[j] = e;
        mbetab[j] = mbe;
        xtab[j] = xt;
    }
    __syncwarp();
    int nsnr = 0, totsnr = 0;
    float stab[2] = {0.0f, 0.0f};
    for (int r = 0; r < 2; r++) {  // spreading (spdsmr.c:188-277)
        const int i = lane + 32 * r;
        if (i < npart) {
            const int p = T->spd_off_l[i], n = T->spd_cnt_l[i], k = T->spd_w0_l[i];
            float sp = 0.1f;
            for (int j = 0; j < n; j++) sp += w[k + j] * xtab[p + j];
            sp = (0.03f * 0.1f * 0.35f) * mb_exp(T, (int)((1.0f / 0.30f) * mb_log(T, sp))) + w[i];
            stab[r] = sp;
            const int snr = mbetab[i] - mb_log(T, w[i] + sp);
            if (snr > 0) nsnr++;
            totsnr += (snr > -200 ? snr : -200);
            snrv[i] = snr;
        }
    }
    __syncwarp();
    int snrvar = 0;
    for (int r = 0; r < 2; r++) {
        const int i = lane + 32 * r;
        if (i < npart) snrvar += iabs(snrv[i] - (i > 0 ? snrv[i - 1] : 0));
    }
    nsnr = __reduce_add_sync(0xffffffffu, nsnr);
    totsnr = __reduce_add_sync(0xffffffffu, totsnr);
    snrvar = __reduce_add_sync(0xffffffffu, snrvar);
    int d = 0;
    if (nsnr > 0) {
        int d0 = round_away(1.3f * (totsnr / npart) - 850);
        int itmp = snrvar / npart;
        int dv = (500 - itmp) < 0 ? (500 - itmp) : 0;
        d = d0 + dv;
        d = d > -2000 ? d : -2000;
        d = d < 600 ? d : 600;
    }
    d += 300;
    const int dm0 = (300 - d) >> 4;
    for (int r = 0; r < 2; r++) {
        const int i = lane + 32 * r;
        if (i < npart2) {
            const int m = i >> 1;
            const int t13 = (m - 13) > 0 ? (m - 13) : 0;
            const int dm = dm0 * t13 > 0 ? dm0 * t13 : 0;
            const float a = mb_exp(T, d + dm);
            R->thr[i] = a * stab[r];
        }
    }
}

__device__ __forceinline__ int ms_measure_long_warp(const EncTables *T, const float *x0, const float *x1, int lane) {
    int cm = 0;
    const int nsf = T->cfg.nsf[0];
    if (lane < nsf) {  // one scale-factor band per lane (bitallo3.cpp:698-744)
        const int i = lane, n = T->nBand_l[i], k0 = T->startBand_l[i];
        float el = 100.0f, er = 100.0f, t = 0.0f;
        for (int k = k0; k < k0 + n; k++) {
            float a = x0[k] * x0[k];
            float b = x1[k] * x1[k];
            float c = x0[k] * x1[k];
            el += a;
            er += b;
            t += c;
        }
        float es = el + er, ed = es;
        t = t + t;
        es = es + t;
        ed = ed - t;
        int mblr = mb_log(T, el + er) - mb_log(T, el > er ? el : er);
        int mbsd = mb_log(T, es + ed) - mb_log(T, es > ed ? es : ed);
        int q = 75 - iabs(mblr - 120);
        int psd = q > 0 ? q : 0;
        int h = (mbsd >> 1) + 120;
        mbsd = mbsd < h ? mbsd : h;
        mbsd += psd;
        cm = n * (mblr - mbsd);
    }
    return __reduce_add_sync(0xffffffffu, cm);
}

// Short blocks, lane-parallel (same per-item operations and orders as psy_short_stage1 / ms_measure_short of
// psy_core.h): the partition energies are 3 x 32 independent ordered sums, the spreading 3 x npart/2 independent pairs
// of ordered sums, the M/S measure 3 x nsf independent bands with integer votes.  (On one lane a short granule cost
// several long ones and held its whole block back.)
__device__ __forceinline__ void psy_short_stage1_warp(const EncTables *T, const float *xr /*[3][192]*/, PsyRaw *R,
                                                      float *e /*[96] scratch*/, int lane) {
    const int nmap = T->psy_emap_n_s;
    const int npart = T->psy_npart_s;
    const float *w = T->w_spd_s;
    for (int it = lane; it < 96; it += 32) {
        const int win = it >> 5, j = it & 31;
        float s = 0.0f;
        if (j < nmap) {
            const int n = T->psy_nsum_s[j];
            const float *x = xr + 192 * win + T->psy_start_s[j];
            for (int q = 0; q < n; q++) s += x[q] * x[q];
        }
        e[it] = s;
    }
    __syncwarp();
    const int mpart = (npart + 1) >> 1;
    for (int it = lane; it < 3 * mpart; it += 32) {
        const int win = it / mpart, m = it - win * mpart, i = 2 * m;
        int k = 0;  // first weight of partition i: the weights of all partitions follow each other
        for (int q = 0; q < i; q++) k += T->spd_cnt_s[q];
        const float *ew = e + 32 * win;
        int p = T->spd_off_s[i], n = T->spd_cnt_s[i];
        float s0 = 0.5f;
        for (int j = 0; j < n; j++, k++) s0 += w[k] * ew[p + j];
        p = T->spd_off_s[i + 1];
        n = T->spd_cnt_s[i + 1];
        float t0 = 0.5f;
        for (int j = 0; j < n; j++, k++) t0 += w[k] * ew[p + j];
        R->thr[16 * win + m] = s0 + t0;
    }
    __syncwarp();
}
__device__ __forceinline__ int ms_measure_short_warp(const EncTables *T, const float *x0, const float *x1, int lane) {
    const int nsf = T->cfg.nsf_s[0];
    int d = 0;
    for (int it = lane; it < 3 * nsf; it += 32) {
        const int win = it / nsf, i = it - win * nsf;
        const int k0 = 192 * win + T->startBand_s[i], n = T->nBand_s[i];
        float s0 = 0.0f, s1 = 0.0f;
        for (int k = k0; k < k0 + n; k++) {
            float a = x0[k] * x0[k];
            const float b = x1[k] * x1[k];
            s0 += (a + b);
            a = a - b;
            if (a < 0.0f) a = -a;
            s1 += a;
        }
        if ((double)s1 > 0.80 * (double)s0) d++;
        if ((double)s1 > 0.95 * (double)s0) d += 2;
    }
    d = __reduce_add_sync(0xffffffffu, d);
    return (nsf - d) << 10;
}

__global__ void __launch_bounds__(128) k_psy_stage1(const EncTables *tabs, const StreamDev *st, ChunkBufs cb, int K0,
                                                    int nstreams) {
    __shared__ __align__(16) float s_x[4][2][576];
    __shared__ float s_xtab[4][44];
    __shared__ int s_mbe[4][44], s_snr[4][44];
    __shared__ float s_e[4][96];
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int q = (int)(wid % cb.NG), s = (int)(wid / cb.NG);
    if (s >= nstreams) return;
    const StreamDev sd = st[s];
    if (K0 + q >= sd.ngran) return;
    const EncTables *T = tabs + sd.cfg;
    const long long o = (long long)s * cb.NG + q;
    const int bt = cb.gi[o].block_type;
    const float *x0 = cb.xr + o * 2 * 576;
    {
        __shared__ unsigned long long s_bar[4];
        const WarpTma t = tma_begin(&s_bar[wl], lane, 2304u * sd.nch);
        tma_row(t, s_x[wl][0], x0, 2304u * sd.nch, lane);  // the granule's rows lie one after the other
        tma_wait(t);
    }
    for (int c = 0; c < sd.nch; c++) {
        PsyRaw *R = cb.raw + o * 2 + c;
        if (bt != 2) psy_long_stage1_warp(T, s_x[wl][c], R, s_xtab[wl], s_mbe[wl], s_snr[wl], lane);
        else psy_short_stage1_warp(T, s_x[wl][c], R, s_e[wl], lane);
        __syncwarp();
    }
    int m = 0;
    if (sd.nch == 2) {
        if (bt != 2) m = ms_measure_long_warp(T, s_x[wl][0], s_x[wl][1], lane);
        else m = ms_measure_short_warp(T, s_x[wl][0], s_x[wl][1], lane);
    }
    if (lane == 0) cb.ms_raw[o] = m;
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

// ---- K5c: psychoacoustic stage 2 (pre-echo memory), one warp per (stream, channel), sequential over the chunk,
// one partition pair (long) / one partition (short) per lane; the carried state lives in shared memory while the
// chunk is scanned.  Same per-index operations as psy_long_stage2 / psy_short_stage2 (psy_core.h).
__global__ void __launch_bounds__(128) k_psy_stage2(const EncTables *tabs, const StreamDev *st, PsyState *psy, ChunkBufs cb,
                                                    int K0, int nstreams) {
    __shared__ float s_echo[4][64];
    __shared__ SigMask s_sm[4][36];
    const int id = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int s = id >> 1, ch = id & 1;
    if (s >= nstreams) return;
    const StreamDev sd = st[s];
    if (ch >= sd.nch) return;
    const EncTables *T = tabs + sd.cfg;
    PsyState *P = psy + (long long)s * 2 + ch;
    float *echo = s_echo[wl];
    SigMask *sm = s_sm[wl];
    for (int i = lane; i < 64; i += 32) echo[i] = P->echo[i];
    for (int i = lane; i < 36; i += 32) sm[i] = P->sm[i];
    __syncwarp();
    const int npl = T->psy_npart_l, mps = (T->psy_npart_s + 1) >> 1;
    for (int q = 0; q < cb.NG && K0 + q < sd.ngran; q++) {
        const GranuleInfo g = cb.gi[(long long)s * cb.NG + q];
        const PsyRaw *R = cb.raw + ((long long)s * cb.NG + q) * 2 + ch;
        if (g.block_type != 2) {
            const int i = 2 * lane;
            if (i < npl) {  // spdsmr.c:279-316
                float s1 = R->thr[i];
                float t = echo[i];
                echo[i] = (float)(2.0 * s1);
                if (g.block_type != 3) {
                    if (s1 > t) {
                        float x = 0.1f * s1;
                        s1 = t;
                        if (s1 < x) s1 = x;
                    }
                }
                float s2 = R->thr[i + 1];
                t = echo[i + 1];
                echo[i + 1] = 2.0f * s2;
                if (g.block_type != 3) {
                    if (s2 > t) {
                        float x = 0.1f * s2;
                        s2 = t;
                        if (s2 < x) s2 = x;
                    }
                }
                float e0 = R->e[i], e1 = R->e[i + 1];
                float emax = e0;
                if (emax < e1) emax = e1;
                sm[lane].sig = e0 + e1;
                sm[lane].mask = (e0 * s1 + e1 * s2) / emax;
            }
        } else {
            const int i = lane;
            if (i < mps) {  // spdsmr.c:110-181
                float k0 = R->thr[i], k1 = R->thr[16 + i], k2 = R->thr[32 + i];
                float m0 = echo[i];
                float m1 = (float)(2.0 * k0);
                float m2 = (float)(2.0 * k1);
                echo[i] = (float)(2.0 * k2);
                if (g.block_type_prev == 2) {
                    float t = k0;
                    if (t > m0) {
                        float tmp = 0.1f * t;
                        k0 = (m0 > tmp) ? m0 : tmp;
                    }
                }
                {
                    float t = k1;
                    if (t > m1) {
                        float tmp = 0.1f * t;
                        k1 = (m1 > tmp) ? m1 : tmp;
                    }
                }
                {
                    float t = k2;
                    if (t > m2) {
                        float tmp = 0.1f * t;
                        k2 = (m2 > tmp) ? m2 : tmp;
                    }
                }
                sm[i].mask = k0;
                sm[12 + i].mask = k1 + 0.1f * k0;
                sm[24 + i].mask = k2 + 0.1f * k1;
                sm[i].sig = 0.0f;
                sm[12 + i].sig = 0.0f;
                sm[24 + i].sig = 0.0f;
            }
        }
        __syncwarp();
        float *o = (float *)(cb.sm + (((long long)s * cb.NG + q) * 2 + ch) * 36);
        const float *src = (const float *)sm;
        for (int k = lane; k < 72; k += 32) o[k] = src[k];
        __syncwarp();
    }
    for (int i = lane; i < 64; i += 32) P->echo[i] = echo[i];
    for (int i = lane; i < 36; i += 32) P->sm[i] = sm[i];
}

// ---- K5d: prepare pass, one warp per (stream, granule): state-free part of the rate-loop prologue (long blocks)
// long_prepare (prepare.h) with the granule staged in shared memory (TMA bulk copy): x = the two spectra (rewritten in place to
// magnitudes / mid-side like the global copy), q = per-line squares and then |x|^(3/4).  The ordered band sums and band
// maxima -- dependent chains as long as a band, one band per lane -- read shared memory instead of global memory, and
// the final contents of xr and PrepGranule are exactly long_prepare's (including the squares it leaves in x34 above the
// last line it raises to the 3/4).
__device__ __forceinline__ void band_sums2_sm(const EncTables *T, const float *v0, const float *v1, int nbands, float *out0,
                                              float *out1, int lane) {
    for (int i = lane; i < nbands; i += 32) {
        const int k0 = T->startBand_l[i], n = T->nBand_l[i];
        float a = 0.0f, b = 0.0f;
        for (int k = k0; k < k0 + n; k++) {
            a += v0[k];
            b += v1[k];
        }
        out0[i] = a;
        if (out1) out1[i] = b;
    }
}
__device__ __forceinline__ void band_bounds_sm(const EncTables *T, PrepGranule *P, const float *y, int ch, int nbands, int lane) {
    for (int i = lane; i < nbands; i += 32) {
        const int k0 = T->startBand_l[i], n = T->nBand_l[i];
        float m = 0.0f;
        for (int k = k0; k < k0 + n; k++)
            if (y[k] > m) m = y[k];
        P->x34max[ch][i] = m;
        const float t = (0.017716950f * mb_log(T, m) + (104.585000f - 100.0f + 8.0f));
        int g0 = round_away(t);
        if (g0 < 0) g0 = 0;
        P->gzero[ch][i] = g0;
        P->gmin[ch][i] = (g0 - kGminOffset) > 0 ? (g0 - kGminOffset) : 0;
    }
}
__device__ __forceinline__ void long_prepare_warp(const EncTables *T, int ms, float *xr, PrepGranule *P, float (*x)[576],
                                                  float (*q)[576], unsigned long long *bar, int lane) {
    const int nch = T->cfg.nchan;
    const WarpTma t = tma_begin(bar, lane, 2 * 2304);
    tma_row(t, x[0], xr, 2 * 2304, lane);  // both rows (the buffer always has two, one after the other)
    for (int w = lane; w < 36; w += 32) (&P->sign[0][0])[w] = 0;
    tma_wait(t);
    if (!ms) {
        for (int ch = 0; ch < nch; ch++) {
            const int nb = T->cfg.nsf3[ch], nl = T->startBand_l[nb];
            if (lane == 0) P->nlines[ch] = nl;
            for (int w = 0; 32 * w < nl; w++) {
                const int k = 32 * w + lane;
                int sgn = 0;
                if (k < nl) {
                    float v = x[ch][k];
                    if (!(v >= 0.0f)) {
                        sgn = 1;
                        v = -v;
                        x[ch][k] = v;
                        xr[576 * ch + k] = v;
                    }
                    q[ch][k] = v * v;
                }
                const unsigned bits = __ballot_sync(0xffffffffu, sgn);
                if (lane == 0) P->sign[ch][w] = bits;
            }
            __syncwarp();
            band_sums2_sm(T, q[ch], q[ch], nb, P->xsxx[ch], nullptr, lane);
            __syncwarp();
            const int nmax = T->cfg.nbmax3[ch];
            for (int k = lane; k < (nl > nmax ? nl : nmax); k += 32) {
                const float v = k < nmax ? pow34(T, x[ch][k]) : q[ch][k];
                q[ch][k] = v;
                P->x34[ch][k] = v;
            }
            __syncwarp();
            band_bounds_sm(T, P, q[ch], ch, nb, lane);
        }
        return;
    }
    const int nsf0 = T->cfg.nsf[0];
    const int nl = T->startBand_l[nsf0];
    const int nrot = nl + (T->cfg.hf_flag ? T->nBand_l[21] : 0);
    if (lane == 0) P->nlines[0] = P->nlines[1] = nrot;
    for (int k = lane; k < nl; k += 32) {
        q[0][k] = x[0][k] * x[0][k];
        q[1][k] = x[1][k] * x[1][k];
    }
    __syncwarp();
    band_sums2_sm(T, q[0], q[1], nsf0, P->xsxx[0], P->xsxx[1], lane);
    __syncwarp();
    for (int w = 0; 32 * w < nrot; w++) {
        const int k = 32 * w + lane;
        int sm_ = 0, sd_ = 0;
        if (k < nrot) {
            float m = (x[0][k] + x[1][k]);
            float d = (x[0][k] - x[1][k]);
            if (m < 0.0f) { sm_ = 1; m = -m; }
            if (d < 0.0f) { sd_ = 1; d = -d; }
            x[0][k] = m;
            x[1][k] = d;
            xr[k] = m;
            xr[576 + k] = d;
            q[0][k] = m * m;
            q[1][k] = d * d;
        }
        const unsigned bm = __ballot_sync(0xffffffffu, sm_), bd = __ballot_sync(0xffffffffu, sd_);
        if (lane == 0) {
            P->sign[0][w] = bm;
            P->sign[1][w] = bd;
        }
    }
    __syncwarp();
    band_sums2_sm(T, q[0], q[1], nsf0, P->e2[0], P->e2[1], lane);
    __syncwarp();
    for (int ch = 0; ch < 2; ch++) {
        const int nmax = T->cfg.nbmax2[ch];
        const int hi = nmax > nrot ? nmax : nrot;
        for (int k = lane; k < hi; k += 32) {
            float v;
            if (k < nmax) v = pow34(T, x[ch][k]);
            else v = q[ch][k];  // (k < nrot: the square long_prepare leaves there)
            q[ch][k] = v;
            P->x34[ch][k] = v;
        }
    }
    __syncwarp();
    for (int ch = 0; ch < nch; ch++) band_bounds_sm(T, P, q[ch], ch, T->cfg.nsf2[ch], lane);
}

__global__ void __launch_bounds__(128) k_prepare(const EncTables *tabs, const StreamDev *st, ChunkBufs cb, int K0,
                                                 int nstreams) {
    __shared__ __align__(16) float s_x[4][2][576];
    __shared__ __align__(16) float s_q[4][2][576];
    __shared__ unsigned long long s_bar[4];
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int q = (int)(wid % cb.NG), s = (int)(wid / cb.NG);
    if (s >= nstreams) return;
    const StreamDev sd = st[s];
    if (K0 + q >= sd.ngran) return;
    const long long o = (long long)s * cb.NG + q;
    if (cb.gi[o].block_type == 2) return;
    const EncTables *T = tabs + sd.cfg;
    if (T->cfg.allocator == 1) return;  // CBitAllo1 strips signs / rotates for itself
    // the flag the allocator is called with (mp3enc.cpp:1556 / :1880: MPEG-2 mono passes the configured ms_flag)
    const int ms = (T->cfg.h_id == 0 && sd.nch != 2) ? T->cfg.ms_flag : (int)cb.ms[o];
    long_prepare_warp(T, ms, cb.xr + o * 2 * 576, cb.prep + o, s_x[wl], s_q[wl], &s_bar[wl], lane);
}

__global__ void k_prepare_init(int *msmem, PsyState *psy, int nstreams) {
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= 2 * nstreams) return;
    psy_state_init(psy + id);
    if ((id & 1) == 0) msmem[id >> 1] = 0;
}

}  // namespace hmp3
