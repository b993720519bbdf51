// TEST INFRASTRUCTURE ONLY.  Host build of the kernel bodies (hmp3_b200/csrc/*.h compiled by g++ with
// -ffp-contract=off) so that parity against the oracle can be checked in a container without a GPU.
// The product library never links or loads this file; the shipped path is CUDA-only.
#include <string.h>
#include <stdlib.h>
#include <vector>

#include "../../hmp3_b200/csrc/enc_init.h"
#include "../../hmp3_b200/csrc/analysis.h"

using namespace hmp3;

// DC-blocking input filter (filter2.c:112-147) over the whole stream incl. the zero tail, when -S1 is on
// (also the plain copy of float input when the filter is off: `pcmf` != null means the source is float PCM)
static std::vector<float> dc_filtered(const EncTables *T, const int16_t *pcm, const float *pcmf, long nsamples, int nch,
                                      long len, float tail = 0.0f) {
    std::vector<float> f;
    if (!T->cfg.filter_select) {
        if (pcmf) {
            f.assign((size_t)len * nch, tail);
            for (long n = 0; n < nsamples && n < len; n++)
                for (int c = 0; c < nch; c++) f[(size_t)n * nch + c] = pcmf[n * nch + c];
        }
        return f;
    }
    f.assign((size_t)len * nch, 0.0f);
    for (int c = 0; c < nch; c++) {
        float d = 0.0f;
        for (long n = 0; n < len; n++) {
            const float x = n < nsamples ? (pcmf ? pcmf[n * nch + c] : (float)pcm[n * nch + c]) : tail;
            const float t = (x - d);
            d = d + T->cfg.dc_alpha * t;
            f[(size_t)n * nch + c] = t;
        }
    }
    return f;
}
static void poly_granule(const EncTables *T, const int16_t *pcm, long nsamples, const std::vector<float> &pf, long len,
                         int nch, int c, long j, float *o) {
    for (int t = 0; t < 18; t++) {
        if (pf.empty()) polyphase_item(T, pcm, nsamples, nch, c, j, t, o);
        else polyphase_item_f(T, pf.data(), len, nch, c, j, t, o);
    }
}

extern "C" {

int sim_resolve(const hmp3_control *ec, int *out /*40*/) {
    EncTables *T = new EncTables;
    int unsup = 0;
    int r = build_tables(ec, T, &unsup);
    const EncConfig &C = T->cfg;
    int v[] = {C.nchan, C.h_id, C.sr_index, C.nband, C.band_limit, C.nsb, C.nsb_limit, C.nsb_hybrid, C.nsb_limit_ms1,
               C.ave_target_bits, C.framebytes, C.main_framebytes, C.side_bytes, C.pad_remainder, C.pad_divisor,
               C.ms_flag, C.is_flag, C.frame_driver, 0, C.ivbr_min, C.ivbr_max, C.vbr_pool_target,
               C.short_block_threshold, C.h_mode, C.br_index, C.totbitrate, C.samprate, C.band_limit_stereo,
               C.sf_bit_max, C.nsf_stereo, C.head[0], C.head[1], C.head[2], C.head[3], C.hf_flag,
               C.filter_select * 2 + (C.nchan - 1)};
    for (unsigned i = 0; i < sizeof(v) / sizeof(int); i++) out[i] = v[i];
    out[38] = unsup;
    out[39] = r;
    delete T;
    return r;
}

// named table copy-out for boundary tests
int sim_table(const hmp3_control *ec, const char *name, void *dst, int nbytes) {
    EncTables *T = new EncTables;
    int r = build_tables(ec, T, nullptr);
    if (!r) { delete T; return -1; }
    struct { const char *n; const void *p; int sz; } tabs[] = {
        {"win", T->win, sizeof(T->win)}, {"csa", T->csa, sizeof(T->csa)}, {"m18_w", T->m18_w, sizeof(T->m18_w)},
        {"m18_w2", T->m18_w2, sizeof(T->m18_w2)}, {"m18_c", T->m18_c, sizeof(T->m18_c)},
        {"m6_v", T->m6_v, sizeof(T->m6_v)}, {"m6_v2", T->m6_v2, sizeof(T->m6_v2)}, {"m6_c", &T->m6_c, 4},
        {"psy_nsum_l", T->psy_nsum_l, sizeof(T->psy_nsum_l)}, {"psy_nsum_s", T->psy_nsum_s, sizeof(T->psy_nsum_s)},
        {"spd_cnt_l", T->spd_cnt_l, sizeof(T->spd_cnt_l)}, {"spd_off_l", T->spd_off_l, sizeof(T->spd_off_l)},
        {"spd_cnt_s", T->spd_cnt_s, sizeof(T->spd_cnt_s)}, {"spd_off_s", T->spd_off_s, sizeof(T->spd_off_s)},
        {"w_spd_l", T->w_spd_l, sizeof(T->w_spd_l)}, {"w_spd_s", T->w_spd_s, sizeof(T->w_spd_s)},
        {"psy_n", &T->psy_npart_l, 16}, {"vbr_main_framebytes", T->cfg.vbr_main_framebytes, 64},
        {"cnt_lut", T->cnt_lut, sizeof(T->cnt_lut)}, {"gain", T->gain, sizeof(T->gain)},
        {"igain34", T->igain34, sizeof(T->igain34)}, {"ix43", T->ix43, sizeof(T->ix43)},
    };
    int ret = -2;
    for (auto &t : tabs)
        if (!strcmp(t.n, name)) {
            int n = t.sz < nbytes ? t.sz : nbytes;
            memcpy(dst, t.p, n);
            ret = n;
        }
    delete T;
    return ret;
}

// Phase A over a whole clip, sequential reference driver of the per-item routines.
// Encode granules K = 0..ngran-1.  Outputs (any may be null):
//   sbt    [ngran][nch][576]   P[K] (frequency-inverted, band-major)
//   ginfo  [ngran][4]          block_type, block_type_prev, short_cur, short_next
//   xr     [ngran][nch][576]
//   sigmask[ngran][nch][36][2] {sig,mask}
//   ms_raw [ngran]             M/S measure without hysteresis (long) / short measure
//   att    [ngran][nch][9]     attack energies of P[K]
//   raw    [ngran][nch][92]    psychoacoustic stage-1 record
int sim_analysis(const hmp3_control *ec, const int16_t *pcm_any, long nsamples, int ngran, float *sbt_out, int *ginfo,
                 float *xr_out, float *sigmask, int *ms_raw, int *att, float *raw_out) {
    const int is_float = 0;
    const float tail = 0.0f;
    EncTables *T = new EncTables;
    if (!build_tables(ec, T, nullptr)) { delete T; return -1; }
    const int nch = T->cfg.nchan;
    const int mpeg2 = T->cfg.h_id == 0;
    // P[j] for j = -3 .. ngran-1 (index j+3)
    std::vector<float> P((size_t)(ngran + 3) * nch * 576, 0.0f);
    std::vector<int> E((size_t)(ngran + 3) * nch * 9, 0);
    const std::vector<float> pf = dc_filtered(T, is_float ? nullptr : (const int16_t *)pcm_any, is_float ? (const float *)pcm_any : nullptr, nsamples, nch, 576L * ngran, tail);
    const int16_t *pcm = is_float ? nullptr : (const int16_t *)pcm_any;
    for (long j = -3; j < ngran; j++)
        for (int c = 0; c < nch; c++) {
            float *o = &P[((j + 3) * nch + c) * 576];
            poly_granule(T, pcm, nsamples, pf, 576L * ngran, nch, c, j, o);
            for (int k = 0; k < 9; k++) E[((j + 3) * nch + c) * 9 + k] = attack_energy(T, o, k, mpeg2);
        }
    SwitchState sw;
    switch_state_init(&sw);
    float echo[2][64];
    for (int c = 0; c < 2; c++)
        for (int i = 0; i < 64; i++) echo[c][i] = 1.0e20f;
    SigMask sm[2][36];
    for (int c = 0; c < 2; c++)
        for (int i = 0; i < 36; i++) sm[c][i].sig = sm[c][i].mask = 100.0f;
    std::vector<float> xr((size_t)nch * 576);
    PsyRaw raw;
    for (int K = 0; K < ngran; K++) {
        const int *e0 = &E[((K - 1 + 3) * nch + 0) * 9];
        const int *e1 = &E[((K - 1 + 3) * nch + (nch - 1)) * 9];
        GranuleInfo g = switch_step(T, &sw, e0, e1);
        if (ginfo) {
            ginfo[4 * K + 0] = g.block_type; ginfo[4 * K + 1] = g.block_type_prev;
            ginfo[4 * K + 2] = g.short_cur;  ginfo[4 * K + 3] = g.short_next;
        }
        for (int c = 0; c < nch; c++) {
            const float *prev = &P[((K - 3 + 3) * nch + c) * 576];
            const float *cur = &P[((K - 2 + 3) * nch + c) * 576];
            float *x = &xr[c * 576];
            for (int sb = 0; sb < 32; sb++) hybrid_item(T, prev, cur, g.block_type, sb, x);
            if (g.block_type != 2)
                for (int sb = 0; sb < 32; sb++) alias_item(T, sb, x);
            if (g.block_type != 2) {
                psy_long_stage1(T, x, &raw);
                psy_long_stage2(T, &raw, echo[c], g.block_type, sm[c]);
            } else {
                psy_short_stage1(T, x, &raw);
                psy_short_stage2(T, &raw, echo[c], g.block_type_prev, sm[c]);
            }
            if (raw_out) memcpy(raw_out + ((size_t)K * nch + c) * 92, &raw, sizeof(PsyRaw));
            if (xr_out) memcpy(xr_out + ((size_t)K * nch + c) * 576, x, 576 * sizeof(float));
            if (sbt_out) memcpy(sbt_out + ((size_t)K * nch + c) * 576, &P[((K + 3) * nch + c) * 576], 576 * sizeof(float));
            if (sigmask) memcpy(sigmask + ((size_t)K * nch + c) * 72, sm[c], 72 * sizeof(float));
            if (att) memcpy(att + ((size_t)K * nch + c) * 9, &E[((K + 3) * nch + c) * 9], 9 * sizeof(int));
        }
        if (ms_raw) {
            if (nch == 2)
                ms_raw[K] = (g.block_type != 2) ? ms_measure_long(T, &xr[0], &xr[576]) : ms_measure_short(T, &xr[0], &xr[576]);
            else ms_raw[K] = 0;
        }
    }
    delete T;
    return 0;
}

}  // extern "C"

#include "../../hmp3_b200/csrc/rate_phased.h"

extern "C" {

// Whole-clip encode on the host build of the kernel bodies (Phase A sequential + Phase B), CLI
// semantics.  Returns bytes written to `out`, or <0.  `trace` (optional) receives per granule:
// 27 GR ints x2 ch, sf_l 23 x2, sf_s 39 x2, ix 576 x2, ms flag, = 1333 ints per granule? (see tests/simmod.py)
long sim_encode_clip_any(const hmp3_control *ec, const void *pcm_any, int is_float, float tail, long nsamples,
                         unsigned char *out, long out_cap, int *trace, int max_trace_granules, int *nframes_out);
long sim_encode_clip(const hmp3_control *ec, const int16_t *pcm, long nsamples, unsigned char *out, long out_cap,
                     int *trace, int max_trace_granules, int *nframes_out) {
    return sim_encode_clip_any(ec, pcm, 0, 0.0f, nsamples, out, out_cap, trace, max_trace_granules, nframes_out);
}
long sim_encode_clip_any(const hmp3_control *ec, const void *pcm_any, int is_float, float tail, long nsamples,
                         unsigned char *out, long out_cap, int *trace, int max_trace_granules, int *nframes_out) {
    EncTables *T = new EncTables;
    if (!build_tables(ec, T, nullptr)) { delete T; return -1; }
    const int nch = T->cfg.nchan;
    const int mpeg2 = T->cfg.h_id == 0;
    long calls = (nsamples + 3 * 1153 + 1152) / 1152;  // the CLI's main loop (pipeline.cu: calls_for)
    int ngran_real = (int)(2 * calls);
    int ngran = ngran_real + 2 * 12;
    std::vector<float> P((size_t)(ngran + 3) * nch * 576, 0.0f);
    std::vector<int> E((size_t)(ngran + 3) * nch * 9, 0);
    const std::vector<float> pf = dc_filtered(T, is_float ? nullptr : (const int16_t *)pcm_any, is_float ? (const float *)pcm_any : nullptr, nsamples, nch, 576L * ngran, tail);
    const int16_t *pcm = is_float ? nullptr : (const int16_t *)pcm_any;
    for (long j = -3; j < ngran; j++)
        for (int c = 0; c < nch; c++) {
            float *o = &P[((j + 3) * nch + c) * 576];
            poly_granule(T, pcm, nsamples, pf, 576L * ngran, nch, c, j, o);
            for (int k = 0; k < 9; k++) E[((j + 3) * nch + c) * 9 + k] = attack_energy(T, o, k, mpeg2);
        }
    SwitchState sw;
    switch_state_init(&sw);
    std::vector<GranuleInfo> gi(ngran);
    std::vector<float> xr((size_t)ngran * 2 * 576, 0.0f);
    std::vector<PsyRaw> raw((size_t)ngran * 2);
    std::vector<int> msr(ngran, 0);
    for (int K = 0; K < ngran; K++) {
        const int *e0 = &E[((K - 1 + 3) * nch + 0) * 9];
        const int *e1 = &E[((K - 1 + 3) * nch + (nch - 1)) * 9];
        gi[K] = switch_step(T, &sw, e0, e1);
        for (int c = 0; c < nch; c++) {
            const float *prev = &P[((K - 3 + 3) * nch + c) * 576];
            const float *cur = &P[((K - 2 + 3) * nch + c) * 576];
            float *x = &xr[((size_t)K * 2 + c) * 576];
            for (int sb = 0; sb < 32; sb++) hybrid_item(T, prev, cur, gi[K].block_type, sb, x);
            if (gi[K].block_type != 2) {
                for (int sb = 0; sb < 32; sb++) alias_item(T, sb, x);
                psy_long_stage1(T, x, &raw[(size_t)K * 2 + c]);
            } else psy_short_stage1(T, x, &raw[(size_t)K * 2 + c]);
        }
        if (nch == 2) {
            const float *x0 = &xr[((size_t)K * 2) * 576];
            msr[K] = gi[K].block_type != 2 ? ms_measure_long(T, x0, x0 + 576) : ms_measure_short(T, x0, x0 + 576);
        }
    }
    // the scans and the prepare pass that follow stage 1 (on the device: k_ms_scan, k_psy_stage2, k_prepare)
    std::vector<signed char> msf(ngran + 2, 0);
    std::vector<SigMask> smk((size_t)(ngran + 2) * 72);
    std::vector<PrepGranule> prep(ngran + 2);
    {
        const bool stereo_ms = (nch == 2 && T->cfg.ms_flag);
        int mem = 0;
        for (int K = 0; K + 1 < ngran; K += 2) {
            int f0 = 0, f1 = 0;
            if (stereo_ms) {
                const int a = ms_scan_step(&mem, gi[K].block_type, msr[K]);
                const int b2 = ms_scan_step(&mem, gi[K + 1].block_type, msr[K + 1]);
                if (!mpeg2) f0 = f1 = ((a + b2) >= 0);
                else { f0 = (a >= 0); f1 = (b2 >= 0); }
            }
            msf[K] = (signed char)f0;
            msf[K + 1] = (signed char)f1;
        }
        PsyState ps[2];
        psy_state_init(&ps[0]);
        psy_state_init(&ps[1]);
        for (int K = 0; K < ngran; K++) {
            for (int c = 0; c < nch; c++) {
                if (gi[K].block_type != 2) psy_long_stage2(T, &raw[(size_t)K * 2 + c], ps[c].echo, gi[K].block_type, ps[c].sm);
                else psy_short_stage2(T, &raw[(size_t)K * 2 + c], ps[c].echo, gi[K].block_type_prev, ps[c].sm);
                for (int i = 0; i < 36; i++) smk[((size_t)K * 2 + c) * 36 + i] = ps[c].sm[i];
            }
            if (gi[K].block_type != 2) {
                const int ms = (mpeg2 && nch != 2) ? T->cfg.ms_flag : (int)msf[K];
                long_prepare(T, ms, &xr[(size_t)K * 2 * 576], &prep[K]);
            }
        }
    }
    RateState *R = new RateState;
    RateCold *Rcold = new RateCold;
    rate_state_init(T, R, Rcold);
    const bool nested = getenv("HMP3_SIM_NESTED") && atoi(getenv("HMP3_SIM_NESTED")) != 0;
    std::vector<unsigned char> mainbuf((size_t)(ngran + 4) * 2100, 0);
    std::vector<FrameRec> frames(ngran + 4);
    // run granule pair by pair so that traces can be taken after each call; the packing pass of the frames a
    // call recorded follows it immediately (on the device it is a separate kernel over all frames of a chunk)
    std::vector<PackGc> pack(4);
    int bad = 0;
    for (int K = 0; K + 1 < ngran && !R->finished; K += 2) {
        const int f0 = R->frames;
        // allocator 0 runs the phase machine (what the device kernel schedules); HMP3_SIM_NESTED=1 the nested drivers
        if (T->cfg.allocator == 0 && !nested)
            rate_run_chunk_phased(T, R, K, 2, ngran, ngran_real, &gi[K], &xr[(size_t)K * 2 * 576], &smk[(size_t)K * 72],
                                  &prep[K], &msf[K], pack.data(), frames.data());
        else
            rate_run_chunk(T, R, K, 2, ngran, ngran_real, &gi[K], &xr[(size_t)K * 2 * 576], &smk[(size_t)K * 72], &prep[K],
                           &msf[K], pack.data(), frames.data());
        for (int f = f0; f < R->frames; f++)
            bad |= pack_frame(T, &frames[f], pack.data() + (size_t)(frames[f].granule0 - K) * 2, mainbuf.data());
        if (trace)
            for (int q = 0; q < 2; q++) {
                int Kq = K + q;
                if (Kq >= max_trace_granules) continue;
                int *t = trace + (size_t)Kq * 1400;
                memset(t, 0, 1400 * sizeof(int));
                for (int c = 0; c < nch; c++) {
                    memcpy(t + 27 * c, &R->gr[q][c], 27 * sizeof(int));
                    memcpy(t + 54 + 23 * c, R->sf[q][c].l, 23 * sizeof(int));
                    memcpy(t + 100 + 39 * c, R->sf[q][c].s, 39 * sizeof(int));
                }
                // ix is shared by both granules of the call: only meaningful for the last granule processed
                for (int k = 0; k < 2 * 576; k++) t[200 + k] = (&R->ix[0][0])[k];
                t[1360] = R->L.mnr;
                t[1361] = R->byte_pool;
                t[1362] = frames[R->frames - 1].head[3];
                t[1363] = Rcold->A1.bitadjust_save[0];
                t[1364] = Rcold->A1.bitadjust_save[1];
                t[1365] = Rcold->A1.call_count;
                memcpy(&t[1366], &Rcold->A1.ave_alpha_nmr, 4);
                memcpy(&t[1367], &Rcold->A1.alpha_nmr, 4);
                for (int i = 0; i < 21; i++) t[1368 + i] = Rcold->A1.gsf_save[1][i];
            }
    }
    if (bad) { delete R; delete Rcold; delete T; return -3; }
    long total = 0;
    const int nf = R->frames_done;
    for (int f = 0; f < nf; f++) {
        const FrameRec &fr = frames[f];
        int sz = frame_bytes(T, &fr);
        if (total + sz > out_cap) { total = -2; break; }
        memcpy(out + total, fr.head, 4);
        memcpy(out + total + 4, fr.side, T->cfg.side_bytes);
        memcpy(out + total + 4 + T->cfg.side_bytes, mainbuf.data() + fr.main_start, fr.mf_bytes);
        total += sz;
    }
    if (nframes_out) *nframes_out = nf;
    delete R; delete Rcold;
    delete T;
    return total;
}

}  // extern "C"
