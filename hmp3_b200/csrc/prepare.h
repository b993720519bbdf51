// Phase A, second half: the parts of the reference's rate-loop prologue and psychoacoustic model that do not
// depend on the serial state (bit reservoir, MNR feedback, noise-target adaptation) and can therefore run in
// parallel over all granules, plus the two O(1)-state scans that feed them:
//   * M/S decision with its hysteresis memory              (bitallo3.cpp:691-696, 745-753; mp3enc.cpp:1537-1546)
//   * psychoacoustic stage 2 (pre-echo memory -> sig/mask)  (spdsmr.c:109-113, 279-319)
//   * long-block "startup" heavy part: sign strip or L/R -> M/S rotation, band energies, |x|^(3/4), band maxima
//     and the quantiser step range                          (bitallo3.cpp:816-898, 902-1065)
// The serial stage (rate_long.h) then only derives the noise targets from these.
#pragma once
#include "psy_core.h"

namespace hmp3 {

constexpr int kGminOffset = 70;  // bitallos.h:61

// Per long-block granule, both channels.
struct PrepGranule {
    float x34[2][576];      // |x|^(3/4) (also the scratch for the per-line squares while preparing)
    unsigned sign[2][18];   // sign bit of line k = bit (k & 31) of word (k >> 5); lines not visited read 0
    float xsxx[2][22];      // band energies of the left / right spectra
    float e2[2][22];        // M/S granules: band energies of mid / side
    float x34max[2][22];
    int gzero[2][22], gmin[2][22];
    int nlines[2];          // lines per channel whose sign / magnitude were rewritten
};

// Persistent per-(stream, channel) state of psychoacoustic stage 2.
struct PsyState {
    float echo[64];      // CMp3Enc::ecsave[ch][0] (init 1e20)
    SigMask sm[36];      // CMp3Enc::sig_mask[ch] (entries a granule type does not write keep their value)
};
HMP3_HD void psy_state_init(PsyState *p) {
    for (int i = 0; i < 64; i++) p->echo[i] = 1.0e20f;
    for (int i = 0; i < 36; i++) p->sm[i].sig = p->sm[i].mask = 100.0f;
}

// One granule of the M/S hysteresis scan: returns the correlation measure with memory applied.
HMP3_HD int ms_scan_step(int *memory, int block_type, int ms_raw) {
    if (block_type == 2) {
        *memory = 0;
        return ms_raw;
    }
    const int cm = ms_raw + *memory;
    *memory = (cm > 0) ? 5000 : -5000;
    return cm;
}

#if HMP3_COOP
// sums of v[] over each of the first nbands long bands, accumulated in line order: bands dealt over the lanes
HMP3_HD void long_band_sums(const EncTables *T, const float *v, int nbands, float *out) {
    HMP3_SYNC();
    for (int i = HMP3_LANE; i < nbands; i += HMP3_W) {
        out[i] = sum_seq(v + T->startBand_l[i], T->nBand_l[i], 0.0f);
    }
    HMP3_SYNC();
}
// Two rows at once: a band's ordered sum is a dependent chain as long as the band (up to ~100 lines for one lane
// while the others wait), so two independent chains per pass halve the passes; each sum keeps its own order.
HMP3_HD void long_band_sums2(const EncTables *T, const float *v0, const float *v1, int nbands, float *out0, float *out1) {
    HMP3_SYNC();
    for (int i = HMP3_LANE; i < nbands; i += HMP3_W) {
        const int k0 = T->startBand_l[i], n = T->nBand_l[i];
        float a = 0.0f, b = 0.0f;
        for (int k = k0; k < k0 + n; k++) {
            a += v0[k];
            b += v1[k];
        }
        out0[i] = a;
        out1[i] = b;
    }
    HMP3_SYNC();
}
#endif

// band maxima of |x|^(3/4) and the step range [gmin, gzero] (bitallo3.cpp:881-896)
HMP3_FN void long_prepare_bounds(const EncTables *T, PrepGranule *P, int ch, int nbands) {
#if HMP3_COOP
    HMP3_SYNC();
    for (int i = HMP3_LANE; i < nbands; i += HMP3_W) {
#else
    for (int i = 0; i < nbands; i++) {
#endif
        const float *y = P->x34[ch] + T->startBand_l[i];
        const int n = T->nBand_l[i];
        float m = 0.0f;
        for (int k = 0; k < n; k++)
            if (y[k] > m) m = y[k];
        P->x34max[ch][i] = m;
        float t = (0.017716950f * mb_log(T, m) + (104.585000f - 100.0f + 8.0f));
        int g0 = round_away(t);
        if (g0 < 0) g0 = 0;
        P->gzero[ch][i] = g0;
        P->gmin[ch][i] = (g0 - kGminOffset) > 0 ? (g0 - kGminOffset) : 0;
    }
    HMP3_SYNC();
}

// The state-free part of CBitAllo3::startup (ms == 0) / startup_ms2 (ms != 0) for one long-block granule.
// xr [2][576] is modified in place exactly as the reference does (magnitudes; mid/side without 1/sqrt2).
HMP3_FN void long_prepare(const EncTables *T, int ms, float *xr, PrepGranule *P) {
    if (T->cfg.allocator == 1) return;  // CBitAllo1 strips signs / rotates for itself (smr_adj, bitallo1.cpp:640-908)
    const int nch = T->cfg.nchan;
    for (int ch = 0; ch < 2; ch++)
#if HMP3_COOP
        for (int w = HMP3_LANE; w < 18; w += HMP3_W) P->sign[ch][w] = 0;
    HMP3_SYNC();
#else
        for (int w = 0; w < 18; w++) P->sign[ch][w] = 0;
#endif
    if (!ms) {
        for (int ch = 0; ch < nch; ch++) {
            float *x = xr + 576 * ch;
            float *sq = P->x34[ch];
            const int nb = T->cfg.nsf3[ch], nl = T->startBand_l[nb];
            P->nlines[ch] = nl;
#if HMP3_COOP
            for (int w = 0; 32 * w < nl; w++) {
                unsigned bits = 0;
                for (int h = 0; h < 32 / HMP3_W; h++) {
                    const int k = 32 * w + HMP3_W * h + HMP3_LANE;
                    int s = 0;
                    if (k < nl) {
                        float v = x[k];
                        if (!(v >= 0.0f)) { s = 1; v = -v; x[k] = v; }
                        sq[k] = v * v;
                    }
                    bits |= gballot(s) << (HMP3_W * h);
                }
                if (HMP3_LANE == 0) P->sign[ch][w] = bits;
            }
            long_band_sums(T, sq, nb, P->xsxx[ch]);
#else
            for (int i = 0; i < nb; i++) {
                const int n = T->nBand_l[i], k0 = T->startBand_l[i];
                float e = 0.0f;
                for (int k = k0; k < k0 + n; k++) {
                    if (!(x[k] >= 0.0f)) { P->sign[ch][k >> 5] |= 1u << (k & 31); x[k] = -x[k]; }
                    e += x[k] * x[k];
                }
                P->xsxx[ch][i] = e;
            }
#endif
        }
        for (int ch = 0; ch < nch; ch++) {
            const float *x = xr + 576 * ch;
#if HMP3_COOP
            HMP3_SYNC();
            for (int k = HMP3_LANE; k < T->cfg.nbmax3[ch]; k += HMP3_W) P->x34[ch][k] = pow34(T, x[k]);
#else
            for (int k = 0; k < T->cfg.nbmax3[ch]; k++) P->x34[ch][k] = pow34(T, x[k]);
#endif
            long_prepare_bounds(T, P, ch, T->cfg.nsf3[ch]);
        }
        return;
    }
    const int nsf0 = T->cfg.nsf[0];
    const int nl = T->startBand_l[nsf0];
    const int nrot = nl + (T->cfg.hf_flag ? T->nBand_l[21] : 0);  // the pseudo band above sfb 21 is rotated too
    P->nlines[0] = P->nlines[1] = nrot;
    float *sq0 = P->x34[0], *sq1 = P->x34[1];
#if HMP3_COOP
    for (int k = HMP3_LANE; k < nl; k += HMP3_W) {
        sq0[k] = xr[k] * xr[k];
        sq1[k] = xr[576 + k] * xr[576 + k];
    }
    long_band_sums2(T, sq0, sq1, nsf0, P->xsxx[0], P->xsxx[1]);
    for (int w = 0; 32 * w < nrot; w++) {
        unsigned bm = 0, bd = 0;
        for (int h = 0; h < 32 / HMP3_W; h++) {
            const int k = 32 * w + HMP3_W * h + HMP3_LANE;
            int sm_ = 0, sd_ = 0;
            if (k < nrot) {
                float m = (xr[k] + xr[576 + k]);
                float d = (xr[k] - xr[576 + k]);
                if (m < 0.0f) { sm_ = 1; m = -m; }
                if (d < 0.0f) { sd_ = 1; d = -d; }
                xr[k] = m;
                xr[576 + k] = d;
                sq0[k] = m * m;
                sq1[k] = d * d;
            }
            bm |= gballot(sm_) << (HMP3_W * h);
            bd |= gballot(sd_) << (HMP3_W * h);
        }
        if (HMP3_LANE == 0) {
            P->sign[0][w] = bm;
            P->sign[1][w] = bd;
        }
    }
    long_band_sums2(T, sq0, sq1, nsf0, P->e2[0], P->e2[1]);
    HMP3_SYNC();
    for (int ch = 0; ch < 2; ch++)
        for (int k = HMP3_LANE; k < T->cfg.nbmax2[ch]; k += HMP3_W) P->x34[ch][k] = pow34(T, xr[576 * ch + k]);
#else
    for (int i = 0; i < nsf0; i++) {
        const int n = T->nBand_l[i], k0 = T->startBand_l[i];
        float el = 0.0f, er = 0.0f;
        for (int k = k0; k < k0 + n; k++) {
            el += xr[k] * xr[k];
            er += xr[576 + k] * xr[576 + k];
        }
        P->xsxx[0][i] = el;
        P->xsxx[1][i] = er;
    }
    for (int k = 0; k < nrot; k++) {
        float m = (xr[k] + xr[576 + k]);
        float d = (xr[k] - xr[576 + k]);
        if (m < 0.0f) { P->sign[0][k >> 5] |= 1u << (k & 31); m = -m; }
        if (d < 0.0f) { P->sign[1][k >> 5] |= 1u << (k & 31); d = -d; }
        xr[k] = m;
        xr[576 + k] = d;
    }
    for (int i = 0; i < nsf0; i++) {
        const int n = T->nBand_l[i], k0 = T->startBand_l[i];
        float em = 0.0f, ed = 0.0f;
        for (int k = k0; k < k0 + n; k++) {
            em += xr[k] * xr[k];
            ed += xr[576 + k] * xr[576 + k];
        }
        P->e2[0][i] = em;
        P->e2[1][i] = ed;
    }
    for (int ch = 0; ch < 2; ch++)
        for (int k = 0; k < T->cfg.nbmax2[ch]; k++) P->x34[ch][k] = pow34(T, xr[576 * ch + k]);
#endif
#if HMP3_COOP
    if (nch == 2 && T->cfg.nsf2[0] == T->cfg.nsf2[1]) {  // both channels' band maxima in one pass (two chains per lane)
        HMP3_SYNC();
        for (int i = HMP3_LANE; i < T->cfg.nsf2[0]; i += HMP3_W) {
            const int k0 = T->startBand_l[i], n = T->nBand_l[i];
            const float *y0 = P->x34[0] + k0, *y1 = P->x34[1] + k0;
            float m0 = 0.0f, m1 = 0.0f;
            for (int k = 0; k < n; k++) {
                if (y0[k] > m0) m0 = y0[k];
                if (y1[k] > m1) m1 = y1[k];
            }
            for (int ch = 0; ch < 2; ch++) {
                const float m = ch ? m1 : m0;
                P->x34max[ch][i] = m;
                float t = (0.017716950f * mb_log(T, m) + (104.585000f - 100.0f + 8.0f));
                int g0 = round_away(t);
                if (g0 < 0) g0 = 0;
                P->gzero[ch][i] = g0;
                P->gmin[ch][i] = (g0 - kGminOffset) > 0 ? (g0 - kGminOffset) : 0;
            }
        }
        HMP3_SYNC();
        return;
    }
#endif
    for (int ch = 0; ch < nch; ch++) long_prepare_bounds(T, P, ch, T->cfg.nsf2[ch]);
}

}  // namespace hmp3
