// MDCT-domain psychoacoustic model and M/S correlation measure.
// Reference: emapLong/emapShort emap.c:61-121, spd_smrLongEcho / spd_smrShort spdsmr.c:64-319,
// CBitAllo3::ms_correlation2 bitallo3.cpp:682-754, CBitAlloShort::ms_correlation2Short bitallos.cpp:377-415.
//
// The model is split where the only cross-granule dependency sits:
//   stage 1 (stateless, parallel over every granule-channel): partition energies, spreading,
//            tonality offset -> unclamped thresholds;
//   stage 2 (O(npart) per granule, sequential in time per channel): pre-echo clamp against the
//            previous granule's thresholds (`echo` memory) and the merge to coder bands.
#pragma once
#include "enc_tables.h"

namespace hmp3 {

struct SigMask { float sig, mask; };

// Stage-1 result of one granule-channel.
struct PsyRaw {
    // long : e[i] = absolute threshold + partition energy, thr[i] = scaled spread threshold (i < npart)
    // short: thr[w*16 + m] = spread threshold of window w, coder band m
    float e[44];
    float thr[48];
};

HMP3_HD int iabs(int x) { return x < 0 ? -x : x; }

// ---- long blocks, stage 1 (emap.c:96-121; spdsmr.c:188-277)
HMP3_HD void psy_long_stage1(const EncTables *T, const float *xr, PsyRaw *R) {
    const int nmap = T->psy_emap_n_l;
    const int npart = T->psy_npart_l;
    const int npart2 = (npart + 1) & ~1;
    const float *w = T->w_spd_l;
    float xtab[44], etab[44];
    int mbetab[44];
    {
        int k = 0;
        for (int j = 0; j < 44; j++) {
            float s = 0.0f;
            if (j < nmap) {
                int n = T->psy_nsum_l[j];
                for (int q = 0; q < n; q++, k++) s += xr[k] * xr[k];
            }
            if (j < npart2) {
                float t = w[j] + s;
                etab[j] = t;
                int mbe = mb_log(T, t);
                mbetab[j] = mbe;
                xtab[j] = mb_exp(T, (int)(0.30f * mbe));
            } else {
                etab[j] = 0.0f;
                mbetab[j] = 0;
                xtab[j] = 0.0f;
            }
        }
    }
    float stab[44];
    int nsnr = 0, snrvar = 0, totsnr = 0, snr0 = 0;
    for (int i = 0; i < npart; i++) {
        int p = T->spd_off_l[i], n = T->spd_cnt_l[i], k = T->spd_w0_l[i];
        float s = 0.1f;
        for (int j = 0; j < n; j++, k++) s += w[k] * xtab[p + j];
        s = (0.03f * 0.1f * 0.35f) * mb_exp(T, (int)((1.0f / 0.30f) * mb_log(T, s))) + w[i];
        stab[i] = s;
        int snr = mbetab[i] - mb_log(T, w[i] + s);
        if (snr > 0) nsnr++;
        totsnr += (snr > -200 ? snr : -200);
        snrvar += iabs(snr - snr0);
        snr0 = snr;
    }
    for (int i = npart; i < 44; i++) stab[i] = 0.0f;
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
    int dm0 = (300 - d) >> 4;
    int m = 0;
    for (int i = 0; i < npart; i += 2, m++) {
        int t13 = (m - 13) > 0 ? (m - 13) : 0;
        int dm = dm0 * t13 > 0 ? dm0 * t13 : 0;
        float a = mb_exp(T, d + dm);
        R->thr[i] = a * stab[i];
        R->thr[i + 1] = a * stab[i + 1];
    }
    for (int i = 0; i < 44; i++) R->e[i] = etab[i];
}

// ---- long blocks, stage 2: pre-echo clamp + merge of partition pairs to coder bands (spdsmr.c:279-316)
HMP3_HD void psy_long_stage2(const EncTables *T, const PsyRaw *R, float *echo /*[64] state*/, int block_type,
                             SigMask *sm) {
    const int npart = T->psy_npart_l;
    int m = 0;
    for (int i = 0; i < npart; i += 2, m++) {
        float s1 = R->thr[i];
        float t = echo[i];
        echo[i] = (float)(2.0 * s1);
        if (block_type != 3) {
            if (s1 > t) {
                float x = 0.1f * s1;
                s1 = t;
                if (s1 < x) s1 = x;
            }
        }
        float s2 = R->thr[i + 1];
        t = echo[i + 1];
        echo[i + 1] = 2.0f * s2;
        if (block_type != 3) {
            if (s2 > t) {
                float x = 0.1f * s2;
                s2 = t;
                if (s2 < x) s2 = x;
            }
        }
        float e0 = R->e[i], e1 = R->e[i + 1];
        float emax = e0;
        if (emax < e1) emax = e1;
        sm[m].sig = e0 + e1;
        sm[m].mask = (e0 * s1 + e1 * s2) / emax;
    }
}

// ---- short blocks, stage 1 (emap.c:61-92; spdsmr.c:64-108): linear spreading per window
HMP3_HD void psy_short_stage1(const EncTables *T, const float *xr /*[3][192]*/, PsyRaw *R) {
    const int nmap = T->psy_emap_n_s;
    const int npart = T->psy_npart_s;
    const float *w = T->w_spd_s;
    float e[3][32];
    {
        int k = 0;
        for (int j = 0; j < 32; j++) {
            float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
            if (j < nmap) {
                int n = T->psy_nsum_s[j];
                for (int q = 0; q < n; q++, k++) {
                    s0 += xr[k] * xr[k];
                    s1 += xr[192 + k] * xr[192 + k];
                    s2 += xr[384 + k] * xr[384 + k];
                }
            }
            e[0][j] = s0;
            e[1][j] = s1;
            e[2][j] = s2;
        }
    }
    int k = 0, m = 0;
    for (int i = 0; i < npart; i += 2, m++) {
        int p = T->spd_off_s[i], n = T->spd_cnt_s[i];
        float s0 = 0.5f, s1 = 0.5f, s2 = 0.5f;
        for (int j = 0; j < n; j++, k++) {
            s0 += w[k] * e[0][p + j];
            s1 += w[k] * e[1][p + j];
            s2 += w[k] * e[2][p + j];
        }
        p = T->spd_off_s[i + 1];
        n = T->spd_cnt_s[i + 1];
        float t0 = 0.5f, t1 = 0.5f, t2 = 0.5f;
        for (int j = 0; j < n; j++, k++) {
            t0 += w[k] * e[0][p + j];
            t1 += w[k] * e[1][p + j];
            t2 += w[k] * e[2][p + j];
        }
        R->thr[m] = s0 + t0;
        R->thr[16 + m] = s1 + t1;
        R->thr[32 + m] = s2 + t2;
    }
}

// ---- short blocks, stage 2: window-to-window pre-echo clamp (spdsmr.c:110-181).  sm is [3][12].
HMP3_HD void psy_short_stage2(const EncTables *T, const PsyRaw *R, float *echo, int block_type_prev, SigMask *sm) {
    const int npart = T->psy_npart_s;
    const int mpart = (npart + 1) >> 1;
    for (int i = 0; i < mpart; i++) {
        float k0 = R->thr[i], k1 = R->thr[16 + i], k2 = R->thr[32 + i];
        float m0 = echo[i];
        float m1 = (float)(2.0 * k0);
        float m2 = (float)(2.0 * k1);
        echo[i] = (float)(2.0 * k2);
        if (block_type_prev == 2) {
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

// ---- M/S correlation measure of one granule, without the frame-to-frame hysteresis term
// (bitallo3.cpp:698-744).  x0/x1 = left/right spectra, long blocks.
HMP3_HD int ms_measure_long(const EncTables *T, const float *x0, const float *x1) {
    int cm = 0, k = 0;
    const int nsf = T->cfg.nsf[0];
    for (int i = 0; i < nsf; i++) {
        int n = T->nBand_l[i];
        float el = 100.0f, er = 100.0f, t = 0.0f;
        for (int j = 0; j < n; j++, k++) {
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
        cm += n * (mblr - mbsd);
    }
    return cm;
}
// short blocks (bitallos.cpp:377-415); x laid out [3][192]
HMP3_HD int ms_measure_short(const EncTables *T, const float *x0, const float *x1) {
    int d = 0;
    const int nsf = T->cfg.nsf_s[0];
    for (int w = 0; w < 3; w++) {
        int k = 0;
        for (int i = 0; i < nsf; i++) {
            int n = T->nBand_s[i];
            float s0 = 0.0f, s1 = 0.0f;
            for (int j = 0; j < n; j++, k++) {
                float a = x0[192 * w + k] * x0[192 * w + k];
                float b = x1[192 * w + k] * x1[192 * w + k];
                s0 += (a + b);
                a = a - b;
                if (a < 0.0f) a = -a;
                s1 += a;
            }
            if ((double)s1 > 0.80 * (double)s0) d++;
            if ((double)s1 > 0.95 * (double)s0) d += 2;
        }
    }
    return (nsf - d) << 10;
}

}  // namespace hmp3
