// Rate loop for short blocks (block type 2): three windows of 192 lines per channel, per-window gains
// (subblock_gain), no pre-emphasis, fixed Huffman regions.  Behaviour follows CBitAlloShort
// (bitallos.cpp:202-1503, bitallosc.cpp:296-487, cnts.c:85-305); integer outputs are bit-exact given
// identical inputs.
#pragma once
#include "rate_long.h"

namespace hmp3 {

struct ShortRate {
    int mnr, nchan, ms;
    int max_bits, max_target, min_target, target, pool_bits, active_lines, feedback_bits, delta_mnr;
    int huff_bits[2];
    float xsxx[2][3][16], x34max[2][3][16];
    int noise0[2][3][16], nt[2][3][16], noise[2][3][16], snr[2][3][16];
    int ixmax[2][3][16], gzero[2][3][16], gmin[2][3][16], gsf[2][3][16], sf[2][3][16], active[2][3][16];
    int subgain[2][3], G[2][3], GG[2], sf_scale[2];
    RegionPlan plan[2];
    float x34[2][3][192];
    alignas(16) QLine ix[2][3][192];
    unsigned char sign[2][3][192];
    unsigned char sign_t[2][576];     // the signs in transmission order (one byte per line), before they are packed
    alignas(16) QLine quad_scratch[580];
};

HMP3_FN void short_rate_init(ShortRate *S) {
    // everything the allocator reads before writing starts at zero (bitallos.cpp:85-110, 128-198)
    unsigned char *p = (unsigned char *)S;
    for (unsigned i = 0; i < sizeof(ShortRate); i++) p[i] = 0;
}

// The short-block loops run over (channel, window, band) items that are independent of each other and short
// (4..30 lines).  On the device the items are dealt over the lanes of the stream's group and each lane runs the
// plain sequential code for its items; the host build visits them in order.  it = (ch * 3 + w) * 16 + i.
#if HMP3_COOP
#define HMP3_SHORT_ITEMS(it) for (int it = HMP3_LANE; it < 96; it += HMP3_W)
#else
#define HMP3_SHORT_ITEMS(it) for (int it = 0; it < 96; it++)
#endif

HMP3_FN void short_step_bounds(const EncTables *T, ShortRate *S) {
    HMP3_SYNC();
    HMP3_SHORT_ITEMS(it) {
        const int ch = it / 48, w = (it >> 4) % 3, i = it & 15;
        if (ch >= S->nchan || i >= T->cfg.nsf_s[ch]) continue;
        const float *y = S->x34[ch][w] + T->startBand_s[i];
        const int n = T->nBand_s[i];
        float m = 0.0f;
        for (int k = 0; k < n; k++)
            if (y[k] > m) m = y[k];
        S->x34max[ch][w][i] = m;
        S->gzero[ch][w][i] = imax_(0, round_away((0.017716950f * mb_log(T, m) + (104.585000f - 100.0f + 8.0f))));
        S->gmin[ch][w][i] = imax_(0, S->gzero[ch][w][i] - kGminOffset);
    }
    HMP3_SYNC();
}

// pull targets of audible bands halfway to their mean when the mean is high (bitallos.cpp:700-745)
HMP3_FN void short_flatten_targets(const EncTables *T, ShortRate *S) {
    int na = 1, a = 0;
    for (int ch = 0; ch < S->nchan; ch++)
        for (int w = 0; w < 3; w++)
            for (int i = 0; i < T->cfg.nsf_s[ch]; i++)
                if (S->snr[ch][w][i] > 0) { a += S->nt[ch][w][i]; na++; }
    a = a / na;
    if (a <= 500) return;
    for (int ch = 0; ch < S->nchan; ch++)
        for (int w = 0; w < 3; w++)
            for (int i = 0; i < T->cfg.nsf_s[ch]; i++)
                if (S->snr[ch][w][i] > 0) S->nt[ch][w][i] = (S->nt[ch][w][i] + a) >> 1;
}

// xr: [2][3][192] (window-major within the channel); sm: [2][3][12]
HMP3_FN void short_startup_lr(const EncTables *T, ShortRate *S, float *xr, const SigMask *sm) {  // bitallos.cpp:455-546
    const int mnr = S->mnr;
    int lines = 0;
    HMP3_SYNC();
    HMP3_SHORT_ITEMS(it) {
        const int ch = it / 48, w = (it >> 4) % 3, i = it & 15;
        if (ch >= S->nchan || i >= T->cfg.nsf_s[ch]) continue;
        float *x = xr + 576 * ch + 192 * w + T->startBand_s[i];
        unsigned char *s = S->sign[ch][w] + T->startBand_s[i];
        const int n = T->nBand_s[i];
        float e = 0.0f;
        for (int k = 0; k < n; k++) {
            if (x[k] >= 0.0f) s[k] = 0;
            else { s[k] = 1; x[k] = -x[k]; }
            e += x[k] * x[k];
        }
        S->xsxx[ch][w][i] = e;
        const int cbw = T->log_cbw_s[i];
        S->noise0[ch][w][i] = mb_log(T, e) - cbw;
        if (S->noise0[ch][w][i] < -2000) {
            S->nt[ch][w][i] = S->noise0[ch][w][i] + 1000;
            S->snr[ch][w][i] = -1000;
        } else {
            int mask = (mb_log(T, sm[36 * ch + 12 * w + i].mask) - cbw);
            S->nt[ch][w][i] = nt_dropout_guard(S->noise0[ch][w][i], mask - mnr);
            S->snr[ch][w][i] = S->noise0[ch][w][i] - S->nt[ch][w][i];
            lines += n;
        }
    }
#if HMP3_COOP
    lines = gsum(lines);
#endif
    S->active_lines = lines;
    HMP3_SYNC();
    short_flatten_targets(T, S);
    HMP3_SYNC();
    for (int ch = 0; ch < S->nchan; ch++)
        for (int w = 0; w < 3; w++)
#if HMP3_COOP
            for (int k = HMP3_LANE; k < T->cfg.nbmax_s[ch]; k += HMP3_W)
#else
            for (int k = 0; k < T->cfg.nbmax_s[ch]; k++)
#endif
                S->x34[ch][w][k] = pow34(T, xr[576 * ch + 192 * w + k]);
    short_step_bounds(T, S);
}

HMP3_FN void short_startup_ms(const EncTables *T, ShortRate *S, float *xr, const SigMask *sm) {  // bitallos.cpp:549-697
    const int nsf0 = T->cfg.nsf_s[0];
    const int so = (int)(&S->sign[1][0][0] - &S->sign[0][0][0]);  // sign[1] follows sign[0]
    int lines = 0;
    HMP3_SYNC();
    HMP3_SHORT_ITEMS(it) {
        const int w = it >> 4, i = it & 15;  // items = (window, band); both channels are handled together
        if (w >= 3 || i >= nsf0) continue;
        float *x = xr + 192 * w + T->startBand_s[i];
        unsigned char *s = S->sign[0][w] + T->startBand_s[i];
        const int n = T->nBand_s[i];
        float el = 0.0f, er = 0.0f;
        for (int k = 0; k < n; k++) {
            el += x[k] * x[k];
            er += x[576 + k] * x[576 + k];
        }
        for (int k = 0; k < n; k++) {
            float m = (x[k] + x[576 + k]);
            float d = (x[k] - x[576 + k]);
            s[k] = s[so + k] = 0;
            if (m < 0.0f) { s[k] = 1; m = -m; }
            if (d < 0.0f) { s[so + k] = 1; d = -d; }
            x[k] = m;
            x[576 + k] = d;
        }
        float em = 0.0f, ed = 0.0f;
        for (int k = 0; k < n; k++) {
            em += x[k] * x[k];
            ed += x[576 + k] * x[576 + k];
        }
        S->xsxx[0][w][i] = el;
        S->xsxx[1][w][i] = er;
        const int cbw = T->log_cbw_s[i];
        int ntl, ntr;
        int n0l = mb_log(T, el) - cbw;
        if (n0l < -2000) ntl = 10000;
        else {
            ntl = nt_dropout_guard(n0l, (mb_log(T, sm[12 * w + i].mask) - cbw) - S->mnr);
            lines += n;
        }
        int n0r = mb_log(T, er) - cbw;
        if (n0r < -2000) ntr = 10000;
        else {
            ntr = nt_dropout_guard(n0r, (mb_log(T, sm[36 + 12 * w + i].mask) - cbw) - S->mnr);
            lines += n;
        }
        const int nsum = mb_log(T, em) - cbw, ndiff = mb_log(T, ed) - cbw;
        S->noise0[0][w][i] = nsum;
        S->noise0[1][w][i] = ndiff;
        const int xnt = imin_(ntr, ntl) + 300;
        S->nt[1][w][i] = S->nt[0][w][i] = xnt;
        if (ndiff < xnt) S->nt[0][w][i] = mb_logsub(T, xnt, ndiff) - 200;
        if (nsum < xnt) S->nt[1][w][i] = mb_logsub(T, xnt, nsum) - 200;
        S->snr[0][w][i] = S->noise0[0][w][i] - S->nt[0][w][i];
        S->snr[1][w][i] = S->noise0[1][w][i] - S->nt[1][w][i];
    }
#if HMP3_COOP
    lines = gsum(lines);
#endif
    S->active_lines = lines;
    HMP3_SYNC();
    short_flatten_targets(T, S);
    HMP3_SYNC();
    for (int w = 0; w < 3; w++)
        for (int ch = 0; ch < 2; ch++)
#if HMP3_COOP
            for (int k = HMP3_LANE; k < T->cfg.nbmax_s[ch]; k += HMP3_W)
#else
            for (int k = 0; k < T->cfg.nbmax_s[ch]; k++)
#endif
                S->x34[ch][w][k] = pow34(T, xr[576 * ch + 192 * w + k]);
    short_step_bounds(T, S);
}

HMP3_FN void short_seek_initial(const EncTables *T, ShortRate *S) {  // bitallos.cpp:748-776
    HMP3_SYNC();
    HMP3_SHORT_ITEMS(it) {
        const int ch = it / 48, w = (it >> 4) % 3, i = it & 15;
        if (ch >= S->nchan || i >= T->cfg.nsf_s[ch]) continue;
        float g4 = 0.017716950f * mb_log(T, S->x34max[ch][w][i]) + (88.411238f - 100.0f + 8.0f);
        float d = (1.00f / 110.5f) * (1800 - (2 * 8) * i - (S->noise0[ch][w][i] - S->nt[ch][w][i]));
        float g = g4 + d;
        int v = round_away(g);
        v = imin_(v, S->gzero[ch][w][i]);
        v = imax_(v, S->gmin[ch][w][i]);
        S->gsf[ch][w][i] = v;
    }
    HMP3_SYNC();
}

HMP3_FN void short_seek_actual(const EncTables *T, ShortRate *S, const float *xr) {  // bitallos.cpp:841-886
    HMP3_SYNC();
    HMP3_SHORT_ITEMS(it) {
        const int ch = it / 48, w = (it >> 4) % 3, i = it & 15;
        if (ch >= S->nchan || i >= T->cfg.nsf_s[ch]) continue;
        const float *y34 = S->x34[ch][w] + T->startBand_s[i];
        const float *y = xr + 576 * ch + 192 * w + T->startBand_s[i];
        const int target = S->nt[ch][w][i];
        const int n = T->nBand_s[i];
        int s = S->gsf[ch][w][i];
        if (S->noise0[ch][w][i] > target) {
            const int logn = T->log_cbw_s[i];
            int noise = band_noise_seq(T, y34, y, s, n, logn);
            int dn = noise - target;
            if (dn > 100) s = seek_finer(T, y34, y, s, n, logn, target, dn, &noise);
            else if (dn < -100) s = seek_coarser(T, y34, y, s, n, logn, target, dn, &noise);
            S->gsf[ch][w][i] = s;
            S->noise[ch][w][i] = noise;
        } else {
            S->gsf[ch][w][i] = S->gzero[ch][w][i] + 5;
            S->noise[ch][w][i] = S->noise0[ch][w][i];
        }
    }
    HMP3_SYNC();
}

HMP3_FN void short_quantise(const EncTables *T, ShortRate *S, bool tuned) {  // bitallos.cpp:889-936
    HMP3_SYNC();
    HMP3_SHORT_ITEMS(it) {
        const int ch = it / 48, w = (it >> 4) % 3, i = it & 15;
        if (ch >= S->nchan || i >= T->cfg.nsf_s[ch]) continue;
        const float *x = S->x34[ch][w] + T->startBand_s[i];
        QLine *q = S->ix[ch][w] + T->startBand_s[i];
        const int n = T->nBand_s[i];
        S->ixmax[ch][w][i] = tuned ? quant_tuned(T, x, q, S->gsf[ch][w][i], n, true, -.30f)
                                   : quant_plain(T, x, q, S->gsf[ch][w][i], n);
    }
    HMP3_SYNC();
}

// per-window gains, scale factors on the coded grid, steps recomputed (bitallos.cpp:1195-1317)
HMP3_FN void short_scale_factors(const EncTables *T, ShortRate *S) {
    for (int ch = 0; ch < S->nchan; ch++) {
        const int nsf = T->cfg.nsf_s[ch];
        S->sf_scale[ch] = 0;
        for (int w = 0; w < 3; w++) {
            int gtop = -1;
            for (int i = 0; i < nsf; i++) {
                S->gsf[ch][w][i] = imax_(S->gsf[ch][w][i], S->gmin[ch][w][i]);
                S->active[ch][w][i] = 0;
                if (S->gsf[ch][w][i] < S->gzero[ch][w][i]) {
                    S->active[ch][w][i] = -1;
                    gtop = imax_(gtop, S->gsf[ch][w][i]);
                }
            }
            S->G[ch][w] = gtop;
        }
        S->GG[ch] = imax_(imax_(S->G[ch][0], S->G[ch][1]), S->G[ch][2]);
        for (int w = 0; w < 3; w++) {
            int gtop = S->G[ch][w];
            if (gtop < 0) {
                S->subgain[ch][w] = 0;
                for (int i = 0; i < nsf; i++) {
                    S->sf[ch][w][i] = 0;
                    S->gsf[ch][w][i] = S->gzero[ch][w][i];
                }
            } else {
                S->subgain[ch][w] = imin_((S->GG[ch] - gtop) & (~7), 7 * 8);
                gtop = S->GG[ch] - S->subgain[ch][w];
                S->G[ch][w] = gtop;
                for (int i = 0; i < nsf; i++) {
                    S->sf[ch][w][i] = 0;
                    if (S->active[ch][w][i]) S->sf[ch][w][i] = gtop - S->gsf[ch][w][i];
                }
            }
        }
        {
            int sp0 = 0;
            for (int i = 0; i < nsf; i++)
                for (int w = 0; w < 3; w++)
                    if (S->active[ch][w][i]) sp0 |= (((i < 6) ? 31 : 15) - S->sf[ch][w][i]);
            S->sf_scale[ch] = (sp0 >= 0) ? 0 : 1;
        }
        if (S->sf_scale[ch] == 0) {
            for (int w = 0; w < 3; w++) {
                if (S->G[ch][w] < 0) continue;
                for (int i = 0; i < nsf; i++) {
                    if ((S->noise[ch][w][i] > S->nt[ch][w][i])) S->sf[ch][w][i]++;
                    S->sf[ch][w][i] = imin_(S->G[ch][w], S->sf[ch][w][i]);
                    S->sf[ch][w][i] &= (~1);
                }
            }
        } else {
            for (int w = 0; w < 3; w++) {
                if (S->G[ch][w] < 0) continue;
                for (int i = 0; i < nsf; i++) {
                    int s = S->sf[ch][w][i] & (~3);
                    int d = S->sf[ch][w][i] - s;
                    int dN = S->noise[ch][w][i] - S->nt[ch][w][i] + 150 * d;
                    if ((dN > 250)) {
                        s = s + 4;
                        s = imin_(S->G[ch][w], s) & (~3);
                    }
                    S->sf[ch][w][i] = s;
                }
            }
        }
        for (int w = 0; w < 3; w++)
            if (S->G[ch][w] >= 0)
                for (int i = 0; i < nsf; i++) {
                    const int hi = ((i < 6) ? 30 : 14) << S->sf_scale[ch];
                    if (S->sf[ch][w][i] > hi) S->sf[ch][w][i] = hi;
                    else if (S->sf[ch][w][i] < 0) S->sf[ch][w][i] = 0;
                }
        for (int w = 0; w < 3; w++) {
            if (S->G[ch][w] < 0) continue;
            for (int i = 0; i < nsf; i++)
                if (S->active[ch][w][i]) {
                    S->gsf[ch][w][i] = S->G[ch][w] - S->sf[ch][w][i];
                    if (S->gsf[ch][w][i] >= S->gzero[ch][w][i]) {
                        S->gsf[ch][w][i] = S->gzero[ch][w][i];
                        S->sf[ch][w][i] = 0;
                    }
                }
        }
    }
}

// fixed regions: region 0 = first three short bands (of all windows), one big region after it, count1
// region = whole bands in transmission order padded to quads (bitallosc.cpp:296-420)
HMP3_FN int short_plan_regions(const EncTables *T, ShortRate *S, int ch) {
    const int ncb = T->cfg.nsf_s[ch];
    const int(*ixmax)[16] = S->ixmax[ch];
    const int *start = T->startBand_s;
    RegionPlan *P = &S->plan[ch];
    int i;
    int cb0 = 3, cb1, cb2;
    for (i = ncb - 1; i >= 0; i--)
        if (ixmax[0][i] > 0 || ixmax[1][i] > 0 || ixmax[2][i] > 0) break;
    cb2 = i + 1;
    for (; i >= 0; i--)
        if (ixmax[0][i] > 1 || ixmax[1][i] > 1 || ixmax[2][i] > 1) break;
    cb1 = i + 1;
    cb1 = imax_(cb1, 3);
    cb2 = imax_(cb2, cb1);
    const int nbig = start[cb1];
    int rmax0 = 0, rmax1 = 0;
    for (i = 0; i < cb0; i++) rmax0 = imax_(rmax0, imax_(ixmax[0][i], imax_(ixmax[1][i], ixmax[2][i])));
    for (; i < cb1; i++) rmax1 = imax_(rmax1, imax_(ixmax[0][i], imax_(ixmax[1][i], ixmax[2][i])));
    const int c0 = count_class_of(T, rmax0), c1 = count_class_of(T, rmax1);
    const int n0 = start[cb0];
    // pairs never straddle windows or bands (band widths are even), so the three windows are simply summed
    int bits = 0;
    {
        // candidates are compared on the sum over the three windows (cnts.c:85-305)
        unsigned s0 = 0, s1 = 0;
        CountResult r;
        for (int reg = 0; reg < 2; reg++) {
            const int c = reg ? c1 : c0;
            const int a = reg ? n0 : 0, b = reg ? nbig : n0;
            const int nc = T->cnt_ncand[c];
            r.bits = r.index = 0;
            if (nc != 0 && b - a > 0) {
                const uint32_t(*lut)[2] = T->cnt_lut[c];
                s0 = s1 = 0;
                for (int w = 0; w < 3; w++) {
                    const QLine *q = S->ix[ch][w];
                    for (int k = a; k < b; k += 2) {
                        int u = q[k], v = q[k + 1];
                        if (c >= 7) { u = u > 15 ? 15 : u; v = v > 15 ? 15 : v; }
                        else { u &= 15; v &= 15; }
                        const uint32_t *e = lut[u * 16 + v];
                        s0 += e[0];
                        s1 += e[1];
                    }
                }
                int b0 = (int)(s0 & 0xFFFF), b1 = (int)((s0 >> 16) & 0xFFFF);
                if (b0 < b1) { r.bits = b0; r.index = 0; }
                else { r.bits = b1; r.index = 1; }
                if (nc == 4) {
                    b0 = (int)(s1 & 0xFFFF);
                    b1 = (int)((s1 >> 16) & 0xFFFF);
                    if (b0 <= r.bits) { r.bits = b0; r.index = 2; }
                    if (b1 <= r.bits) { r.bits = b1; r.index = 3; }
                }
            }
            bits += r.bits;
            P->table[reg] = T->cnt_tables[c][r.index];
        }
    }
    P->table[2] = 0;
    int k = 0;
    QLine *qs = S->quad_scratch;
    for (i = cb1; i < cb2; i++)
        for (int w = 0; w < 3; w++)
            for (int j = start[i]; j < start[i + 1]; j++) qs[k++] = S->ix[ch][w][j];
    qs[k] = qs[k + 1] = qs[k + 2] = 0;
    k = (k + 3) & (~3);
    const int nquads = k >> 2;
    CountResult r = count_quads(qs, nquads);
    bits += r.bits;
    P->table[3] = r.index;
    P->cb[0] = cb0;
    P->cb[1] = cb1;
    P->cb[2] = cb2;
    P->nbig = nbig;
    P->nquads = nquads;
    P->bits = bits;
    return bits;
}
HMP3_FN int short_count(const EncTables *T, ShortRate *S) {
    int bits = 0;
    for (int ch = 0; ch < S->nchan; ch++) {
        S->huff_bits[ch] = short_plan_regions(T, S, ch);
        bits += S->huff_bits[ch];
    }
    return bits;
}
HMP3_FN void short_plan_to_side(const EncTables *T, const RegionPlan *P, GrSide *g) {  // bitallosc.cpp:423-487
    if (P->bits <= 0) {
        g->table_select[0] = g->table_select[1] = g->table_select[2] = 0;
        g->big_values = 0;
        g->region0_count = g->region1_count = 0;
        g->aux_nreg[0] = g->aux_nreg[1] = g->aux_nreg[2] = 0;
        g->aux_nquads = 0;
        g->count1table_select = 0;
        return;
    }
    g->table_select[0] = P->table[0];
    g->table_select[1] = P->table[1];
    g->table_select[2] = P->table[2];
    g->count1table_select = P->table[3];
    g->big_values = 3 * (P->nbig >> 1);
    g->region0_count = 0;
    g->region1_count = 0;
    int n0 = T->startBand_s[P->cb[0]], n1 = T->startBand_s[P->cb[1]], n2 = T->startBand_s[P->cb[2]];
    if (n2 > P->nbig) n2 = P->nbig;
    if (n1 > n2) n1 = n2;
    if (n0 > n1) n0 = n1;
    n2 = n2 - n1;
    n1 = n1 - n0;
    g->aux_nreg[0] = 3 * (n0 >> 1);
    g->aux_nreg[1] = 3 * (n1 >> 1);
    g->aux_nreg[2] = 3 * (n2 >> 1);
    g->aux_nquads = P->nquads;
}

// one short granule (bitallos.cpp:1455-1503 with the control loops :1320-1452)
HMP3_FN void short_allocate(const EncTables *T, ShortRate *S, const float *xr) {
    if (S->mnr < -200) S->min_target = imax_(S->min_target, (3 * S->target) >> 2);
    short_seek_initial(T, S);
    short_seek_actual(T, S, xr);
    short_scale_factors(T, S);
    short_quantise(T, S, true);
    int bits = short_count(T, S);
    S->feedback_bits = bits;
    if (bits < S->min_target) {
        for (int k = 0; k < 10; k++) {
            for (int ch = 0; ch < S->nchan; ch++)
                for (int w = 0; w < 3; w++)
                    for (int i = 0; i < T->cfg.nsf_s[ch]; i++) S->gsf[ch][w][i] = imax_(S->gsf[ch][w][i] - 1, 0);
            short_scale_factors(T, S);
            short_quantise(T, S, true);
            bits = short_count(T, S);
            if (bits >= S->min_target) break;
        }
    }
    if (bits > S->max_target) {
        const int f = (250 * 1024) / (S->active_lines + 10);
        int dN = imax_((f * (bits - S->max_target)) >> 10, 40);
        S->delta_mnr = 0;
        for (int k = 0; k < 10; k++) {
            S->delta_mnr += dN;
            for (int ch = 0; ch < S->nchan; ch++)
                for (int w = 0; w < 3; w++)
                    for (int i = 0; i < T->cfg.nsf_s[ch]; i++) S->nt[ch][w][i] += dN;
            short_seek_actual(T, S, xr);
            short_scale_factors(T, S);
            short_quantise(T, S, false);
            bits = short_count(T, S);
            if (bits <= S->max_target) break;
            dN = imax_((f * (bits - S->max_target)) >> 10, 40);
        }
    }
    if (bits > S->max_bits) {
        for (int k = 0; k < 100; k++) {
            for (int ch = 0; ch < S->nchan; ch++)
                for (int w = 0; w < 3; w++)
                    for (int i = 0; i < T->cfg.nsf_s[ch]; i++) S->gsf[ch][w][i] = imin_(127, S->gsf[ch][w][i] + 1);
            short_scale_factors(T, S);
            short_quantise(T, S, false);
            bits = short_count(T, S);
            if (bits <= S->max_bits) break;
        }
    }
    if (bits > kPart23Max) {
        bool over = false;
        for (int ch = 0; ch < S->nchan; ch++)
            if (S->huff_bits[ch] > kPart23Max) over = true;
        if (over)
            for (int k = 0; k < 100; k++) {
                for (int ch = 0; ch < S->nchan; ch++)
                    if (S->huff_bits[ch] > kPart23Max)
                        for (int w = 0; w < 3; w++)
                            for (int i = 0; i < T->cfg.nsf_s[ch]; i++)
                                S->gsf[ch][w][i] = imin_(127, S->gsf[ch][w][i] + 1);
                short_scale_factors(T, S);
                short_quantise(T, S, false);
                bits = short_count(T, S);
                if ((S->huff_bits[0] <= kPart23Max) && (S->huff_bits[1] <= kPart23Max)) break;
            }
    }
}

// Full short-block granule: returns the feedback bit count.  gr/sf_out are the persistent side-info
// records of this granule; ix_out/sign_out receive the lines in transmission order
// (bitallos.cpp:202-372).
HMP3_FN int short_granule(const EncTables *T, ShortRate *S, float *xr, const SigMask *sm, int nchan, int min_bits,
                          int target_bits, int max_bits, int pool_bits, ScaleFac *sf_out, GrSide *gr, QLine *ix_out,
                          unsigned *sign_out /*[2][18]*/, int ms, int mnr) {
    S->mnr = mnr;
    if (T->cfg.h_id == 0) S->mnr = imin_(S->mnr, 850);
    S->ms = ms;
    S->nchan = nchan;
    S->max_bits = imin_(4000 * nchan, max_bits);
    S->min_target = min_bits < 0 ? 0 : min_bits;
    S->target = target_bits;
    S->pool_bits = pool_bits;
    S->max_target = S->target + ((614 * S->pool_bits) >> 10);
    S->max_target = (S->max_bits + S->max_target) >> 1;
    S->max_target = imin_(S->max_bits, S->max_target);
    if (ms) short_startup_ms(T, S, xr, sm);
    else short_startup_lr(T, S, xr, sm);
    if (S->active_lines <= 0) {
        for (int ch = 0; ch < nchan; ch++) {
            GrSide *g = gr + ch;
            g->global_gain = 0;
            g->window_switching_flag = 1;
            g->block_type = 2;
            g->mixed_block_flag = 0;
            g->preflag = 0;
            g->scalefac_scale = 0;
            g->table_select[0] = g->table_select[1] = g->table_select[2] = 0;
            g->subblock_gain[0] = g->subblock_gain[1] = g->subblock_gain[2] = 0;
            g->big_values = 0;
            g->region0_count = g->region1_count = 0;
            g->count1table_select = 0;
            g->aux_nquads = 0;
            g->aux_bits = 0;
            g->aux_not_null = 0;
            g->aux_nreg[0] = g->aux_nreg[1] = g->aux_nreg[2] = 0;
            for (int w = 0; w < 3; w++)
                for (int j = 0; j < 12; j++) sf_out[ch].s[w][j] = 0;
        }
        S->feedback_bits = 0;
        return 0;
    }
    short_allocate(T, S, xr);
    if (ms) {
        S->GG[0] -= 2;
        S->GG[1] -= 2;
    }
    S->GG[0] = imax_(S->GG[0], 0);
    S->GG[1] = imax_(S->GG[1], 0);
    for (int ch = 0; ch < nchan; ch++) {
        GrSide *g = gr + ch;
        g->global_gain = imin_(S->GG[ch] + (4 * 32 + 14), 255);
        g->window_switching_flag = 1;
        g->block_type = 2;
        g->mixed_block_flag = 0;
        g->preflag = 0;
        g->scalefac_scale = S->sf_scale[ch];
        g->aux_bits = S->huff_bits[ch];
        g->aux_not_null = S->huff_bits[ch];
        g->subblock_gain[0] = S->subgain[ch][0] >> 3;
        g->subblock_gain[1] = S->subgain[ch][1] >> 3;
        g->subblock_gain[2] = S->subgain[ch][2] >> 3;
        short_plan_to_side(T, &S->plan[ch], g);
    }
    // scale factors on the coded grid (bitallos.cpp:418-452)
    for (int ch = 0; ch < nchan; ch++) {
        const int sh = S->sf_scale[ch] == 0 ? 1 : 2;
        for (int w = 0; w < 3; w++) {
            for (int i = 0; i < T->cfg.nsf_s[ch]; i++) S->sf[ch][w][i] >>= sh;
            for (int i = 0; i < 12; i++) sf_out[ch].s[w][i] = S->sf[ch][w][i];
        }
    }
    // lines in transmission order: [band][window][line] (bitallos.cpp:329-366); the signs of the lines written are
    // merged into the persistent sign words
    for (int ch = 0; ch < nchan; ch++) {
        QLine *dst = ix_out + 576 * ch;
        unsigned *ds = sign_out + 18 * ch;
        const int ncb = S->plan[ch].cb[2];
        const int total = 3 * T->startBand_s[ncb];
#if HMP3_COOP
        const int lane = HMP3_LANE;
        HMP3_SYNC();
        for (int k = lane; k < 576; k += HMP3_W) dst[k] = 0;
        HMP3_SYNC();
        for (int it = lane; it < 3 * ncb; it += HMP3_W) {  // one (band, window) per lane
            const int i = it / 3, w = it - 3 * i;
            const int n = T->nBand_s[i], s0 = T->startBand_s[i], base = 3 * s0 + w * n;
            for (int j = 0; j < n; j++) {
                dst[base + j] = S->ix[ch][w][s0 + j];
                S->sign_t[ch][base + j] = S->sign[ch][w][s0 + j];
            }
        }
        HMP3_SYNC();
        for (int w = 0; 32 * w < total; w++) {
            unsigned bits = 0;
            for (int h = 0; h < 32 / HMP3_W; h++) {
                const int k = 32 * w + HMP3_W * h + lane;
                bits |= gballot(k < total ? (S->sign_t[ch][k] & 1) : 0) << (HMP3_W * h);
            }
            if (lane == 0) merge_sign_word(ds + w, bits, w, total);
        }
        HMP3_SYNC();
#else
        for (int k = 0; k < 576; k++) dst[k] = 0;
        int k = 0;
        for (int i = 0; i < ncb; i++)
            for (int w = 0; w < 3; w++)
                for (int j = T->startBand_s[i]; j < T->startBand_s[i + 1]; j++) {
                    dst[k] = S->ix[ch][w][j];
                    S->sign_t[ch][k] = S->sign[ch][w][j];
                    k++;
                }
        for (int w = 0; 32 * w < total; w++) {
            unsigned bits = 0;
            for (int b = 0; b < 32 && 32 * w + b < total; b++) bits |= (unsigned)(S->sign_t[ch][32 * w + b] & 1) << b;
            merge_sign_word(ds + w, bits, w, total);
        }
#endif
    }
    return S->feedback_bits;
}

}  // namespace hmp3
