// Shared pieces of the rate loop: side-info records, quantiser passes, noise measurement and the
// Huffman region planner / bit counter.  Integer results are bit-exact with the reference; float
// expressions keep the reference's evaluation order (compile with --fmad=false / -ffp-contract=off).
#pragma once
#include <math.h>
#include "enc_tables.h"

namespace hmp3 {

// Granule side information (same fields, order and persistence as GR, pub/l3e.h:71-96: fields that a
// block type does not write keep their previous value, e.g. subblock_gain on start/stop blocks).
struct GrSide {
    int part2_3_length, big_values, global_gain, scalefac_compress, window_switching_flag, block_type,
        mixed_block_flag, table_select[3], subblock_gain[3], region0_count, region1_count, preflag, scalefac_scale,
        count1table_select, aux_nquads, aux_bits, aux_not_null, aux_nreg[3], block_type_prev, short_flag_current,
        short_flag_next;
};
struct ScaleFac {  // pub/l3e.h:114-120
    int l[23];
    int s[3][13];
};

// Result of one Huffman region planning pass (the reference keeps it in the file-static save[ch],
// bitalloc.cpp:78-86 / bitallos.h:224-232).
struct RegionPlan {
    int table[4];  // big-value tables of regions 0..2, count1 table select
    int cb[3];     // region ends in scale-factor bands
    int nbig, nquads, bits;
};

// A quantised spectral magnitude.  The step range (gmin = gzero - 70) bounds them by 8191, so 16 bits hold them:
// half the state bytes, and a pair (the unit of the Huffman big-value tables) is one 32-bit word.
typedef short QLine;

HMP3_HD int imin_(int a, int b) { return a < b ? a : b; }
HMP3_HD int imax_(int a, int b) { return a > b ? a : b; }

// (hot sequential loops are unrolled by hand: see sum_seq in enc_tables.h)
HMP3_HD int max_seq(const QLine *v, int n) {  // max(0, v[0], ..., v[n-1])
    int m = 0, k = 0;
    for (; k < n; k++) m = imax_(m, v[k]);
    return m;
}

// ------------------------------------------------------------------ quantiser passes (l3math.c)
// plain rounding quantiser (l3math.c:655-671)
HMP3_FN int quant_plain(const EncTables *T, const float *x34, QLine *ix, int g, int n) {
    const float ig = T->igain34[g];
    int m = 0;
    for (int i = 0; i < n; i++) {
        int q = (int)(ig * x34[i] + (0.5f - 0.0946f));
        ix[i] = (QLine)q;
        if (q > m) m = q;
    }
    return m;
}
// RD-tuned quantiser: magnitude-dependent rounding offset (l3math.c:674-694); `r0` replaces the offset of
// magnitude class 0 (l3math.c:697-725 when given), clamp_lo mirrors the extra lower clamp of that variant.
HMP3_FN int quant_tuned(const EncTables *T, const float *x34, QLine *ix, int g, int n, bool alt, float r0) {
    const float ig = T->igain34[g];
    int m = 0;
    for (int i = 0; i < n; i++) {
        float t = ig * x34[i] + (0.5f - 0.4375f);
        int c = (int)t;
        if (c > 31) c = 31;
        if (alt && c < 0) c = 0;
        float off = (alt && c == 0) ? r0 : T->quantB_round[c];
        int q = (int)(t - off);
        ix[i] = (QLine)q;
        if (q > m) m = q;
    }
    return m;
}
// quantised value of a band maximum, actual and 10x scale (l3math.c:746-780)
HMP3_HD int quant_tuned_peak(const EncTables *T, float x34max, int g) {
    float t = T->igain34[g] * x34max + (0.5f - 0.4375f);
    int c = (int)t;
    if (c > 31) c = 31;
    return (int)(t - T->quantB_round[c]);
}
HMP3_HD int quant_tuned_peak10(const EncTables *T, float x34max, int g) {
    float t = T->igain34[g] * x34max + (0.5f - 0.4375f);
    int c = (int)t;
    if (c > 31) c = 31;
    return (int)(10.0f * (t - T->quantB_round[c]) + (0.5f - 5.0f));
}

// x^(4/3) of a quantised value (table below 256, libm above: l3math.c:527-534)
HMP3_HD float dequant43(const EncTables *T, int q) {
    if (q >= 0 && q < 256) return T->ix43[q];
    return (float)pow((double)q, (4.0 / 3.0));
}

// quantisation noise of one band at step g, in millibels relative to the band width (l3math.c:511-544)
// the plain sequential loop (host build; on the device also used one band per lane where bands are short)
HMP3_HD float noise_line(const EncTables *T, float ig, float gn, float x34, float x) {  // squared error of one line
    float t = (ig * x34 + (0.0f - 0.0946f));
    int q = (int)(t + ((f2u(t) >> 31) ? -0.5f : 0.5f));
    float xh;
    if (q >= 0 && q < 256) xh = gn * T->ix43[q];
    else xh = (float)((double)gn * pow((double)q, (4.0 / 3.0)));
    float d = x - xh;
    return d * d;
}
HMP3_HD int band_noise_seq(const EncTables *T, const float *x34, const float *x, int g, int n, int logn) {
    const float ig = T->igain34[g], gn = T->gain[g];
    float acc = 0.0f;
    int i = 0;
    for (; i < n; i++) acc += noise_line(T, ig, gn, x34[i], x[i]);
    return mb_log(T, 1.0e-12f + acc) - logn;
}
HMP3_HD float dequant43_sq(const EncTables *T, int q) {  // (q^(4/3))^2 as band_refit_gain takes it
    float v;
    if (q < 256) v = T->ix43[q];
    else v = (float)(pow((double)q, (4.0 / 3.0)));
    return v * v;
}
HMP3_HD int band_refit_gain_seq(const EncTables *T, const QLine *q, const float *x, int n) {
    float sqq = 0, sxx = 0;
    int i = 0;
    for (; i < n; i++) {
        sqq += dequant43_sq(T, q[i]);
        sxx += x[i] * x[i];
    }
    return 54 * mb_log(T, sxx / sqq) + (8 << 13);
}

HMP3_FN int band_noise(const EncTables *T, const float *x34, const float *x, int g, int n, int logn) {
    const float ig = T->igain34[g], gn = T->gain[g];
    float acc = 0.0f;
#if HMP3_COOP
    // lanes square the errors of 32 lines at a time; the sum is then taken in line order by every lane
    const int lane = HMP3_LANE;
    for (int i0 = 0; i0 < n; i0 += HMP3_W) {
        const int i = i0 + lane;
        float dd = 0.0f;
        if (i < n) {
            float t = (ig * x34[i] + (0.0f - 0.0946f));
            int q = (int)(t + ((f2u(t) >> 31) ? -0.5f : 0.5f));
            float xh;
            if (q >= 0 && q < 256) xh = gn * T->ix43[q];
            else xh = (float)((double)gn * pow((double)q, (4.0 / 3.0)));
            float d = x[i] - xh;
            dd = d * d;
        }
        acc = gsum_ordered(acc, dd, (n - i0) < HMP3_W ? (n - i0) : HMP3_W);
    }
#else
    for (int i = 0; i < n; i++) {
        float t = (ig * x34[i] + (0.0f - 0.0946f));
        int q = (int)(t + ((f2u(t) >> 31) ? -0.5f : 0.5f));
        float xh;
        if (q >= 0 && q < 256) xh = gn * T->ix43[q];
        else xh = (float)((double)gn * pow((double)q, (4.0 / 3.0)));
        float d = x[i] - xh;
        acc += d * d;
    }
#endif
    return mb_log(T, 1.0e-12f + acc) - logn;
}

// gain (scaled by 2^13) that best maps quantised values back onto the spectrum (l3math.c:1087-1114)
HMP3_FN int band_refit_gain(const EncTables *T, const QLine *q, const float *x, int n) {
    float sqq = 0, sxx = 0;
#if HMP3_COOP
    const int lane = HMP3_LANE;
    for (int i0 = 0; i0 < n; i0 += HMP3_W) {
        const int i = i0 + lane;
        float vv = 0.0f, xx = 0.0f;
        if (i < n) {
            float v;
            if (q[i] < 256) v = T->ix43[q[i]];
            else v = (float)(pow((double)q[i], (4.0 / 3.0)));
            vv = v * v;
            xx = x[i] * x[i];
        }
        const int m = (n - i0) < HMP3_W ? (n - i0) : HMP3_W;
        sqq = gsum_ordered(sqq, vv, m);
        sxx = gsum_ordered(sxx, xx, m);
    }
#else
    for (int i = 0; i < n; i++) {
        float v;
        if (q[i] < 256) v = T->ix43[q[i]];
        else v = (float)(pow((double)q[i], (4.0 / 3.0)));
        sqq += v * v;
        sxx += x[i] * x[i];
    }
#endif
    return 54 * mb_log(T, sxx / sqq) + (8 << 13);
}

// ------------------------------------------------------------------ Huffman bit counting
struct CountResult { int bits, index; };

HMP3_HD int count_class_of(const EncTables *T, int m) {
    if (m <= 22) return T->cnt_class_of_max[m];
    if (m <= 30) return 10;
    if (m <= 46) return 11;
    if (m <= 78) return 12;
    if (m <= 142) return 13;
    if (m <= 270) return 14;
    if (m <= 526) return 15;
    if (m <= 1038) return 16;
    if (m <= 2062) return 17;
    return 18;
}

// bits of n values (n/2 pairs) under the candidate tables of class c; ties go to the higher candidate
// (cnt.c:96-292)
// (device: ix + even offsets are 4-byte aligned -- band edges and region ends are even -- so a pair is one load)
HMP3_HD void load_pair(const QLine *p, int *a, int *b) {
#if HMP3_COOP
    const unsigned w = *(const unsigned *)p;
    *a = (int)(short)(w & 0xffffu);
    *b = (int)(short)(w >> 16);
#else
    *a = p[0];
    *b = p[1];
#endif
}
HMP3_FN CountResult count_pairs(const EncTables *T, int c, const QLine *ix, int n) {
    CountResult r;
    r.bits = r.index = 0;
    const int nc = T->cnt_ncand[c];
    if (nc == 0 || n <= 0) return r;
    const uint32_t(*lut)[2] = T->cnt_lut[c];
    unsigned s0 = 0, s1 = 0;
#if HMP3_COOP
    const int i_first = 2 * HMP3_LANE, i_step = 2 * HMP3_W;
#else
    const int i_first = 0, i_step = 2;
#endif
    if (c >= 7) {  // escape tables: values above 15 use the row/column of 15
        for (int i = i_first; i < n; i += i_step) {
            int a, b;
            load_pair(ix + i, &a, &b);
            a = a > 15 ? 15 : a;
            b = b > 15 ? 15 : b;
            s0 += lut[a * 16 + b][0];
        }
    } else if (nc == 2) {
        for (int i = i_first; i < n; i += i_step) {
            int a, b;
            load_pair(ix + i, &a, &b);
            s0 += lut[(a & 15) * 16 + (b & 15)][0];
        }
    } else {
        for (int i = i_first; i < n; i += i_step) {
            int a, b;
            load_pair(ix + i, &a, &b);
            const uint32_t *e = lut[(a & 15) * 16 + (b & 15)];
            s0 += e[0];
            s1 += e[1];
        }
    }
#if HMP3_COOP
    s0 = gsum(s0);
    s1 = gsum(s1);
#endif
    int b0 = (int)(s0 & 0xFFFF), b1 = (int)((s0 >> 16) & 0xFFFF);
    if (b0 < b1) { r.bits = b0; r.index = 0; }
    else { r.bits = b1; r.index = 1; }
    if (nc == 4) {
        b0 = (int)(s1 & 0xFFFF);
        b1 = (int)((s1 >> 16) & 0xFFFF);
        if (b0 <= r.bits) { r.bits = b0; r.index = 2; }
        if (b1 <= r.bits) { r.bits = b1; r.index = 3; }
    }
    return r;
}

// count1 region: table A (variable length) against table B (4 bits), sign bits included (cnt.c:295-326)
HMP3_FN CountResult count_quads(const QLine *ix, int nquads) {
    CountResult r;
    r.bits = r.index = 0;
    if (nquads <= 0) return r;
    int a = 0, b = 0;
#if HMP3_COOP
    for (int i = HMP3_LANE; i < nquads; i += HMP3_W) {
        const int k = 4 * i;
#else
    for (int i = 0, k = 0; i < nquads; i++, k += 4) {
#endif
        int q0, q1, q2, q3;
        load_pair(ix + k, &q0, &q1);
        load_pair(ix + k + 2, &q2, &q3);
        int j = ((q0 << 3) + (q1 << 2) + (q2 << 1) + q3) & 15;
        int ones = (j & 1) + ((j >> 1) & 1) + ((j >> 2) & 1) + ((j >> 3) & 1);
        a += kQuadLenA[j] + ones;
        b += 4 + ones;
    }
#if HMP3_COOP
    a = gsum(a);
    b = gsum(b);
#endif
    if (a < b) { r.bits = a; r.index = 0; }
    else { r.bits = b; r.index = 1; }
    return r;
}

// region lengths in bands as a function of the number of big-value bands (bitalloc.cpp:124-203)
HMP3_HD void region_split_rule(int nbands, int *r0, int *r1) {
    *r0 = kRegion0[nbands];
    *r1 = kRegion1[nbands];
}

// Region planning + bit count for a long-block granule channel (block types 0 / 1,3).
// ixmax[] = per-band maxima, ix[] = the persistent quantised-line buffer (lines past the last coded band
// keep whatever earlier granules left there, as in the reference).  bitalloc.cpp:470-754.
HMP3_FN int plan_regions_long(const EncTables *T, int block_type, const int *ixmax, const QLine *ix, int ncb,
                              RegionPlan *P) {
    const int *start = T->startBand_l, *width = T->nBand_l;
    int i, j, n;
    int cb[4], rmax[3];
    const bool fixed = (block_type != 0);  // window switching: region0 = 8 bands, one big region after it
    for (i = ncb - 1; i >= 0; i--)
        if (ixmax[i] > 0) break;
    cb[3] = i + 1;
    for (; i >= 0; i--)
        if (ixmax[i] > 1) break;
    cb[2] = i + 1;
    const int keep = fixed ? 8 : 2;
    if (fixed) {
        cb[0] = 8;
        cb[2] = imax_(cb[2], keep);
        cb[3] = imax_(cb[3], cb[2]);
        cb[1] = cb[0];
    } else if (cb[2] < 2) {
        cb[2] = 2;
        if (cb[3] < cb[2]) cb[3] = cb[2];
    }
    j = start[cb[2]];
    n = width[cb[2] - 1];
    for (i = 0; i < n; i++) {
        j--;
        if (ix[j] > 1) break;
    }
    int nbig = (j + 2) & (~1);
    if (nbig < start[keep]) nbig = start[keep];
    j = start[cb[3]];
    n = width[cb[3] - 1];
    for (i = 0; i < n; i++) {
        j--;
        if (ix[j] > 0) break;
    }
    int nquads = (j + 4 - nbig) >> 2;
    if (fixed) nquads = imax_(nquads, 0);
    if (!fixed) {
        int r0, r1;
        region_split_rule(cb[2], &r0, &r1);
        cb[0] = r0;
        cb[1] = r0 + r1;
        if (cb[0] < 1) cb[0] = 1;
        if (cb[1] <= cb[0]) cb[1] = cb[0] + 1;
        if (cb[1] > cb[0] + 8) cb[1] = cb[0] + 8;
    }
    rmax[0] = rmax[1] = rmax[2] = 0;
    for (i = 0; i < cb[0]; i++) rmax[0] = imax_(rmax[0], ixmax[i]);
    if (!fixed)
        for (; i < cb[1]; i++) rmax[1] = imax_(rmax[1], ixmax[i]);
    for (; i < cb[2]; i++) rmax[2] = imax_(rmax[2], ixmax[i]);
    int cls[3];
    for (i = 0; i < 3; i++) cls[i] = count_class_of(T, rmax[i]);
    if (!fixed) {
        // shrink a region whose table is richer than its upper neighbour's (bitalloc.cpp:563-595)
        if (T->cnt_tmax[cls[2]] < T->cnt_tmax[cls[1]]) {
            for (j = cb[1] - 1; j > cb[0]; j--)
                if (ixmax[j] > T->cnt_tmax[cls[2]]) break;
            cb[1] = j + 1;
        }
        if (T->cnt_tmax[cls[1]] < T->cnt_tmax[cls[0]]) {
            n = cb[1] - 8;
            if (n < 1) n = 1;
            for (j = cb[0] - 1; j > n; j--)
                if (ixmax[j] > T->cnt_tmax[cls[1]]) break;
            cb[0] = j + 1;
        }
    }
    const int n0 = start[cb[0]], n1 = start[cb[1]];
    CountResult r = count_pairs(T, cls[0], ix, n0);
    int bits = r.bits;
    P->table[0] = T->cnt_tables[cls[0]][r.index];
    if (!fixed) {
        r = count_pairs(T, cls[1], ix + n0, n1 - n0);
        bits += r.bits;
        P->table[1] = T->cnt_tables[cls[1]][r.index];
    }
    r = count_pairs(T, cls[2], ix + n1, nbig - n1);
    bits += r.bits;
    P->table[2] = T->cnt_tables[cls[2]][r.index];
    if (fixed) P->table[1] = P->table[2];
    r = count_quads(ix + nbig, nquads);
    bits += r.bits;
    P->table[3] = r.index;
    P->cb[0] = cb[0];
    P->cb[1] = cb[1];
    P->cb[2] = cb[2];
    P->nbig = nbig;
    P->nquads = nquads;
    P->bits = bits;
    return bits;
}

// side-info fields that follow from a region plan (bitalloc.cpp:758-811)
HMP3_FN void plan_to_side(const EncTables *T, const RegionPlan *P, GrSide *g) {
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
    g->big_values = P->nbig >> 1;
    g->region0_count = P->cb[0] - 1;
    g->region1_count = imax_((P->cb[1] - P->cb[0]) - 1, 0);
    int n0 = T->startBand_l[P->cb[0]], n1 = T->startBand_l[P->cb[1]], n2 = T->startBand_l[P->cb[2]];
    if (n2 > P->nbig) n2 = P->nbig;
    if (n1 > n2) n1 = n2;
    if (n0 > n1) n0 = n1;
    n2 = n2 - n1;
    n1 = n1 - n0;
    g->aux_nreg[0] = n0 >> 1;
    g->aux_nreg[1] = n1 >> 1;
    g->aux_nreg[2] = n2 >> 1;
    g->aux_nquads = P->nquads;
}

}  // namespace hmp3
