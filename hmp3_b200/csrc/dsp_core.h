// Analysis front end of the Layer III path: polyphase filterbank, attack energies, hybrid window +
// 18/6-point MDCT, alias reduction.  Every function is a pure per-unit routine (one time slot, one
// sub-band, one boundary) so that the CUDA kernels map one thread to one unit; the same bodies are
// compiled for the host by the test-only simulator.  Floating-point evaluation order follows the
// reference exactly (bit-exact parity; compile with --fmad=false / -ffp-contract=off).
#pragma once
#include "enc_tables.h"

namespace hmp3 {

// ------------------------------------------------------------------------------------------------
// Polyphase analysis, one time slot: 512-sample window (newest sample first) -> 32 sub-band values.
// Reference: window() sbt.c:57-108 and fidct_L3() sbt.c:134-257.
// `fetch(i)` returns window sample i (i = 0 newest .. 511 oldest); out[k*stride] receives band k.
// ------------------------------------------------------------------------------------------------
template <int N, int M>
HMP3_HD void dct32_split(const float *x, float *f) {
    // M blocks of N: even samples to the low half, odd samples to the high half with a running
    // alternating difference taken from the top (sbt.c:134-160)
#pragma unroll
    for (int blk = 0; blk < M; blk++) {
        const int b0 = blk * N;
        constexpr int H = N / 2;
#pragma unroll
        for (int j = 0; j < H; j++) f[b0 + j] = x[b0 + 2 * j];
        f[b0 + N - 1] = x[b0 + N - 1];
#pragma unroll
        for (int j = H - 2; j >= 0; j--) f[b0 + H + j] = x[b0 + 2 * j + 1] - f[b0 + H + j + 1];
    }
}
template <int N, int M>
HMP3_HD void dct32_merge(const float *x, float *f, const float *c) {
    // butterflies with the 2cos twiddles (sbt.c:163-184)
#pragma unroll
    for (int blk = 0; blk < M; blk++) {
        const int b0 = blk * N;
        constexpr int H = N / 2;
#pragma unroll
        for (int j = 0; j < H; j++) {
            float tw = c[j] * x[b0 + j + H];
            float t = x[b0 + j];
            f[b0 + j] = t + tw;
            f[b0 + N - 1 - j] = t - tw;
        }
    }
}

template <class Fetch>
HMP3_HD void polyphase_slot(const EncTables *T, Fetch fetch, float *out, int stride) {
    float a[32], b[32];
    // window fold: two 8-term sums per output, accumulated in table order
    {
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; j++) s += T->polyA[0][j] * fetch(T->polyIa[0] + 64 * j);
        b[0] = s;
    }
#pragma unroll
    for (int k = 1; k < 32; k++) {
        float s1 = 0.0f, s2 = 0.0f;
        const int ia = T->polyIa[k], ib = T->polyIb[k];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            s1 += T->polyA[k][j] * fetch(ia + 64 * j);
            s2 += T->polyB[k][j] * fetch(ib + 64 * j);
        }
        b[k] = s1 + s2;
    }
    const float *c = T->dct32;
    dct32_split<32, 1>(b, a);
    dct32_split<16, 2>(a, b);
    dct32_split<8, 4>(b, a);
    dct32_split<4, 8>(a, b);
    dct32_merge<2, 16>(b, a, c + 16 + 8 + 4 + 2);
    dct32_merge<4, 8>(a, b, c + 16 + 8 + 4);
    dct32_merge<8, 4>(b, a, c + 16 + 8);
    dct32_merge<16, 2>(a, b, c + 16);
#pragma unroll
    for (int k = 0; k < 16; k++) {  // last stage writes band-major (sbt.c:222-236)
        float tw = c[k] * b[k + 16];
        out[stride * k] = b[k] + tw;
        out[stride * (31 - k)] = b[k] - tw;
    }
}

// Frequency inversion applied before the hybrid transform: odd time index of odd sub-bands, for the
// band pairs below `nsb` (hwin.c:282-294).  Returns true if sample (sb,t) is negated.
HMP3_HD bool freq_inverted(int sb, int t, int nsb) { return (sb & 1) && (t & 1) && (sb - 1) < nsb; }

// ------------------------------------------------------------------------------------------------
// Attack detector energies of one granule: 9 millibel energies over pairs of time slots
// (detect.c:53-107, 147-195).  sbt is band-major [32][18].
// ------------------------------------------------------------------------------------------------
HMP3_HD int attack_energy(const EncTables *T, const float *sbt, int k /*0..8*/, int mpeg2) {
    const int start = mpeg2 ? 8 : 4, nb = mpeg2 ? 20 : 14;
    const float *y = sbt + 18 * start + 2 * k;
    float sum = 7.0e4f;
    for (int i = 0; i < nb; i++) {
        float x = y[0] * y[0];
        sum += x;
        x = y[1] * y[1];
        sum += x;
        y += 18;
    }
    return mb_log(T, sum);
}

// Peak of (energy - max of the preceding window) over the newest 12 (or 11) energies (detect.c:109-139,197-227)
// e[] is the 32-entry history with the newest granule at [23..31].
HMP3_HD int attack_measure(const int *e, int prev_short, int mpeg2) {
    int m = 0;
    for (int j = prev_short ? 18 : 17; j < 29; j++) {
        int a1 = e[j - 4] > e[j - 5] ? e[j - 4] : e[j - 5];
        int a2 = e[j - 2] > e[j - 3] ? e[j - 2] : e[j - 3];
        int a = a1 > a2 ? a1 : a2;
        if (!mpeg2) {
            int a0 = e[j - 6] > e[j - 7] ? e[j - 6] : e[j - 7];
            a = a > a0 ? a : a0;
        }
        int d = e[j] - a;
        m = m > d ? m : d;
    }
    return m;
}

// block type from (previous type, current short flag, next short flag) (mp3enc.cpp:82-87)
HMP3_HD int block_type_rule(int prev, int cur, int next) {
    const int tab[16] = {0, 1, 2, 2, 3, 2, 2, 2, 3, 2, 2, 2, 0, 1, 2, 2};
    return tab[prev * 4 + cur * 2 + next];
}

// ------------------------------------------------------------------------------------------------
// 18-point forward MDCT core on a pre-windowed, folded vector f[18] (emdct.c:104-187).
// ------------------------------------------------------------------------------------------------
HMP3_HD void mdct18_core(const EncTables *T, const float *f, float *y) {
    const float *w = T->m18_w, *w2 = T->m18_w2;
    float a[9], b[9];
#pragma unroll
    for (int p = 0; p < 4; p++) {
        float g1 = w[p] * f[p];
        float g2 = w[17 - p] * f[17 - p];
        float ap = g1 + g2;
        float bp = w2[p] * (g1 - g2);
        g1 = w[8 - p] * f[8 - p];
        g2 = w[9 + p] * f[9 + p];
        float a8p = g1 + g2;
        float b8p = w2[8 - p] * (g1 - g2);
        a[p] = ap + a8p;
        a[5 + p] = ap - a8p;
        b[p] = bp + b8p;
        b[5 + p] = bp - b8p;
    }
    {
        float g1 = w[4] * f[4];
        float g2 = w[13] * f[13];
        a[4] = g1 + g2;
        b[4] = w2[4] * (g1 - g2);
    }
    const float(*c)[4] = T->m18_c;
    // even outputs from a[], odd from b[], then the running difference y[k] -= y[k-1]
    float e0 = 0.5f * (a[0] + a[1] + a[2] + a[3] + a[4]);
    float o0 = 0.5f * (b[0] + b[1] + b[2] + b[3] + b[4]);
    float e1 = c[1][0] * a[5] + c[1][1] * a[6] + c[1][2] * a[7] + c[1][3] * a[8];
    float o1 = c[1][0] * b[5] + c[1][1] * b[6] + c[1][2] * b[7] + c[1][3] * b[8] - o0;
    float e2 = c[2][0] * a[0] + c[2][1] * a[1] + c[2][2] * a[2] + c[2][3] * a[3] - a[4];
    float o2 = c[2][0] * b[0] + c[2][1] * b[1] + c[2][2] * b[2] + c[2][3] * b[3] - b[4] - o1;
    float e3 = c[3][0] * (a[5] - a[7] - a[8]);
    float o3 = c[3][0] * (b[5] - b[7] - b[8]) - o2;
    float e4 = c[4][0] * a[0] + c[4][1] * a[1] + c[4][2] * a[2] + c[4][3] * a[3] + a[4];
    float o4 = c[4][0] * b[0] + c[4][1] * b[1] + c[4][2] * b[2] + c[4][3] * b[3] + b[4] - o3;
    float e5 = c[5][0] * a[5] + c[5][1] * a[6] + c[5][2] * a[7] + c[5][3] * a[8];
    float o5 = c[5][0] * b[5] + c[5][1] * b[6] + c[5][2] * b[7] + c[5][3] * b[8] - o4;
    float e6 = 0.5f * (a[0] + a[2] + a[3]) - a[1] - a[4];
    float o6 = 0.5f * (b[0] + b[2] + b[3]) - b[1] - b[4] - o5;
    float e7 = c[7][0] * a[5] + c[7][1] * a[6] + c[7][2] * a[7] + c[7][3] * a[8];
    float o7 = c[7][0] * b[5] + c[7][1] * b[6] + c[7][2] * b[7] + c[7][3] * b[8] - o6;
    float e8 = c[8][0] * a[0] + c[8][1] * a[1] + c[8][2] * a[2] + c[8][3] * a[3] + a[4];
    float o8 = c[8][0] * b[0] + c[8][1] * b[1] + c[8][2] * b[2] + c[8][3] * b[3] + b[4] - o7;
    y[0] = e0;
    float r = o0 - e0;   y[1] = r;
    r = e1 - r;          y[2] = r;
    r = o1 - r;          y[3] = r;
    r = e2 - r;          y[4] = r;
    r = o2 - r;          y[5] = r;
    r = e3 - r;          y[6] = r;
    r = o3 - r;          y[7] = r;
    r = e4 - r;          y[8] = r;
    r = o4 - r;          y[9] = r;
    r = e5 - r;          y[10] = r;
    r = o5 - r;          y[11] = r;
    r = e6 - r;          y[12] = r;
    r = o6 - r;          y[13] = r;
    r = e7 - r;          y[14] = r;
    r = o7 - r;          y[15] = r;
    r = e8 - r;          y[16] = r;
    r = o8 - r;          y[17] = r;
}

// Long-block hybrid for one sub-band: window-and-fold of (previous, current) 18 samples then the
// 18-point core (hwin.c:147-172).  prev/cur are the frequency-inverted sub-band samples.
HMP3_HD void hybrid_long_band(const EncTables *T, const float *prev, const float *cur, int bt, float *out) {
    const float *w = T->win[bt];
    float f[18];
#pragma unroll
    for (int j = 0; j < 9; j++) {
        f[j] = w[26 - j] * cur[8 - j] + w[27 + j] * cur[9 + j];
        f[9 + j] = w[j] * prev[j] + w[17 - j] * prev[17 - j];
    }
    mdct18_core(T, f, out);
}

// Short-block hybrid for one sub-band: three overlapped 12-sample windows -> three 6-point
// transforms, written window-major at out[192*w + k] (hwin.c:228-268, emdct.c:252-303).
HMP3_HD void hybrid_short_band(const EncTables *T, const float *x1, const float *x2, float *out /* stride 192 */) {
    const float *w = T->win[2];
    float f[18];
    f[0] = w[8] * x1[14] + w[9] * x1[15];
    f[1] = w[7] * x1[13] + w[10] * x1[16];
    f[2] = w[6] * x1[12] + w[11] * x1[17];
    f[3] = w[0] * x1[6] + w[5] * x1[11];
    f[4] = w[1] * x1[7] + w[4] * x1[10];
    f[5] = w[2] * x1[8] + w[3] * x1[9];
    f[6] = w[8] * x2[2] + w[9] * x2[3];
    f[7] = w[7] * x2[1] + w[10] * x2[4];
    f[8] = w[6] * x2[0] + w[11] * x2[5];
    f[9] = w[0] * x1[12] + w[5] * x1[17];
    f[10] = w[1] * x1[13] + w[4] * x1[16];
    f[11] = w[2] * x1[14] + w[3] * x1[15];
    f[12] = w[8] * x2[8] + w[9] * x2[9];
    f[13] = w[7] * x2[7] + w[10] * x2[10];
    f[14] = w[6] * x2[6] + w[11] * x2[11];
    f[15] = w[0] * x2[0] + w[5] * x2[5];
    f[16] = w[1] * x2[1] + w[4] * x2[4];
    f[17] = w[2] * x2[2] + w[3] * x2[3];
    const float *v = T->m6_v, *v2 = T->m6_v2;
    const float c87 = T->m6_c;
#pragma unroll
    for (int win = 0; win < 3; win++) {
        const float *g = f + 6 * win;
        float a[6];
#pragma unroll
        for (int p = 0; p < 3; p++) {
            float g1 = v[p] * g[p];
            float g2 = v[5 - p] * g[5 - p];
            a[p] = g1 + g2;
            a[3 + p] = v2[p] * (g1 - g2);
        }
        float a02 = (a[0] + a[2]);
        float b02 = (a[3] + a[5]);
        float c0 = a02 + a[1];
        float c1 = b02 + a[4];
        float c2 = c87 * (a[0] - a[2]);
        float c3 = c87 * (a[3] - a[5]) - c1;
        c1 = c1 - c0;
        c2 = c2 - c1;
        float c4 = a02 - a[1] - a[1];
        float c5 = b02 - a[4] - a[4] - c3;
        c3 = c3 - c2;
        c4 = c4 - c3;
        c5 = c5 - c4;
        float *o = out + 192 * win;
        o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3; o[4] = c4; o[5] = c5;
    }
}

// Alias-reduction butterflies across the boundary above sub-band `sb` (hwin.c:298-319).
// `last` = the final processed boundary, where only the lower side is scaled.
HMP3_HD void alias_boundary(const EncTables *T, float *x /* xr + 18*sb */, bool last) {
    if (!last) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float a = x[17 - i], b = x[18 + i];
            x[17 - i] = a * T->csa[0][i] + b * T->csa[1][i];
            x[18 + i] = b * T->csa[0][i] - a * T->csa[1][i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) x[17 - i] = x[17 - i] * T->csa[0][i];
    }
}

}  // namespace hmp3
