// The second allocator of the reference, CBitAllo1 (bitallo1.cpp:96-1813): used for dual channel (-M2) and whenever
// intensity stereo is in use (-N below the coded bandwidth; MPEG-2 joint stereo at low rates).  Long blocks only.
// It works on estimates -- bits and noise as functions of the band maximum -- in two stages, quantises once and
// then steps all bands together until the counted bits fit.  Plain sequential code on both the host and the device
// (every lane of the stream's group runs it uniformly): these configurations are a completeness item, not the
// throughput path.
//
// The reference is C++, so its per-granule calls log(x) / log10(x) / sqrt(x) on float arguments are glibc's
// logf / log10f / sqrtf.  log10f is far from correctly rounded (it differs from the rounded double result for 4 % of
// its arguments), so both are restated here operation for operation: logf as glibc >= 2.27 computes it (table of 16
// intervals, cubic in double precision; sysdeps/ieee754/flt-32/e_logf.c, constants read from the libm of the build
// image and checked against it on 3e7 arguments, bit for bit), log10f as fdlibm's e_log10f.c on top of it.
#pragma once
#include "rate_common.h"

namespace hmp3 {

HMP3_CONST_TABLE double kLogfTab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2}, {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1.0p+0, 0x0.0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};

HMP3_HD float glibc_logf(float x) {  // positive finite normal or subnormal x (all this allocator passes)
#ifdef HMP3_A1_DOUBLE_LOG
    return (float)log((double)x);
#endif
    unsigned ix = f2u(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return -1.0f / 0.0f;
        if (ix == 0x7f800000u) return x;
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return 0.0f / 0.0f;
        ix = f2u(x * 8388608.0f);
        ix -= 23u << 23;
    }
    const unsigned tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) % 16u);
    const int k = (int)tmp >> 23;
    const unsigned iz = ix - (tmp & (0x1ffu << 23));
    const double invc = kLogfTab[i][0], logc = kLogfTab[i][1];
    const double z = (double)u2f(iz);
    const double r = z * invc - 1.0;
    const double y0 = logc + (double)k * 0x1.62e42fefa39efp-1;
    const double r2 = r * r;
    double y = 0x1.5575b0be00b6ap-2 * r + -0x1.ffffef20a4123p-2;
    y = -0x1.00ea348b88334p-2 * r2 + y;
    y = y * r2 + (y0 + r);
    return (float)y;
}
HMP3_HD float glibc_log10f(float x) {
#ifdef HMP3_A1_DOUBLE_LOG
    return (float)log10((double)x);
#endif
    const float two25 = 3.3554432000e+07f, ivln10 = 4.3429449201e-01f, log10_2hi = 3.0102920532e-01f,
                log10_2lo = 7.9034151668e-07f;
    int hx = (int)f2u(x), k = 0;
    if (hx < 0x00800000) {
        if ((hx & 0x7fffffff) == 0) return -two25 / 0.0f;
        if (hx < 0) return 0.0f / 0.0f;
        k -= 25;
        x *= two25;
        hx = (int)f2u(x);
    }
    if (hx >= 0x7f800000) return x + x;
    k += (hx >> 23) - 127;
    const int i = (int)(((unsigned)k & 0x80000000u) >> 31);
    hx = (hx & 0x007fffff) | ((0x7f - i) << 23);
    const float y = (float)(k + i);
    const float z = y * log10_2lo + ivln10 * glibc_logf(u2f((unsigned)hx));
    return z + y * log10_2hi;
}

HMP3_CONST_TABLE unsigned char kA1Pre2[21] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 17, 17, 17, 17, 19, 19, 21, 21, 21, 19};
HMP3_CONST_TABLE unsigned char kA1Pre4[21] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 35, 35, 35, 35, 39, 39, 43, 43, 43, 39};

// State of one stream's CBitAllo1.  The first block persists from granule to granule.
struct Allo1 {
    int gsf[2][21], gsf_save[2][21], sf[2][21];
    int bitadjust, bitadjust_save[2], call_count;
    float running_a, ave_alpha_nmr, alpha_nmr;
    // per call
    int nchan, ms, min_bits, max_bits, target_bits, max_cnt_bits, target0_bits, target0_min, target0_max;
    int noise_from_lines;  // which noise estimate function_noise_cb uses (stage 2: from the lines)
    float dBG, dGdB, x34mm;
    int gzero[2][21], gmin[2][21], ixmax[2][21], last_gsf[2][21];
    float xsxx[2][21], x34max[2][21], mask[2][21], noise[2][21];
    int G[2], preemp[2], sf_scale[2], huff_bits[2];
    RegionPlan plan[2];
};

HMP3_FN void allo1_init(const EncTables *T, Allo1 *A) {  // bitallo1.cpp:107-203
    unsigned char *p = (unsigned char *)A;
    for (unsigned i = 0; i < sizeof(Allo1); i++) p[i] = 0;
    A->bitadjust = -100;
    A->bitadjust_save[0] = A->bitadjust_save[1] = -100;
    for (int c = 0; c < 2; c++)
        for (int j = 0; j < T->cfg.nsf[c]; j++) A->gsf[c][j] = A->gsf_save[c][j] = 35;
    A->running_a = (1.0f / 20.0f);
    A->ave_alpha_nmr = 40.0f;
}

// CBitAllo1::ms_correlation2 (bitallo1.cpp:393-433): x = the granule's two spectra
HMP3_FN int allo1_ms_measure(const EncTables *T, const float *x0, const float *x1) {
    int d = 0, k = 0;
    const int n0 = T->cfg.nsf[0];
    for (int i = 0; i < n0; i++) {
        const int n = T->nBand_l_iso[i];
        float s0 = 0.0f, s1 = 0.0f;
        for (int j = 0; j < n; j++, k++) {
            float a = x0[k] * x0[k];
            const float b = x1[k] * x1[k];
            s0 += (a + b);
            a = fabsf(a - b);
            s1 += a;
        }
        if (s1 > 0.80 * s0) d++;
        if (s1 > 0.95 * s0) d += 2;
    }
    return n0 - 3 * d;
}

// x[ch] = pointer to the channel's 576 lines (the channels of a dual-channel call are allocated one at a time)
struct Allo1Io {
    float *xr[2];
    float *x34[2];
    QLine *ix[2];
    unsigned *sign[2];  // sign words of the channel (18 per channel)
    const SigMask *sm[2];
};

HMP3_HD void a1_set_sign(unsigned *w, int k, int s) {
    if (s) w[k >> 5] |= 1u << (k & 31);
    else w[k >> 5] &= ~(1u << (k & 31));
}
HMP3_HD float a1_mask_of(const EncTables *T, float sig, float mask, float xsxx, int i) {
    const float r = sig / (mask * (0.1f + 0.0001f * xsxx));
    if (r < 1.0e-10f) return 100.0f;
    return (float)(-10.0 * glibc_log10f(r) - T->a1_log_cbw[i]);
}

// signs, band energies, masks (bitallo1.cpp:640-679)
HMP3_FN void allo1_smr_adj(const EncTables *T, Allo1 *A, Allo1Io *io, const int *nsf) {
    for (int ch = 0; ch < A->nchan; ch++) {
        int k = 0;
        float *x = io->xr[ch];
        for (int i = 0; i < nsf[ch]; i++) {
            float e = 1.0e-12f;
            const int n = T->nBand_l_iso[i];
            for (int j = 0; j < n; j++, k++) {
                int s = 0;
                if (x[k] < 0.0f) {
                    s = 1;
                    x[k] = -x[k];
                }
                a1_set_sign(io->sign[ch], k, s);
                e += x[k] * x[k];
            }
            A->xsxx[ch][i] = e;
        }
    }
    for (int ch = 0; ch < A->nchan; ch++)
        for (int i = 0; i < nsf[ch]; i++)
            A->mask[ch][i] = a1_mask_of(T, io->sm[ch][i].sig, io->sm[ch][i].mask, A->xsxx[ch][i], i);
}

// the same for joint stereo with intensity coding above band nsf[1] (bitallo1.cpp:683-908)
HMP3_FN void allo1_smr_adj_joint(const EncTables *T, Allo1 *A, Allo1Io *io, const int *nsf) {
    float *x0 = io->xr[0], *x1 = io->xr[1];
    if (A->ms == 0) {
        for (int ch = 0; ch < A->nchan; ch++) {
            int k = 0;
            float *x = io->xr[ch];
            for (int i = 0; i < nsf[1]; i++) {
                float e = 1.0e-12f;
                const int n = T->nBand_l_iso[i];
                for (int j = 0; j < n; j++, k++) {
                    int s = 0;
                    if (x[k] < 0.0f) {
                        s = 1;
                        x[k] = -x[k];
                    }
                    a1_set_sign(io->sign[ch], k, s);
                    e += x[k] * x[k];
                }
                A->xsxx[ch][i] = e;
            }
        }
    } else {
        int k = 0;
        for (int i = 0; i < nsf[1]; i++) {
            float e0 = 1.0e-12f, e1 = 1.0e-12f;
            const int n = T->nBand_l_iso[i];
            for (int j = 0; j < n; j++, k++) {
                e0 += x0[k] * x0[k];
                e1 += x1[k] * x1[k];
                const float a = T->a1_con707 * x0[k];
                const float b = T->a1_con707 * x1[k];
                x0[k] = a + b;
                x1[k] = a - b;
                int s0 = 0, s1 = 0;
                if (x0[k] < 0.0f) {
                    s0 = 1;
                    x0[k] = -x0[k];
                }
                if (x1[k] < 0.0f) {
                    s1 = 1;
                    x1[k] = -x1[k];
                }
                a1_set_sign(io->sign[0], k, s0);
                a1_set_sign(io->sign[1], k, s1);
            }
            A->xsxx[0][i] = e0;
            A->xsxx[1][i] = e1;
        }
    }
    if (A->ms) {  // thin out the side channel
        int k = T->startBand_l[5];
        for (int i = 5; i < nsf[1]; i++) {
            const int n = T->nBand_l_iso[i];
            for (int j = 0; j < n; j += 2, k += 2) {
                const float a = x1[k] * x1[k] + x1[k + 1] * x1[k + 1];
                const float b = T->a1_sparse[i] * (a + x0[k] * x0[k] + x0[k + 1] * x0[k + 1]);
                if (a < b) x1[k] = x1[k + 1] = 0.0f;
            }
        }
    }
    if (T->cfg.is_flag) {  // intensity part: the sum of the channels, scaled to the energy of the pair
        for (int i = nsf[1]; i < nsf[0]; i++) {
            float e0 = 1.0e-12f, e1 = 1.0e-12f;
            const int n = T->nBand_l_iso[i];
            int k = T->startBand_l[i];
            float r = 1.0f;
            for (int j = 0; j < n; j++, k++) {
                e0 += x0[k] * x0[k];
                e1 += x1[k] * x1[k];
                x0[k] = x0[k] + x1[k];
                r += x0[k] * x0[k];
                int s = 0;
                if (x0[k] < 0.0f) {
                    s = 1;
                    x0[k] = -x0[k];
                }
                a1_set_sign(io->sign[0], k, s);
            }
            A->xsxx[0][i] = e0;
            A->xsxx[1][i] = e1;
            if (T->cfg.h_id) {
                r = (float)(sqrt((e0 + e1 + 2.0 * sqrtf(e0 * e1)) / r));
                if (r > 1.5f) r = 1.5f;
            } else {
                const float a = e0 > e1 ? e0 : e1;
                r = (float)(sqrtf(a / r));
                if (r > 1.2f) r = 1.2f;
            }
            k = T->startBand_l[i];
            for (int j = 0; j < n; j++, k++) x0[k] = r * x0[k];
        }
    }
    for (int ch = 0; ch < A->nchan; ch++)
        for (int i = 0; i < nsf[1]; i++)
            A->mask[ch][i] = a1_mask_of(T, io->sm[ch][i].sig, io->sm[ch][i].mask, A->xsxx[ch][i], i);
    if (T->cfg.is_flag) {
        for (int i = nsf[1]; i < nsf[0]; i++) {
            const float r = (io->sm[0][i].sig + io->sm[1][i].sig) /
                            ((io->sm[0][i].mask + io->sm[1][i].mask) * (0.1f + 0.0001f * (A->xsxx[0][i] + A->xsxx[1][i])));
            if (r < 1.0e-10f) A->mask[0][i] = 100.0f;
            else A->mask[0][i] = (float)(-10.0 * glibc_log10f(r) - T->a1_log_cbw[i]);
        }
        for (int i = nsf[1]; i < nsf[0]; i++) {  // intensity position from the energy ratio
            if (T->cfg.h_id) {
                if (A->xsxx[0][i] <= A->xsxx[1][i]) {
                    const int k = (int)(32.0f * A->xsxx[0][i] / A->xsxx[1][i] + 0.5f);
                    A->sf[1][i] = T->a1_is_pos[k];
                } else {
                    const int k = (int)(32.0f * A->xsxx[1][i] / A->xsxx[0][i] + 0.5f);
                    A->sf[1][i] = 6 - T->a1_is_pos[k];
                }
            } else {
                if (A->xsxx[0][i] <= A->xsxx[1][i]) {
                    const int k = (int)(32.0f * A->xsxx[0][i] / A->xsxx[1][i] + 0.5f);
                    A->sf[1][i] = T->a1_is_pos[k];
                    if (A->sf[1][i] != 0) A->sf[1][i] -= 1;
                } else {
                    const int k = (int)(32.0f * A->xsxx[1][i] / A->xsxx[0][i] + 0.5f);
                    A->sf[1][i] = T->a1_is_pos[k];
                }
            }
        }
    }
    if (A->ms)
        for (int i = 0; i < nsf[1]; i++) A->mask[1][i] = A->mask[0][i] = 0.5f * (A->mask[0][i] + A->mask[1][i]);
}

// |x|^(3/4), band maxima, step bounds (bitallo1.cpp:596-636)
HMP3_FN void allo1_x34(const EncTables *T, Allo1 *A, Allo1Io *io, const int *nsf) {
    for (int ch = 0; ch < A->nchan; ch++) {
        const int n = T->startBand_l[nsf[ch]];
        for (int k = 0; k < n; k++) io->x34[ch][k] = pow34(T, io->xr[ch][k]);
    }
    A->x34mm = 0.0f;
    for (int ch = 0; ch < A->nchan; ch++) {
        int k = 0;
        for (int i = 0; i < nsf[ch]; i++) {
            float m = 0.0f;
            const int n = T->nBand_l_iso[i];
            for (int j = 0; j < n; j++, k++)
                if (m < io->x34[ch][k]) m = io->x34[ch][k];
            A->x34max[ch][i] = m;
            if (A->x34mm < m) A->x34mm = m;
            if (m < T->a1_gz_con0) A->gzero[ch][i] = 0;
            else A->gzero[ch][i] = (int)(T->a1_gz_con1 * glibc_logf(m) + T->a1_gz_con2);
            A->gmin[ch][i] = imax_(0, A->gzero[ch][i] - 70);
        }
    }
}

HMP3_FN void allo1_ixmax(const EncTables *T, Allo1 *A, const int *nsf) {  // :912-926
    for (int ch = 0; ch < A->nchan; ch++)
        for (int i = 0; i < nsf[ch]; i++)
            A->ixmax[ch][i] = (int)((0.5f - 0.0946f) + A->x34max[ch][i] * T->igain34[A->gsf[ch][i]]);
}
HMP3_FN int allo1_bit_est(const EncTables *T, Allo1 *A, const int *nsf) {  // :930-954
    int n = 0;
    for (int ch = 0; ch < A->nchan; ch++)
        for (int i = 0; i < nsf[ch]; i++) {
            const int ixm = A->ixmax[ch][i];
            int bits;
            if (ixm < 256) bits = T->a1_bits[ixm];
            else if (ixm < 512) bits = 16 * 11;
            else if (ixm < 2048) bits = 16 * 13;
            else bits = 16 * 15;
            n += T->nBand_l_iso[i] * bits;
        }
    return n >> 4;
}
// move all steps together until the estimate meets `target` (bitallo1.cpp:958-1120; both variants)
HMP3_FN int allo1_bit_seek(const EncTables *T, Allo1 *A, const int *nsf, int target) {
    allo1_ixmax(T, A, nsf);
    int nbits = allo1_bit_est(T, A, nsf);
    int delta = nbits - target;
    if (delta > 0) {
        for (int it = 0; it < 10; it++) {
            if (delta <= 0) break;
            int dG = (int)(A->dGdB * delta);
            if (dG < 1) dG = 1;
            for (int ch = 0; ch < A->nchan; ch++)
                for (int j = 0; j < nsf[ch]; j++) {
                    A->gsf[ch][j] += dG;
                    if (A->gsf[ch][j] > A->gzero[ch][j]) A->gsf[ch][j] = A->gzero[ch][j];
                }
            allo1_ixmax(T, A, nsf);
            nbits = allo1_bit_est(T, A, nsf);
            delta = nbits - target;
        }
        return nbits;
    }
    int mindelta = target >> 2;
    if (mindelta < 100) mindelta = 100;
    delta = -delta;
    if (delta < mindelta) return nbits;
    // the reference's loop counter is shared with its inner band loop (bitallo1.cpp:1005-1030): after one pass it
    // holds the band count of the last channel, so the loop only repeats while that count is below 10
    for (int i = 0; i < 10; i++) {
        int dG = (int)(A->dGdB * delta);
        if (dG < 1) dG = 1;
        int gz = 0;
        for (int ch = 0; ch < A->nchan; ch++)
            for (i = 0; i < nsf[ch]; i++) {
                A->gsf[ch][i] -= dG;
                if (A->gsf[ch][i] < 0) A->gsf[ch][i] = 0;
                gz |= A->gsf[ch][i];
            }
        allo1_ixmax(T, A, nsf);
        nbits = allo1_bit_est(T, A, nsf);
        delta = target - nbits;
        if (delta < mindelta) break;
        if (gz == 0) break;
    }
    return nbits;
}

HMP3_HD float a1_noise_of_max(const EncTables *T, int ixm, int gsf) {
    if (ixm < 256) return T->a1_f_ixmax[ixm] + 1.505f * gsf;
    ixm >>= 5;
    if (ixm > 255) ixm = 255;
    return T->a1_f_big_ixmax[ixm] + 1.505f * gsf;
}
HMP3_FN void allo1_noise(const EncTables *T, Allo1 *A, const int *nsf) {  // :1124-1144
    for (int ch = 0; ch < A->nchan; ch++)
        for (int i = 0; i < nsf[ch]; i++) A->noise[ch][i] = a1_noise_of_max(T, A->ixmax[ch][i], A->gsf[ch][i]);
}
HMP3_FN void allo1_noise_cb(const EncTables *T, Allo1 *A, int i, int ch) {  // :1148-1165
    const int ixm = (int)((0.5f - 0.0946f + 0.002f) + A->x34max[ch][i] * T->igain34[A->gsf[ch][i]]);
    A->ixmax[ch][i] = ixm;
    A->noise[ch][i] = a1_noise_of_max(T, ixm, A->gsf[ch][i]);
}
HMP3_FN void allo1_noise2_cb(const EncTables *T, Allo1 *A, const Allo1Io *io, int i, int ch) {  // :1169-1238
    if (A->gsf[ch][i] == A->last_gsf[ch][i]) return;
    A->last_gsf[ch][i] = A->gsf[ch][i];
    int k = T->startBand_l[i];
    const int n = T->nBand_l_iso[i];
    const float igain = T->igain34[A->gsf[ch][i]];
    float sum = 0.0f;
    for (int j = 0; j < n; j++, k++) {
        int ixm = (int)((0.5f - 0.0946f) + io->x34[ch][k] * igain);
        if (ixm < 256) sum += T->a1_f_ix[ixm];
        else {
            ixm >>= 5;
            if (ixm > 255) ixm = 255;
            sum += T->a1_f_big_ix[ixm];
        }
    }
    A->noise[ch][i] = (float)(10.0f * glibc_log10f(sum) - T->a1_log_cbw[i] + 1.505f * (A->gsf[ch][i]));
}
HMP3_FN void allo1_noise2(const EncTables *T, Allo1 *A, const Allo1Io *io, const int *nsf) {
    for (int ch = 0; ch < A->nchan; ch++)
        for (int i = 0; i < nsf[ch]; i++) allo1_noise2_cb(T, A, io, i, ch);
}
HMP3_FN void allo1_reset_last(Allo1 *A) {
    for (int ch = 0; ch < A->nchan; ch++)
        for (int i = 0; i < 21; i++) A->last_gsf[ch][i] = -9999;
}
HMP3_FN void allo1_quant(const EncTables *T, Allo1 *A, const Allo1Io *io, const int *nsf) {  // :1255-1287
    for (int ch = 0; ch < A->nchan; ch++)
        for (int i = 0; i < nsf[ch]; i++) {
            if (A->gsf[ch][i] == A->last_gsf[ch][i]) continue;
            A->last_gsf[ch][i] = A->gsf[ch][i];
            const int n = T->nBand_l_iso[i];
            int k = T->startBand_l[i];
            if (A->ixmax[ch][i] <= 0) {
                for (int j = 0; j < n; j++, k++) io->ix[ch][k] = 0;
            } else {
                const float igain = T->igain34[A->gsf[ch][i]];
                for (int j = 0; j < n; j++, k++) io->ix[ch][k] = (QLine)(int)((0.5f - 0.0946f) + io->x34[ch][k] * igain);
            }
        }
}

// even out noise-to-mask over the bands (bitallo1.cpp:1291-1395); returns the largest step change made
HMP3_FN int allo1_noise_seek(const EncTables *T, Allo1 *A, const Allo1Io *io, const int *nsf) {
    int n = 0;
    float asum = 0.0f;
    for (int ch = 0; ch < A->nchan; ch++)
        for (int i = 0; i < nsf[ch]; i++)
            if ((A->gsf[ch][i] > 0) && (A->gsf[ch][i] < A->gzero[ch][i])) {
                asum += A->noise[ch][i] - A->mask[ch][i];
                n++;
            }
    if (n <= 1) return 0;
    const float a = asum / n;
    A->alpha_nmr = a;
    int dgmax = 0;
    for (int ch = 0; ch < A->nchan; ch++)
        for (int cb = 0; cb < nsf[ch]; cb++) {
            float dn = A->noise[ch][cb] - A->mask[ch][cb] - a;
            if (dn > 1.0) {
                if (A->gsf[ch][cb] <= 0) continue;
                float dn0 = dn;
                int gsf0 = A->gsf[ch][cb];
                const int gsf00 = gsf0;
                for (int i = 0; i < 50; i++) {
                    if (A->gsf[ch][cb] <= 0) break;
                    const int dg = (int)(0.5f * dn + 0.5f);
                    if (dg <= 0) break;
                    A->gsf[ch][cb] -= dg;
                    if (A->gsf[ch][cb] < 0) A->gsf[ch][cb] = 0;
                    if (A->noise_from_lines) allo1_noise2_cb(T, A, io, cb, ch);
                    else allo1_noise_cb(T, A, cb, ch);
                    dn = A->noise[ch][cb] - A->mask[ch][cb] - a;
                    if (dn < -1.0f) {
                        dn = dn0 = 0.5f * dn0;
                        A->gsf[ch][cb] = gsf0;
                        continue;
                    }
                    dn0 = dn;
                    gsf0 = A->gsf[ch][cb];
                }
                const int dg = gsf00 - A->gsf[ch][cb];
                if (dg > dgmax) dgmax = dg;
            } else if (dn < -1.0f) {
                if (A->gsf[ch][cb] >= A->gzero[ch][cb]) continue;
                float dn0 = dn;
                int gsf0 = A->gsf[ch][cb];
                const int gsf00 = gsf0;
                for (int i = 0; i < 50; i++) {
                    if (A->gsf[ch][cb] >= A->gzero[ch][cb]) break;
                    const int dg = (int)(-0.5f * dn);
                    if (dg <= 0) break;
                    A->gsf[ch][cb] += dg;
                    if (A->gsf[ch][cb] >= A->gzero[ch][cb]) A->gsf[ch][cb] = A->gzero[ch][cb];
                    if (A->noise_from_lines) allo1_noise2_cb(T, A, io, cb, ch);
                    else allo1_noise_cb(T, A, cb, ch);
                    dn = A->noise[ch][cb] - A->mask[ch][cb] - a;
                    if (dn > 1.0f) {
                        dn = dn0 = 0.5f * dn0;
                        A->gsf[ch][cb] = gsf0;
                        continue;
                    }
                    dn0 = dn;
                    gsf0 = A->gsf[ch][cb];
                }
                const int dg = A->gsf[ch][cb] - gsf00;
                if (dg > dgmax) dgmax = dg;
            }
        }
    return dgmax;
}

// choose scalefac_scale / preflag and clamp (bitallo1.cpp:1418-1595)
HMP3_FN void allo1_sf_final(const EncTables *T, Allo1 *A, int ch, const int *nsf) {
    int *sf = A->sf[ch];
    int pre = 0, scale = 0;
    int n = 11;
    if (n > nsf[ch]) n = nsf[ch];
    for (int i = 0; i < n; i++)
        if (sf[i] > 31) {
            scale = 1;
            break;
        }
    if (T->cfg.h_id) {
        if (scale == 0)
            for (int i = 11; i < nsf[ch]; i++)
                if (sf[i] > kA1Pre2[i]) {
                    scale = 1;
                    break;
                }
        const int lim = scale ? 31 : 15, sh = scale ? 2 : 1;
        for (int i = 11; i < nsf[ch]; i++)
            if (sf[i] > lim) {
                pre = 1;
                break;
            }
        if (pre)
            for (int i = 11; i < nsf[ch]; i++)
                if ((sf[i] >> sh) < kPretab[i]) {
                    pre = 0;
                    break;
                }
        const int top = scale ? 63 : 31;
        for (int i = 0; i < n; i++)
            if (sf[i] > top) sf[i] = top;
        for (int i = 11; i < nsf[ch]; i++) {
            const int hi = pre ? (scale ? kA1Pre4[i] : kA1Pre2[i]) : lim;
            if (sf[i] > hi) sf[i] = hi;
        }
        A->preemp[ch] = pre;
    } else {
        if (scale == 0)
            for (int i = 11; i < nsf[ch]; i++)
                if (sf[i] > 15) {
                    scale = 1;
                    break;
                }
        const int top = scale ? 63 : 31, hi = scale ? 31 : 15;
        for (int i = 0; i < n; i++)
            if (sf[i] > top) sf[i] = top;
        for (int i = 11; i < nsf[ch]; i++)
            if (sf[i] > hi) sf[i] = hi;
        A->preemp[ch] = 0;
    }
    A->sf_scale[ch] = scale;
}

// global gains and scale factors from the band steps (bitallo1.cpp:1599-1670)
HMP3_FN int allo1_scale_factors(const EncTables *T, Allo1 *A, const int *nsf) {
    int gmin_all = 999;
    for (int ch = 0; ch < A->nchan; ch++) {
        int gtop = -1;
        for (int i = 0; i < nsf[ch]; i++) {
            A->gsf[ch][i] = imax_(A->gsf[ch][i], A->gmin[ch][i]);
            if ((A->ixmax[ch][i] > 0) && (A->gsf[ch][i] > gtop)) gtop = A->gsf[ch][i];
        }
        if (gtop < 0) {
            for (int i = 0; i < nsf[ch]; i++) {
                A->sf[ch][i] = 0;
                A->gsf[ch][i] = A->gzero[ch][i];
                if (A->gsf[ch][i] > gtop) gtop = A->gsf[ch][i];
            }
            A->preemp[ch] = 0;
            A->sf_scale[ch] = 0;
            A->G[ch] = gtop;
            if (100 < gmin_all) gmin_all = 100;
            continue;
        }
        for (int i = 0; i < nsf[ch]; i++) {
            A->sf[ch][i] = 0;
            if (A->ixmax[ch][i] > 0) A->sf[ch][i] = gtop - A->gsf[ch][i];
        }
        allo1_sf_final(T, A, ch, nsf);
        const int keep = A->sf_scale[ch] == 0 ? ~1 : ~3;
        for (int i = 0; i < nsf[ch]; i++) A->sf[ch][i] &= keep;
        for (int i = 0; i < nsf[ch]; i++) {
            A->gsf[ch][i] = gtop - A->sf[ch][i];
            if (A->gsf[ch][i] > A->gzero[ch][i]) A->gsf[ch][i] = A->gzero[ch][i];
        }
        A->G[ch] = gtop;
        if (gtop < gmin_all) gmin_all = gtop;
    }
    return gmin_all;
}

HMP3_FN int allo1_count(const EncTables *T, Allo1 *A, const Allo1Io *io, const int *nsf) {
    int bits = 0;
    for (int ch = 0; ch < A->nchan; ch++) {
        A->huff_bits[ch] = plan_regions_long(T, 0, A->ixmax[ch], io->ix[ch], nsf[ch], &A->plan[ch]);
        bits += A->huff_bits[ch];
    }
    return bits;
}

// the allocation proper (bitallo1.cpp:1674-1813)
HMP3_FN int allo1_allocate(const EncTables *T, Allo1 *A, const Allo1Io *io, const int *nsf) {
    allo1_reset_last(A);
    A->noise_from_lines = 0;
    int nbits = allo1_bit_seek(T, A, nsf, A->target0_bits);
    for (int i = 0; i < 4; i++) {
        allo1_noise(T, A, nsf);
        const int dsf = allo1_noise_seek(T, A, io, nsf);
        if (dsf <= 0) break;
        nbits = allo1_bit_seek(T, A, nsf, A->target0_bits);
        if (dsf < 2) break;
    }
    A->noise_from_lines = 1;
    for (int i = 0; i < 4; i++) {
        allo1_noise2(T, A, io, nsf);
        const int dsf = allo1_noise_seek(T, A, io, nsf);
        if (dsf <= 0) break;
        int target = (int)(A->target0_bits + 0.5f * A->dBG * (A->alpha_nmr - A->ave_alpha_nmr));
        if (target > A->target0_max) target = A->target0_max;
        else if (target < A->target0_min) target = A->target0_min;
        nbits = allo1_bit_seek(T, A, nsf, target);
        if (dsf < 2) break;
    }
    allo1_reset_last(A);
    allo1_scale_factors(T, A, nsf);
    allo1_ixmax(T, A, nsf);
    allo1_quant(T, A, io, nsf);
    int bits = allo1_count(T, A, io, nsf);
    A->bitadjust = A->bitadjust + ((bits - nbits - A->bitadjust) >> 3);
    int tmp = A->min_bits - bits;
    if (tmp > 0) {
        if (tmp > 200) tmp = 200;
        A->bitadjust = A->bitadjust - (tmp >> 2);
    }
    for (int j = 0; j < 3; j++) {
        if ((A->min_bits - bits) < 50) break;
        int dG = (int)(A->dGdB * (A->min_bits - bits));
        if (dG < 1) dG = 1;
        int gz = 0;
        for (int ch = 0; ch < A->nchan; ch++)
            for (int i = 0; i < nsf[ch]; i++) {
                A->gsf[ch][i] -= dG;
                if (A->gsf[ch][i] < 0) A->gsf[ch][i] = 0;
                gz |= A->gsf[ch][i];
            }
        allo1_scale_factors(T, A, nsf);
        allo1_ixmax(T, A, nsf);
        allo1_quant(T, A, io, nsf);
        bits = allo1_count(T, A, io, nsf);
        if (gz == 0) break;
    }
    for (int j = 0; j < 100; j++) {
        if (bits <= A->max_cnt_bits) break;
        int dG = (int)(A->dGdB * (bits - A->max_cnt_bits));
        if (dG < 1) dG = 1;
        for (int ch = 0; ch < A->nchan; ch++)
            for (int i = 0; i < nsf[ch]; i++) A->gsf[ch][i] += dG;
        const int GG = allo1_scale_factors(T, A, nsf);
        allo1_ixmax(T, A, nsf);
        allo1_quant(T, A, io, nsf);
        bits = allo1_count(T, A, io, nsf);
        if (GG >= 100) break;
    }
    for (int ch = 0; ch < A->nchan; ch++)
        for (int i = 0; i < nsf[ch]; i++)
            if (A->ixmax[ch][i] <= 0) A->sf[ch][i] = 0;
    return bits;
}

// CBitAllo1::BitAllo (bitallo1.cpp:205-389).  nchan = channels allocated by this call (1 for dual channel / mono
// style calls, then ch_arg says which), gr / sf_out = the side-info records and scale factors of those channels.
HMP3_FN void allo1_granule(const EncTables *T, Allo1 *A, Allo1Io *io, int ch_arg, int nchan, int min_bits, int target_bits_arg,
                           int max_bits, ScaleFac *sf_out, GrSide *gr, int ms) {
    int nsf[2];
    // a one-channel call works on "channel 0" of the allocator whichever stream channel it is (nsf[0] = band_limit_left)
    nsf[0] = T->cfg.nsf[0];
    nsf[1] = T->cfg.nsf[1];
    A->ms = ms;
    A->nchan = nchan;
    if (nchan == 1) A->dBG = 0.25f * T->startBand_l[nsf[0]];
    else A->dBG = 0.25f * (T->startBand_l[nsf[0]] + T->startBand_l[nsf[1]]);
    A->dGdB = 1.0f / A->dBG;
    if (nchan == 1) A->bitadjust = A->bitadjust_save[ch_arg];
    A->max_bits = max_bits;
    A->min_bits = min_bits < 0 ? 0 : min_bits;
    A->target_bits = target_bits_arg - (target_bits_arg >> 4);
    if (A->target_bits < A->min_bits) A->target_bits = A->min_bits;
    if (T->cfg.is_flag == 0) allo1_smr_adj(T, A, io, nsf);
    else allo1_smr_adj_joint(T, A, io, nsf);
    allo1_x34(T, A, io, nsf);
    if (A->x34mm < 3.0f) {
        for (int i = 0; i < nchan; i++) {
            GrSide *g = gr + i;
            g->global_gain = 0;
            g->window_switching_flag = 0;
            g->block_type = 0;
            g->mixed_block_flag = 0;
            g->preflag = 0;
            g->scalefac_scale = 0;
            g->table_select[0] = g->table_select[1] = g->table_select[2] = 0;
            g->big_values = 0;
            g->region0_count = g->region1_count = 0;
            g->count1table_select = 0;
            g->aux_nquads = 0;
            g->aux_bits = 0;
            g->aux_not_null = 0;
            g->aux_nreg[0] = g->aux_nreg[1] = g->aux_nreg[2] = 0;
            for (int j = 0; j < 21; j++) sf_out[i].l[j] = 0;
        }
        return;
    }
    A->call_count++;
    if (A->call_count <= 20) A->running_a = 1.0f / A->call_count;
    A->max_cnt_bits = A->max_bits;
    if (A->target_bits < A->min_bits) A->target_bits = A->min_bits;
    A->target0_min = A->target_bits >> 1;
    if (A->target0_min < A->min_bits) A->target0_min = A->min_bits;
    A->target0_max = (A->target_bits + A->max_bits) >> 1;
    if (A->bitadjust > (A->target_bits >> 1)) A->bitadjust = A->target_bits >> 1;
    A->target0_bits = A->target_bits - A->bitadjust;
    A->target0_min -= A->bitadjust;
    A->target0_max -= A->bitadjust;
    if (nchan == 1) {
        for (int i = 0; i < nsf[0]; i++) {
            A->gsf[0][i] = A->gsf_save[ch_arg][i];
            if (A->gsf[0][i] > A->gzero[0][i]) A->gsf[0][i] = A->gzero[0][i];
        }
    } else {
        for (int ch = 0; ch < nchan; ch++)
            for (int i = 0; i < nsf[ch]; i++)
                if (A->gsf[ch][i] > A->gzero[ch][i]) A->gsf[ch][i] = A->gzero[ch][i];
    }
    allo1_allocate(T, A, io, nsf);
    A->ave_alpha_nmr = A->ave_alpha_nmr + A->running_a * (A->alpha_nmr - A->ave_alpha_nmr);
    // ---- output_sf (bitallo1.cpp:546-592)
    for (int ch = 0; ch < nchan; ch++) {
        const int sh = A->sf_scale[ch] == 0 ? 1 : 2;
        for (int i = 0; i < nsf[ch]; i++) A->sf[ch][i] >>= sh;
        if (A->preemp[ch])
            for (int i = 11; i < nsf[ch]; i++) A->sf[ch][i] -= kPretab[i];
    }
    if (T->cfg.is_flag)
        for (int i = nsf[1] - 1; i >= 0; i--) {
            if (A->ixmax[1][i] > 0) break;
            A->sf[1][i] = T->cfg.ill_is_pos;
        }
    for (int ch = 0; ch < nchan; ch++)
        for (int i = 0; i < 21; i++) sf_out[ch].l[i] = A->sf[ch][i];
    for (int i = 0; i < nchan; i++) {
        GrSide *g = gr + i;
        g->global_gain = imin_(A->G[i] + (4 * 32 + 14), 255);
        g->window_switching_flag = 0;
        g->block_type = 0;
        g->mixed_block_flag = 0;
        g->preflag = A->preemp[i];
        g->scalefac_scale = A->sf_scale[i];
        g->aux_bits = A->huff_bits[i];
        g->aux_not_null = A->huff_bits[i];
        plan_to_side(T, &A->plan[i], g);
    }
    if (T->cfg.is_flag) gr[1].aux_not_null = 1;  // the right channel's scale factors carry the intensity positions
    if (nchan == 1) {
        for (int i = 0; i < nsf[0]; i++) A->gsf_save[ch_arg][i] = A->gsf[0][i];
        A->bitadjust_save[ch_arg] = A->bitadjust;
    }
}

}  // namespace hmp3
