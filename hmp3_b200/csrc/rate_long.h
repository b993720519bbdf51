// Rate loop for long block types (0, 1, 3): noise-target computation, per-band step search, scale
// factor selection, quantisation, Huffman region planning and the bit-budget control loops.
// Behaviour follows CBitAllo3 (bitallo3.cpp:484-3149); integer outputs are bit-exact given identical
// inputs.  All state of one stream lives in a LongRate object (persistent fields keep their value
// from granule to granule exactly as the reference object does).
#pragma once
#include "rate_common.h"
#include "psy_core.h"
#include "prepare.h"

namespace hmp3 {

constexpr int kPart23Max = 4021;  // bitallo3.h:61

struct LongRate {
    // ---- carried from granule to granule
    int mnr, pool_fraction, calls, delta_mnr, ms_memory;
    int hf_quant, hf_quant_ch[2], gsf_hf, gsf_hf_ch[2];
    int huff_bits[2];
    int nt_adjust[2][22];
    int ixmax[2][22];
    int sf[2][22], active[2][22];
    // ---- per call
    int nchan, ms, block_type;
    int max_bits, max_target, min_target, target, pool_bits, active_lines;
    float xsxx[2][22], x34max[2][22];
    int snr[2][22], noise0[2][22], noise[2][22], nt[2][22];
    int gzero[2][22], gmin[2][22], gsf[2][22];
    int G[2], preemp[2], sf_scale[2];
    RegionPlan plan[2];
    int gsave[2][22];     // long_more_bits: the steps before the trial pass
    int peak[22], peak10[22];  // long_trade_peaks: quantised band maxima of the channel in hand
    float (*x34)[576];  // |x|^(3/4) of the granule in flight: points into its PrepGranule
    const float *exx[2];  // band energies of the spectra the loop works on (left/right, or mid/side), same place
    int n_exx[2];         // bands they cover
};

#if HMP3_COOP
// One 576-float row of shared memory per stream of the block: per-line scratch of the line-parallel sections
// (step search, sparse-band refit); nothing is kept in it between sections.
extern __shared__ float s_rate_rows[];  // [streams of the block][576], dynamic (a 32-warp block needs 72 KB)
__device__ __forceinline__ float *rate_scratch_row() {
    return s_rate_rows + 576 * ((threadIdx.x / HMP3_W) % (kRateWarpsPerBlock * (32 / HMP3_W)));
}
#endif

HMP3_FN void long_rate_init(const EncTables *T, LongRate *L) {  // bitallo3.cpp:288-480
    L->mnr = T->cfg.initial_mnr;
    L->pool_fraction = T->cfg.vbr_flag ? 614 : 0;
    L->calls = 0;
    L->delta_mnr = 0;
    L->ms_memory = 0;
    L->hf_quant = 0;
    L->hf_quant_ch[0] = L->hf_quant_ch[1] = 0;
    L->gsf_hf = -1;
    L->gsf_hf_ch[0] = L->gsf_hf_ch[1] = -1;
    L->huff_bits[0] = L->huff_bits[1] = 0;
    for (int c = 0; c < 2; c++) {
        for (int i = 0; i < 22; i++) {
            L->nt_adjust[c][i] = 0;
            L->ixmax[c][i] = 0;
            L->sf[c][i] = 0;
            L->active[c][i] = 0;
            L->gzero[c][i] = L->gmin[c][i] = L->gsf[c][i] = 0;
            L->noise[c][i] = L->noise0[c][i] = L->nt[c][i] = L->snr[c][i] = 0;
            L->x34max[c][i] = L->xsxx[c][i] = 0.0f;
        }
        L->G[c] = L->preemp[c] = L->sf_scale[c] = 0;
        L->plan[c].bits = 0;
        L->plan[c].nbig = L->plan[c].nquads = 0;
        for (int i = 0; i < 4; i++) L->plan[c].table[i] = 0;
        for (int i = 0; i < 3; i++) L->plan[c].cb[i] = 0;
    }
}

// ---- scale-factor range tables by (scalefac_scale, preflag) (bitallo3.cpp:87-161)
HMP3_HD int sf_pre_amount(int i) {  // ISO pretab
    return kPretab[i];
}
HMP3_HD int sf_select_limit(int sel, int i) {  // limits used to choose (scale, preflag)
    const int scale = sel >> 1, pre = sel & 1;
    const int base = (i < 11) ? 31 : 15;
    const int step = 2 << scale;  // 2 or 4
    return (scale ? 2 * base : base) + (pre ? step * sf_pre_amount(i) : 0);
}
HMP3_HD int sf_upper(int scale, int pre, int i) {
    const int base = (i < 11) ? 30 : 14;
    const int step = 2 << scale;
    return (scale ? 2 * base : base) + (pre ? step * sf_pre_amount(i) : 0);
}
HMP3_HD int sf_lower(int scale, int pre, int i) {
    const int step = 2 << scale;
    return pre ? step * sf_pre_amount(i) : 0;
}
HMP3_HD int noise_gap_limit(int i) {  // bitallo3.cpp:182-188
    return i < 15 ? 250 : (i == 15 ? 300 : (i < 18 ? 400 : (i < 20 ? 500 : 600)));
}

// ------------------------------------------------------------------ startup
// dropout prevention applied to a band's noise target (bitallo3.cpp:858-865)
HMP3_HD int nt_dropout_guard(int noise0, int nt) {
    int tsnr = noise0 - nt;
    if (tsnr < 300) {
        tsnr = 187 + ((3 * tsnr) >> 3) - tsnr;
        nt -= tsnr;
    }
    return nt;
}

// pull noise targets toward their band-weighted mean (bitallo3.cpp:1069-1126)
HMP3_FN void long_flatten_targets(const EncTables *T, LongRate *L) {
    const int f = T->cfg.nt_flatten;
    if (f == 0) return;
    HMP3_SYNC();
    for (int ch = 0; ch < L->nchan; ch++) {
        const int nsf = T->cfg.nsf[ch];
        int na = 0, nab = 0, ab = 0;  // integer sums: the order of the bands does not matter
        HMP3_FOR_LANES(i, nsf) {
            int th = i < 14 ? 0 : (i < 17 ? 100 : (i == 17 ? 200 : 300));
            if (L->snr[ch][i] > th) {
                na++;
                ab += T->nBand_l[i] * L->nt[ch][i];
                nab += T->nBand_l[i];
            }
        }
        na = 1 + gsum(na);
        nab = 1 + gsum(nab);
        ab = gsum(ab) / nab;
        if (na < 5) continue;
        HMP3_FOR_LANES(i, nsf) {
            int th = i < 14 ? 0 : (i < 17 ? 100 : (i == 17 ? 200 : 300));
            if (L->snr[ch][i] > th) {
                int dmax = imax_(L->snr[ch][i] - 400, 0);
                int d = (f * (ab - L->nt[ch][i])) >> 4;
                d = imin_(d, dmax);
                L->nt[ch][i] = L->nt[ch][i] + d;
            }
        }
    }
    HMP3_SYNC();
}

// Take over what the parallel prepare pass computed for this granule (prepare.h): band energies, step bounds,
// the |x|^(3/4) array (by reference) and the signs (expanded into the persistent sign array for the lines that
// were rewritten, so that lines beyond them keep their old value like the reference's signx buffer).
// Merge the sign bits of the first nl lines (packed 32 per word) into the persistent sign words: lines beyond nl keep
// their old value like the reference's signx buffer.
HMP3_HD void merge_sign_word(unsigned *dst, unsigned src, int w, int nl) {
    if (32 * w + 32 <= nl) *dst = src;
    else if (32 * w < nl) {
        const unsigned m = (1u << (nl - 32 * w)) - 1u;
        *dst = (*dst & ~m) | (src & m);
    }
}
HMP3_FN const PrepGranule *long_adopt_prepared(const EncTables *T, LongRate *L, PrepGranule *P, unsigned *signw /*[2][18]*/,
                                               int n_energy, const int *n_bounds /*[2]*/) {
    L->x34 = P->x34;
#if HMP3_COOP
    const PrepGranule *Q = P;
#if HMP3_W == 32
    {   // the granule's band records (signs, energies, maxima, step bounds: the kilobyte behind the two x34 rows) are
        // touched here for the first time -- bring them to the warp's scratch row in one go (cp.async) instead of one
        // cold global load per little loop below; Q addresses the staged copy with PrepGranule's own field offsets
        float *row = rate_scratch_row();
        const char *src = (const char *)&P->sign[0][0];
        constexpr int kBytes = (int)(sizeof(PrepGranule) - sizeof(P->x34));
        static_assert(kBytes % 8 == 0 && kBytes <= 2304, "band records of a PrepGranule fit the scratch row");
        for (int o = 8 * HMP3_LANE; o < kBytes; o += 8 * 32)
            asm volatile("{ .reg .u64 a; cvta.to.shared.u64 a, %0; cp.async.ca.shared.global [a], [%1], 8; }" ::"l"((char *)row + o), "l"(src + o));
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        Q = (const PrepGranule *)((const char *)row - sizeof(P->x34));
    }
#endif
    HMP3_SYNC();
    for (int ch = 0; ch < L->nchan; ch++) {
        const int ne = n_energy < 0 ? n_bounds[ch] : n_energy;
        for (int i = HMP3_LANE; i < ne; i += HMP3_W) L->xsxx[ch][i] = Q->xsxx[ch][i];
        for (int i = HMP3_LANE; i < n_bounds[ch]; i += HMP3_W) {
            L->x34max[ch][i] = Q->x34max[ch][i];
            L->gzero[ch][i] = Q->gzero[ch][i];
            L->gmin[ch][i] = Q->gmin[ch][i];
        }
        const int nl = Q->nlines[ch];
        for (int w = HMP3_LANE; w < 18; w += HMP3_W) merge_sign_word(signw + 18 * ch + w, Q->sign[ch][w], w, nl);
    }
    HMP3_SYNC();
    return Q;
#else
    for (int ch = 0; ch < L->nchan; ch++) {
        const int ne = n_energy < 0 ? n_bounds[ch] : n_energy;
        for (int i = 0; i < ne; i++) L->xsxx[ch][i] = P->xsxx[ch][i];
        for (int i = 0; i < n_bounds[ch]; i++) {
            L->x34max[ch][i] = P->x34max[ch][i];
            L->gzero[ch][i] = P->gzero[ch][i];
            L->gmin[ch][i] = P->gmin[ch][i];
        }
        for (int w = 0; w < 18; w++) merge_sign_word(signw + 18 * ch + w, P->sign[ch][w], w, P->nlines[ch]);
    }
    return P;
#endif
}

// left/right granule (bitallo3.cpp:816-898): noise targets from the prepared band energies.
HMP3_FN void long_startup_lr(const EncTables *T, LongRate *L, const SigMask *sm /*[2][36]*/, PrepGranule *P,
                             unsigned *signx /*[2][18]*/) {
    const int mnr = L->mnr + 100;
    long_adopt_prepared(T, L, P, signx, -1, T->cfg.nsf3);
    for (int ch = 0; ch < 2; ch++) {
        L->exx[ch] = P->xsxx[ch];
        L->n_exx[ch] = T->cfg.nsf3[ch];
    }
    int lines = 0;
    for (int ch = 0; ch < L->nchan; ch++) {
        HMP3_FOR_LANES(i, T->cfg.nsf[ch]) {
            const int cbw = T->log_cbw_l[i];
            L->noise0[ch][i] = mb_log(T, L->xsxx[ch][i]) - cbw;
            if (L->noise0[ch][i] < -2000) {
                L->nt[ch][i] = L->noise0[ch][i] + 1000;
            } else {
                lines += T->nBand_l[i];
                int mask = mb_log(T, sm[36 * ch + i].mask) - cbw;
                L->nt[ch][i] = nt_dropout_guard(L->noise0[ch][i], mask - mnr + T->taperNT[i]);
            }
            L->snr[ch][i] = L->noise0[ch][i] - L->nt[ch][i];
        }
    }
    L->active_lines = gsum(lines);
    long_flatten_targets(T, L);
}

// mid/side granule (bitallo3.cpp:902-1065): noise targets from the prepared energies (the spectra were rotated to
// |L+R|, |L-R| without 1/sqrt2 by the prepare pass; the global gain is lowered by 2 steps on output instead).
HMP3_FN void long_startup_ms(const EncTables *T, LongRate *L, const SigMask *sm, PrepGranule *P, unsigned *signx) {
    if (T->cfg.vbr_flag == 0 && L->calls > 10 && (L->target - L->min_target) < 100)
        L->mnr = imin_(L->mnr + 50, 2050);
    const int mnr = L->mnr;
    const int nsf0 = T->cfg.nsf[0];
    const PrepGranule *Q = long_adopt_prepared(T, L, P, signx, nsf0, T->cfg.nsf2);
    for (int ch = 0; ch < 2; ch++) {
        L->exx[ch] = P->e2[ch];
        L->n_exx[ch] = nsf0;
    }
    int lines = 0;
    HMP3_FOR_LANES(i, nsf0) {
        const int n = T->nBand_l[i];
        const float el = L->xsxx[0][i], er = L->xsxx[1][i], em = Q->e2[0][i], ed = Q->e2[1][i];
        const int cbw = T->log_cbw_l[i];
        int ntl, ntr;
        int n0l = mb_log(T, el) - cbw;
        if (n0l < -2000) ntl = 10000;
        else {
            ntl = nt_dropout_guard(n0l, (mb_log(T, sm[i].mask) - cbw) - mnr + T->taperNT[i]);
            lines += n;
        }
        int n0r = mb_log(T, er) - cbw;
        if (n0r < -2000) ntr = 10000;
        else {
            ntr = nt_dropout_guard(n0r, (mb_log(T, sm[36 + i].mask) - cbw) - mnr + T->taperNT[i]);
            lines += n;
        }
        L->nt[0][i] = ntl;
        L->nt[1][i] = ntr;
        L->snr[0][i] = n0l - ntl;
        L->snr[1][i] = n0r - ntr;
        L->noise0[0][i] = mb_log(T, em) - cbw;
        L->noise0[1][i] = mb_log(T, ed) - cbw;
    }
    L->active_lines = gsum(lines);
    long_flatten_targets(T, L);
    HMP3_FOR_LANES(i, nsf0) {
        const int nsum = L->noise0[0][i], ndiff = L->noise0[1][i];
        const int xnt = imin_(L->nt[0][i], L->nt[1][i]) + 300;
        L->nt[1][i] = L->nt[0][i] = xnt;
        if (ndiff < xnt) {
            L->nt[0][i] = mb_logsub(T, xnt, ndiff);
            if (i < 16) L->nt[0][i] -= 200;
        }
        if (nsum < xnt) L->nt[1][i] = mb_logsub(T, xnt, nsum);
        L->snr[0][i] = nsum - L->nt[0][i];
        L->snr[1][i] = ndiff - L->nt[1][i];
    }
    HMP3_SYNC();
}

// ------------------------------------------------------------------ per-band step search
HMP3_FN void long_seek_initial(const EncTables *T, LongRate *L) {  // bitallo3.cpp:1130-1160
    HMP3_SYNC();
    for (int ch = 0; ch < L->nchan; ch++)
        HMP3_FOR_LANES(i, T->cfg.nsf[ch]) {
            L->nt_adjust[ch][i] = imax_(L->nt_adjust[ch][i], -400);
            L->nt_adjust[ch][i] = imin_(L->nt_adjust[ch][i], 400);
            float g4 = 0.017716950f * mb_log(T, L->x34max[ch][i]) + (88.411238f - 100.0f + 8.0f);
            float d = (1.00f / 110.5f) * (1800 - 8 * i - (L->noise0[ch][i] - L->nt[ch][i] + L->nt_adjust[ch][i]));
            float g = g4 + d;
            int v = round_away(g);
            v = imin_(v, L->gzero[ch][i]);
            v = imax_(v, L->gmin[ch][i]);
            L->gsf[ch][i] = v;
        }
    HMP3_SYNC();
}

// walk the step of one band toward the noise target, at most 20 steps (bitallo3.cpp:1164-1238); plain
// sequential code, safe to run one band per lane
HMP3_FN int seek_finer(const EncTables *T, const float *y34, const float *y, int s0, int n, int logn, int target,
                       int dn, int *noise_io) {
    int s = s0 - 1;
    int best_abs = iabs(dn), best_noise = *noise_io, best_s = s0;
    const int niter = imin_(s, 20);
    for (int i = 0; i < niter; i++) {
        int tn = band_noise_seq(T, y34, y, s, n, logn);
        int a = iabs(tn - target);
        if (a < best_abs) { best_abs = a; best_noise = tn; best_s = s; }
        if (tn <= target) break;
        s--;
    }
    *noise_io = best_noise;
    return best_s;
}
HMP3_FN int seek_coarser(const EncTables *T, const float *y34, const float *y, int s0, int n, int logn, int target,
                         int dn, int *noise_io) {
    int s = s0;
    int best_abs = iabs(dn), best_noise = *noise_io, best_s = s0;
    for (int i = 0; i < 20; i++) {
        s++;
        int tn = band_noise_seq(T, y34, y, s, n, logn);
        int a = iabs(tn - target);
        if (a < best_abs) { best_abs = a; best_noise = tn; best_s = s; }
        if (tn >= target) break;
    }
    *noise_io = best_noise;
    return best_s;
}

#if HMP3_COOP
// max of v over the lanes named in `seg` (a lane mask inside the group, the caller's lane included); lanes with
// different masks may execute this together
__device__ __forceinline__ int gmax_seg(int v, unsigned seg) {
    int r;
    asm volatile("redux.sync.max.s32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(seg << HMP3_GSHIFT));
    return r;
}
// Device form of the per-band step search: all bands of both channels walk together.  Lane b owns band b of
// each channel (its search state lives in that lane's registers).  Every round (1) all lanes evaluate the
// squared error of every line whose band is still searching, at that band's current trial step, into L->dd;
// (2) each owning lane adds up its band IN LINE ORDER (so the float result is the sequential one) and advances
// its little state machine exactly as seek_finer / seek_coarser do.  Rounds end when no band is searching.
HMP3_FN void long_seek_actual(const EncTables *T, LongRate *L, const float *xr) {  // bitallo3.cpp:1242-1288
    constexpr int NS = (22 + HMP3_W - 1) / HMP3_W;  // band slots per lane: band = lane + HMP3_W * slot
    const int lane = HMP3_LANE;
    // per-line squared errors of one channel at a time, in shared memory (one 576-float row per stream of the block)
    float *dd = rate_scratch_row();
    HMP3_SYNC();
    // channel-major: one channel's bands are searched to the end before the other channel is touched, so that the
    // rounds of a search work on one channel's spectra (4.6 KB) instead of both (the bands are independent: the order
    // of the evaluations does not change any result).  The per-band state machine of a lane lives in scalars
    // (slot 0) plus, for groups narrower than 22 lanes, a second set (slot 1).
    for (int c = 0; c < L->nchan; c++) {
        int mode0 = 0, try0 = 0, target0 = 0, babs0 = 0, bnoise0 = 0, bs0 = 0, iter0 = 0, niter0 = 0;
        int mode1 = 0, try1 = 0, target1 = 0, babs1 = 0, bnoise1 = 0, bs1 = 0, iter1 = 0, niter1 = 0;
#define HMP3_SEEK_INIT(bnd, mode, stry, target)                  \
    if ((bnd) < T->cfg.nsf[c]) {                                 \
        target = L->nt[c][bnd];                                  \
        if (L->noise0[c][bnd] > target) {                        \
            mode = 1;                                            \
            stry = L->gsf[c][bnd];                               \
        } else {                                                 \
            L->gsf[c][bnd] = L->gzero[c][bnd] + 5;               \
            L->noise[c][bnd] = L->noise0[c][bnd];                \
        }                                                        \
    }
        HMP3_SEEK_INIT(lane, mode0, try0, target0)
        if (NS > 1) { HMP3_SEEK_INIT(lane + HMP3_W, mode1, try1, target1) }
#undef HMP3_SEEK_INIT
        const int nl = T->startBand_l[T->cfg.nsf[c]];
        const float *y34 = L->x34[c];
        const float *y = xr + 576 * c;
        // both rows of the channel on their way to L1 before the first round asks for them line by line
        for (int o = 32 * lane; o < nl; o += 32 * HMP3_W) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(y34 + o));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(y + o));
        }
        for (;;) {
            unsigned am = gballot(mode0 != 0);  // bit b = band b of the channel is still searching
            if (NS > 1) am |= gballot(mode1 != 0) << (HMP3_W & 31);
            if (am == 0) break;
            for (int k0 = 0; k0 < nl; k0 += HMP3_W) {
                const int kk = k0 + lane;
                const int bb = (kk < nl) ? (int)T->line_band_l[kk] : 0;
                const bool aa = (kk < nl) && ((am >> bb) & 1u);
                float v34 = 0.0f, vx = 0.0f;
                if (aa) {
                    v34 = y34[kk];
                    vx = y[kk];
                }
                if (gballot(aa) != 0) {
                    int st = gshfl(try0, bb & (HMP3_W - 1));
                    if (NS > 1) {
                        const int st1 = gshfl(try1, bb & (HMP3_W - 1));
                        if (bb >= HMP3_W) st = st1;
                    }
                    if (aa) dd[kk] = noise_line(T, T->igain34[st], T->gain[st], v34, vx);
                }
            }
            HMP3_SYNC();
#define HMP3_SEEK_STEP(bnd, mode, stry, target, babs, bnoise, bs, iter, niter)                \
    if (mode != 0) {                                                                          \
        const float acc = sum_seq(dd + T->startBand_l[bnd], T->nBand_l[bnd], 0.0f);           \
        const int tn = mb_log(T, 1.0e-12f + acc) - T->log_cbw_l[bnd];                         \
        bool done = false;                                                                    \
        if (mode == 1) {                                                                      \
            const int dn = tn - target;                                                       \
            L->nt_adjust[c][bnd] = L->nt_adjust[c][bnd] + (dn >> 3);                          \
            babs = iabs(dn);                                                                  \
            bnoise = tn;                                                                      \
            bs = stry;                                                                        \
            iter = 0;                                                                         \
            if (dn > 100) {                                                                   \
                niter = imin_(stry - 1, 20);                                                  \
                stry = stry - 1;                                                              \
                mode = 2;                                                                     \
                done = niter <= 0;                                                            \
            } else if (dn < -100) {                                                           \
                niter = 20;                                                                   \
                stry = stry + 1;                                                              \
                mode = 3;                                                                     \
            } else done = true;                                                               \
        } else {                                                                              \
            const int a = iabs(tn - target);                                                  \
            if (a < babs) {                                                                   \
                babs = a;                                                                     \
                bnoise = tn;                                                                  \
                bs = stry;                                                                    \
            }                                                                                 \
            iter++;                                                                           \
            if (mode == 2) {                                                                  \
                if (tn <= target || iter >= niter) done = true;                               \
                else stry--;                                                                  \
            } else {                                                                          \
                if (tn >= target || iter >= niter) done = true;                               \
                else stry++;                                                                  \
            }                                                                                 \
        }                                                                                     \
        if (done) {                                                                           \
            L->gsf[c][bnd] = bs;                                                              \
            L->noise[c][bnd] = bnoise;                                                        \
            mode = 0;                                                                         \
        }                                                                                     \
    }
            HMP3_SEEK_STEP(lane, mode0, try0, target0, babs0, bnoise0, bs0, iter0, niter0)
            if (NS > 1) { HMP3_SEEK_STEP(lane + HMP3_W, mode1, try1, target1, babs1, bnoise1, bs1, iter1, niter1) }
#undef HMP3_SEEK_STEP
            HMP3_SYNC();
        }
    }
    HMP3_SYNC();
}
#else
HMP3_FN void long_seek_actual(const EncTables *T, LongRate *L, const float *xr) {  // bitallo3.cpp:1242-1288
    for (int ch = 0; ch < L->nchan; ch++) {
        const float *y34 = L->x34[ch];
        const float *y = xr + 576 * ch;
        for (int i = 0; i < T->cfg.nsf[ch]; i++) {
            const int target = L->nt[ch][i];
            const int n = T->nBand_l[i];
            int s = L->gsf[ch][i];
            if (L->noise0[ch][i] > target) {
                const int logn = T->log_cbw_l[i];
                int noise = band_noise(T, y34, y, s, n, logn);
                int dn = noise - target;
                L->nt_adjust[ch][i] = L->nt_adjust[ch][i] + (dn >> 3);
                if (dn > 100) s = seek_finer(T, y34, y, s, n, logn, target, dn, &noise);
                else if (dn < -100) s = seek_coarser(T, y34, y, s, n, logn, target, dn, &noise);
                L->gsf[ch][i] = s;
                L->noise[ch][i] = noise;
            } else {
                L->gsf[ch][i] = L->gzero[ch][i] + 5;
                L->noise[ch][i] = L->noise0[ch][i];
            }
            y34 += n;
            y += n;
        }
    }
}

#endif

// flatten isolated small peaks in the upper bands of L/R granules (bitallo3.cpp:2216-2299)
// reciprocal of the value that quantises to n + 0.5 under the tuned rounding (bitallo3.cpp:2216-2232)
HMP3_CONST_TABLE float kInvPeak[16] = {1.0f / (0.5f + 0.09460f),  1.0f / (1.5f + 0.02799f),  1.0f / (2.5f + 0.01671f),
    1.0f / (3.5f + 0.01192f),  1.0f / (4.5f + 0.00927f),  1.0f / (5.5f + 0.00758f),
    1.0f / (6.5f + 0.00641f),  1.0f / (7.5f + 0.00556f),  1.0f / (8.5f + 0.00490f),
    1.0f / (9.5f + 0.00439f),  1.0f / (10.5f + 0.00397f), 1.0f / (11.5f + 0.00362f),
    1.0f / (12.5f + 0.00333f), 1.0f / (13.5f + 0.00309f), 1.0f / (14.5f + 0.00287f),
    1.0f / (15.5f + 0.00269f)};
HMP3_HD float db_of(float x) { return (float)(10.0 * log10((double)x)); }
HMP3_FN void long_trade_peaks(const EncTables *T, LongRate *L) {
    for (int ch = 0; ch < L->nchan; ch++) {
        const int nsf = T->cfg.nsf[ch];
        int *peak = L->peak, *peak10 = L->peak10;
        HMP3_SYNC();
        HMP3_FOR_LANES(i, nsf) {
            peak[i] = quant_tuned_peak(T, L->x34max[ch][i], L->gsf[ch][i]);
            peak10[i] = quant_tuned_peak10(T, L->x34max[ch][i], L->gsf[ch][i]);
            L->ixmax[ch][i] = peak[i];
        }
        HMP3_SYNC();
        int i;
        for (i = nsf - 1; i >= 11; i--) {
            if (peak10[i] > 16) break;
            if (peak[i] == 2) {
                float xg = 1.7717f * db_of(L->x34max[ch][i] * (1.0f / (1.5f + 0.02799f)));
                L->gsf[ch][i] = (int)(xg + 1.0f) + 8;
            }
        }
        const int k1 = i + 1;
        if (k1 < 9) continue;
        int k0 = (3 * k1) >> 2;
        if (k0 < 11) k0 = 11;
        if (k0 >= k1) continue;
        int top = 0;
        for (i = k0; i < k1; i++) top = imax_(top, peak[i]);
        if (top <= 2) continue;
        float fetot = 0, fepk = 0;
        for (i = k0; i < k1; i++) {
            float e = T->rnBand_l[i] * L->xsxx[ch][i];
            fetot += e;
            fepk += e * peak10[i];
        }
        float mean_pk = fepk / (1.0f + fetot);
        int tgt = (int)(0.1f * mean_pk + 0.65f);
        if (tgt < 2) tgt = 2;
        if (top <= tgt) continue;
        if (tgt > 15) continue;
        tgt = kPeakSnap[tgt];
        const float factor = kInvPeak[tgt];
        for (i = k0; i < k1; i++)
            if (peak[i] > tgt) {
                float xg = 1.7717f * db_of(L->x34max[ch][i] * factor);
                L->gsf[ch][i] = (int)(xg + 1.0f) + 8;
            }
    }
}

// -HF: decide whether the lines above band 21 can be coded at the granule's gain
// (bitallo3.cpp:2421-2565).  which = channel for L/R granules; -1 = the mid channel of an M/S granule.
HMP3_FN void long_hf_decide(const EncTables *T, LongRate *L, int ch, bool ms) {
    if (L->gzero[ch][21] <= 8) return;
    int gmax0 = 0, gmax1 = 0;
    for (int i = 0; i < 11; i++)
        if (L->gsf[ch][i] < L->gzero[ch][i] && L->gsf[ch][i] > gmax0) gmax0 = L->gsf[ch][i];
    for (int i = 11; i < T->cfg.nsf[ch]; i++)
        if (L->gsf[ch][i] < L->gzero[ch][i] && L->gsf[ch][i] > gmax1) gmax1 = L->gsf[ch][i];
    const int gtar = imax_(0, L->gzero[ch][21] - 5);
    const int gtar2 = imax_(0, L->gzero[ch][21] - 7);
    const int gmax = imax_(gmax0, gmax1);
    if (gtar >= gmax) {
        if (ms) { L->hf_quant = 1; L->gsf_hf = gtar2; }
        else { L->hf_quant_ch[ch] = 1; L->gsf_hf_ch[ch] = gtar2; }
    } else if (gmax0 > gmax1) {
        const int gset = imax_(gtar, gmax1);
        if (L->gzero[ch][21] > gset) {
            for (int i = 0; i < 11; i++)
                if (L->gsf[ch][i] < L->gzero[ch][i] && L->gsf[ch][i] > gset) L->gsf[ch][i] = gset;
            if (ms) L->hf_quant = 1;
            else L->hf_quant_ch[ch] = 1;
        }
    }
}
HMP3_FN void long_hf_adjust_lr(const EncTables *T, LongRate *L) {
    L->gsf_hf_ch[0] = L->gsf_hf_ch[1] = -1;
    for (int ch = 0; ch < L->nchan; ch++) long_hf_decide(T, L, ch, false);
    L->hf_quant = L->hf_quant_ch[0] | L->hf_quant_ch[1];
}
HMP3_HD void long_hf_reset_lr(LongRate *L) {
    L->hf_quant = 0;
    L->hf_quant_ch[0] = L->hf_quant_ch[1] = 0;
    L->gsf_hf_ch[0] = L->gsf_hf_ch[1] = -1;
    L->ixmax[0][21] = L->ixmax[1][21] = 0;
}
HMP3_FN void long_clear_hf_lines(const EncTables *T, QLine *ix, int nch) {  // bitallo3.cpp:1629-1654
    const int b = T->startBand_l[21], n = T->nBand_l[21];
    for (int ch = 0; ch < nch; ch++)
        for (int k = 0; k < n; k++) ix[576 * ch + b + k] = 0;
}

// ------------------------------------------------------------------ scale factors
// choose (scalefac_scale, preflag): first combination whose ranges hold every active band
// (bitallo3.cpp:1793-1888).  Bands are dealt over the lanes; the per-band sign tests are OR-reduced.
HMP3_FN void long_pick_sf_mode(const EncTables *T, LongRate *L, int ch) {
    const int nsf = T->cfg.nsf[ch];
    if (T->cfg.h_id) {
        int sp0 = 0, sp1 = 0, sp2 = 0, sp3 = 0;
        HMP3_FOR_LANES(i, nsf)
            if (L->active[ch][i]) {
                const int s = L->sf[ch][i];
                sp0 |= (sf_select_limit(0, i) - s);
                sp1 |= (sf_select_limit(1, i) - s);
                sp2 |= (sf_select_limit(2, i) - s);
                sp3 |= (sf_select_limit(3, i) - s);
                sp1 |= (s - sf_lower(0, 1, i));
                sp3 |= (s - sf_lower(1, 1, i));
            }
        // only the sign bits matter
        const int neg = gor(((sp0 >> 31) & 1) | ((sp1 >> 31) & 2) | ((sp2 >> 31) & 4) | ((sp3 >> 31) & 8));
        int scale, pre;
        if (!(neg & 1)) { scale = 0; pre = 0; }
        else if (!(neg & 2)) { scale = 0; pre = 1; }
        else if (!(neg & 4)) { scale = 1; pre = 0; }
        else if (!(neg & 8)) { scale = 1; pre = 1; }
        else { scale = 1; pre = 0; }
        L->preemp[ch] = pre;
        L->sf_scale[ch] = scale;
    } else {
        int sp0 = 0;
        HMP3_FOR_LANES(i, nsf)
            if (L->active[ch][i]) sp0 |= (sf_select_limit(0, i) - L->sf[ch][i]);
        const int neg = gor((sp0 >> 31) & 1);
        L->preemp[ch] = 0;
        L->sf_scale[ch] = neg ? 1 : 0;
    }
}

// derive G and scale factors from the per-band steps, round them to the coded grid and recompute the
// steps (bitallo3.cpp:1892-2169).  ms selects the mid/side flavour of the rounding rules.  Every per-band step
// is independent, so bands are dealt over the lanes; the band maximum is a group reduction.
HMP3_FN int long_scale_factors(const EncTables *T, LongRate *L, bool ms) {
    int gmin_all = 999;
    int gtop = -1;
    if (ms && L->hf_quant) gtop = L->gsf_hf;
    HMP3_SYNC();
    for (int ch = 0; ch < L->nchan; ch++) {
        const int nsf = T->cfg.nsf[ch];
        if (!ms) gtop = L->gsf_hf_ch[ch];
        {
            int top = -1;
            HMP3_FOR_LANES(i, nsf) {
                L->gsf[ch][i] = imax_(L->gsf[ch][i], L->gmin[ch][i]);
                L->active[ch][i] = 0;
                if (L->gsf[ch][i] < L->gzero[ch][i]) {
                    L->active[ch][i] = -1;
                    top = imax_(top, L->gsf[ch][i]);
                }
            }
            gtop = imax_(gtop, gmax(top));
        }
        if (gtop < 0) {  // nothing to code in this channel
            int top = -1;
            HMP3_FOR_LANES(i, nsf) {
                L->sf[ch][i] = 0;
                L->gsf[ch][i] = L->gzero[ch][i];
                top = imax_(top, L->gsf[ch][i]);
            }
            gtop = imax_(gtop, gmax(top));
            L->preemp[ch] = 0;
            L->sf_scale[ch] = 0;
            L->G[ch] = gtop;
            gmin_all = imin_(gmin_all, 100);
            HMP3_SYNC();
            continue;  // note: the mid/side flavour carries gtop into the next channel here (bitallo3.cpp:2059-2074)
        }
        HMP3_FOR_LANES(i, nsf) L->sf[ch][i] = (gtop - L->gsf[ch][i]) & L->active[ch][i];
        HMP3_SYNC();
        long_pick_sf_mode(T, L, ch);
        const int scale = L->sf_scale[ch], pre = L->preemp[ch];
        const int dsf = scale == 0 ? 2 : 4;
        HMP3_FOR_LANES(i, nsf) {
            int sfv = L->sf[ch][i];
            if (scale == 0) {
                if (ms) {
                    if (L->active[ch][i]) {
                        if ((L->gzero[ch][i] - L->gsf[ch][i]) < 5) sfv++;
                        else if ((i < 11) && (L->noise[ch][i] > L->nt[ch][i])) sfv++;
                        sfv &= (~1);
                    }
                } else {
                    if ((i < 11) && (L->noise[ch][i] > L->nt[ch][i])) sfv++;
                    sfv &= (~1);
                }
            } else if (!(ms && !L->active[ch][i])) {
                int s = sfv & (~3);
                int d = sfv - s;
                int dN = L->noise[ch][i] - L->nt[ch][i] + 150 * d;
                if (dN > noise_gap_limit(i)) s = s + 4;
                else if (ms && (L->gzero[ch][i] - L->gsf[ch][i] - d) < 5) s = s + 4;
                sfv = ms ? s : (s & L->active[ch][i]);
            }
            const int hi = sf_upper(scale, pre, i), lo = sf_lower(scale, pre, i);
            if (sfv > hi) sfv = hi;
            else if (sfv < lo) sfv = lo;
            L->sf[ch][i] = sfv;
            if (L->active[ch][i]) {
                L->gsf[ch][i] = gtop - sfv;
                if (L->gsf[ch][i] < 0) {
                    L->gsf[ch][i] += dsf;
                    L->sf[ch][i] -= dsf;
                }
                if (L->gsf[ch][i] >= L->gzero[ch][i]) {
                    L->gsf[ch][i] = L->gzero[ch][i] + 5;
                    L->sf[ch][i] = sf_lower(scale, pre, i);
                }
            }
        }
        L->G[ch] = gtop;
        gmin_all = imin_(gmin_all, gtop);
        if (ms) gtop = -1;
        HMP3_SYNC();
    }
    return gmin_all;
}

// try coarser steps on the low bands while the measured noise stays under target (bitallo3.cpp:1348-1399)
HMP3_FN void long_coarsen_low_bands(const EncTables *T, LongRate *L, const float *xr) {
#if HMP3_COOP
    // the low bands are short (4..16 lines) and independent: one (channel, band) per lane, each lane walking its
    // candidates with the plain sequential noise measurement
    HMP3_SYNC();
    for (int it = HMP3_LANE; it < 32; it += HMP3_W) {
        const int ch = it >> 4, i = it & 15;
        if (ch >= L->nchan || i >= imin_(13, T->cfg.nsf[ch])) continue;
        if (!(L->active[ch][i] && (L->gsf[ch][i] < (L->gzero[ch][i] - 5)))) continue;
        const int sdelta = 2 * (1 + L->sf_scale[ch]);
        const int GG = L->G[ch];
        const int scale = L->sf_scale[ch], pre = L->preemp[ch];
        const float *y34 = L->x34[ch] + T->startBand_l[i];
        const float *y = xr + 576 * ch + T->startBand_l[i];
        const int n = T->nBand_l[i];
        int smin = L->sf[ch][i];
        const int g0 = L->gzero[ch][i] - 4;
        int s = imin_(L->sf[ch][i] - sdelta, sf_upper(scale, pre, i));
        const int s0 = sf_lower(scale, pre, i);
        const int logn = T->log_cbw_l[i];
        const int nt = L->nt[ch][i];
        for (; s >= s0; s -= sdelta) {
            const int g = GG - s;
            if (g >= g0) break;
            int nz = band_noise_seq(T, y34, y, g, n, logn);
            if (nz <= nt) {
                L->noise[ch][i] = nz;
                smin = s;
            }
        }
        L->sf[ch][i] = smin;
        L->gsf[ch][i] = imax_(GG - smin, 0);
    }
    HMP3_SYNC();
#else
    for (int ch = 0; ch < L->nchan; ch++) {
        const int sdelta = 2 * (1 + L->sf_scale[ch]);
        const int GG = L->G[ch];
        const float *y34 = L->x34[ch];
        const float *y = xr + 576 * ch;
        const int m = imin_(13, T->cfg.nsf[ch]);
        const int scale = L->sf_scale[ch], pre = L->preemp[ch];
        for (int i = 0; i < m; i++) {
            const int n = T->nBand_l[i];
            if (L->active[ch][i] && (L->gsf[ch][i] < (L->gzero[ch][i] - 5))) {
                int smin = L->sf[ch][i];
                const int g0 = L->gzero[ch][i] - 4;
                int s = imin_(L->sf[ch][i] - sdelta, sf_upper(scale, pre, i));
                const int s0 = sf_lower(scale, pre, i);
                const int logn = T->log_cbw_l[i];
                for (; s >= s0; s -= sdelta) {
                    const int g = GG - s;
                    if (g >= g0) break;
                    int nz = band_noise(T, y34, y, g, n, logn);
                    if (nz <= L->nt[ch][i]) {
                        L->noise[ch][i] = nz;
                        smin = s;
                    }
                }
                L->sf[ch][i] = smin;
                L->gsf[ch][i] = imax_(GG - smin, 0);
            }
            y34 += n;
            y += n;
        }
    }
#endif
}

// re-fit the scale factor of bands whose largest quantised value is 1 or 2 (bitallo3.cpp:1471-1536)
HMP3_FN void long_refit_sparse_bands(const EncTables *T, LongRate *L, const float *xr, const QLine *ix) {
#if HMP3_COOP
    // Line-parallel: every lane squares the de-quantised value of every 32nd line of the bands to refit into the
    // stream's scratch row, then the band's owner adds them up in line order.  The other sum of the fit, the energy
    // of the band's spectrum in the same order, is the one the prepare pass already took (exx).
    constexpr int NS = (22 + HMP3_W - 1) / HMP3_W;
    const int lane = HMP3_LANE;
    float *dd = rate_scratch_row();
    HMP3_SYNC();
    for (int ch = 0; ch < L->nchan; ch++) {
        const int nb = T->cfg.nsf[ch];
        unsigned fm = 0;  // bit b = band b of the channel is refitted
        for (int sl = 0; sl < NS; sl++) {
            const int i = lane + HMP3_W * sl;
            const int m = i < nb ? L->ixmax[ch][i] : 0;
            fm |= gballot((m == 1) || (m == 2)) << ((HMP3_W * sl) & 31);
        }
        if (fm == 0) continue;
        const int nl = T->startBand_l[nb];
        const QLine *q = ix + 576 * ch;
        for (int k0 = 0; k0 < nl; k0 += HMP3_W) {
            const int k = k0 + lane;
            if (k < nl && ((fm >> T->line_band_l[k]) & 1u)) {
                const int v = q[k];
                const float d = (v < 256) ? T->ix43[v] : (float)(pow((double)v, (4.0 / 3.0)));
                dd[k] = d * d;
            }
        }
        HMP3_SYNC();
        const int gscale = L->G[ch] << 13;
        const int scale = L->sf_scale[ch], pre = L->preemp[ch];
        for (int i = lane; i < nb; i += HMP3_W) {
            if (!((fm >> i) & 1u)) continue;
            const int n = T->nBand_l[i], k0 = T->startBand_l[i];
            const float sqq = sum_seq(dd + k0, n, 0.0f);
            float sxx;
            if (i < L->n_exx[ch]) sxx = L->exx[ch][i];
            else {
                sxx = 0.0f;
                for (int k = 0; k < n; k++) sxx += xr[576 * ch + k0 + k] * xr[576 * ch + k0 + k];
            }
            const int t = 54 * mb_log(T, sxx / sqq) + (8 << 13);
            int s;
            if (scale == 0) s = ((gscale - t + (1 << 13)) & (~((1 << 14) - 1))) >> 13;
            else s = ((gscale - t + (1 << 14)) & (~((1 << 15) - 1))) >> 13;
            s = imin_(s, sf_upper(scale, pre, i));
            s = imax_(s, sf_lower(scale, pre, i));
            L->sf[ch][i] = s;
        }
        HMP3_SYNC();
    }
#else
    for (int ch = 0; ch < L->nchan; ch++) {
        const int gscale = L->G[ch] << 13;
        const int scale = L->sf_scale[ch], pre = L->preemp[ch];
        const float *y = xr + 576 * ch;
        const QLine *q = ix + 576 * ch;
        for (int i = 0; i < T->cfg.nsf[ch]; i++) {
            const int n = T->nBand_l[i];
            if ((L->ixmax[ch][i] == 1) || (L->ixmax[ch][i] == 2)) {
                int t = band_refit_gain(T, q, y, n);
                int s;
                if (scale == 0) s = ((gscale - t + (1 << 13)) & (~((1 << 14) - 1))) >> 13;
                else s = ((gscale - t + (1 << 14)) & (~((1 << 15) - 1))) >> 13;
                s = imin_(s, sf_upper(scale, pre, i));
                s = imax_(s, sf_lower(scale, pre, i));
                L->sf[ch][i] = s;
            }
            y += n;
            q += n;
        }
    }
#endif
}

// ------------------------------------------------------------------ quantise + count
HMP3_FN void long_quantise(const EncTables *T, LongRate *L, QLine *ix, bool tuned) {  // bitallo3.cpp:1540-1581
#if HMP3_COOP
    // Every lane quantises every 32nd line with its band's step.  The band maxima come out of the same pass: the
    // lanes of a chunk that hold lines of one band reduce their values among themselves (enc_init's line_seg_l is
    // that lane mask), a band that continues into the next chunk carries its running maximum along, and the first
    // lane of the band's last piece stores the result.
    HMP3_SYNC();
    const int lane = HMP3_LANE;
    for (int ch = 0; ch < L->nchan; ch++) {
        const float *x = L->x34[ch];
        QLine *q = ix + 576 * ch;
        const int nb = T->cfg.nsf[ch], nl = T->startBand_l[nb];
#if HMP3_W == 32
        {   // the channel's |x|^(3/4) row goes to the warp's scratch row first, all of it in flight at once (cp.async):
            // the loop below would otherwise wait for one global load per 32 lines, 18 times in a row
            float *row = rate_scratch_row();
            for (int k = 2 * lane; k < nl; k += 64)
                asm volatile("{ .reg .u64 a; cvta.to.shared.u64 a, %0; cp.async.ca.shared.global [a], [%1], 8; }" ::"l"(row + k), "l"(x + k));
            asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
            HMP3_SYNC();
            x = row;
        }
#endif
        float ig_own = 0.0f;  // W = 32: lane b holds the inverse step of band b
        if (HMP3_W == 32 && lane < nb) ig_own = T->igain34[L->gsf[ch][lane]];
        int carry_b = -1, carry_m = 0;
        for (int k0 = 0; k0 < nl; k0 += HMP3_W) {
            const int k = k0 + lane;
            const bool in = k < nl;
            const int b = in ? (int)T->line_band_l[k] : -2;
            float ig;
            if (HMP3_W == 32) ig = gshfl(ig_own, b & 31);
            else ig = in ? T->igain34[L->gsf[ch][b]] : 0.0f;
            int v = 0;
            unsigned seg = 1u << lane;
            bool first = false, last = false;
            if (in) {
                const float xv = x[k];
                if (tuned) {
                    float t = ig * xv + (0.5f - 0.4375f);
                    int c = (int)t;
                    if (c > 31) c = 31;
                    v = (int)(t - T->quantB_round[c]);
                } else v = (int)(ig * xv + (0.5f - 0.0946f));
                q[k] = (QLine)v;
                if (HMP3_W == 32) {
                    seg = T->line_seg_l[k];
                    const int f = T->line_segflag_l[k];
                    first = f & 1;
                    last = f & 2;
                } else {
                    const int s = T->startBand_l[b], e = s + T->nBand_l[b];
                    const int lo = imax_(s - k0, 0), hi = imin_(e - k0, HMP3_W);
                    seg = (((hi - lo) >= 32) ? 0xffffffffu : ((1u << (hi - lo)) - 1u)) << lo;
                    first = lane == lo;
                    last = e <= k0 + HMP3_W;
                }
            }
            int m = gmax_seg(v, seg);
            if (b == carry_b) m = imax_(m, carry_m);
            if (first && last) L->ixmax[ch][b] = imax_(m, 0);
            carry_b = gshfl(b, HMP3_W - 1);
            carry_m = gshfl(m, HMP3_W - 1);
        }
    }
    HMP3_SYNC();
#else
    for (int ch = 0; ch < L->nchan; ch++) {
        const float *x = L->x34[ch];
        QLine *q = ix + 576 * ch;
        for (int i = 0; i < T->cfg.nsf[ch]; i++) {
            const int n = T->nBand_l[i];
            L->ixmax[ch][i] = tuned ? quant_tuned(T, x, q, L->gsf[ch][i], n, false, 0.0f)
                                    : quant_plain(T, x, q, L->gsf[ch][i], n);
            x += n;
            q += n;
        }
    }
#endif
}
// drop isolated single-valued quads from the top, at most level/16 of them (bitallo3.cpp:1657-1687)
HMP3_FN void sparsify_quads(QLine *q, int n, int level) {
    int c = 0;
    for (int i = 0; i < n; i++) c += q[i];
    c = (level * c) >> 4;
    if (c <= 0) return;
    int dropped = 0;
    for (int i = n - 4; i >= 0; i -= 4)
        if (q[i] + q[i + 1] + q[i + 2] + q[i + 3] == 1) {
            q[i] = q[i + 1] = q[i + 2] = q[i + 3] = 0;
            if (++dropped >= c) break;
        }
}
HMP3_FN void long_quantise_hf(const EncTables *T, LongRate *L, QLine *ix, bool ms) {  // bitallo3.cpp:1690-1736
    const int b = T->startBand_l[21], n = T->nBand_l[21];
    if (ms) {
        L->ixmax[0][21] = quant_tuned(T, L->x34[0] + b, ix + b, L->G[0], n, true, -.30f);
        return;
    }
    for (int ch = 0; ch < L->nchan; ch++)
        if (L->hf_quant_ch[ch]) {
            L->ixmax[ch][21] = quant_tuned(T, L->x34[ch] + b, ix + 576 * ch + b, L->G[ch], n, true, -.30f);
            sparsify_quads(ix + 576 * ch + b, n, 4);
        }
}
HMP3_FN int long_count(const EncTables *T, LongRate *L, const QLine *ix, const int *ncb) {  // bitallo3.cpp:1740-1779
    int bits = 0;
    for (int ch = 0; ch < L->nchan; ch++) {
        L->huff_bits[ch] = plan_regions_long(T, L->block_type, L->ixmax[ch], ix + 576 * ch, ncb[ch], &L->plan[ch]);
        bits += L->huff_bits[ch];
    }
    return bits;
}

// ------------------------------------------------------------------ bit-budget control loops
HMP3_FN int long_more_bits(const EncTables *T, LongRate *L, QLine *ix, int bits0, bool ms) {  // :2569-2721
    const int thres = L->min_target - (L->min_target >> 4);
    if (bits0 > thres) return bits0;
    int(*g)[22] = L->gsave;
    int bits = bits0;
    HMP3_SYNC();
    for (int ch = 0; ch < L->nchan; ch++)
        HMP3_FOR_LANES(i, T->cfg.nsf[ch]) g[ch][i] = L->gsf[ch][i];
    const int hf = T->cfg.hf_flag;
    for (int pass = 0; pass < 11; pass++) {
        const bool undo = (pass == 10) || (pass > 0 && bits >= thres);
        if (undo) {
            // finished stepping down; if that overshot the ceiling go back one step
            if (!(bits > L->max_target)) break;
            for (int ch = 0; ch < L->nchan; ch++)
                HMP3_FOR_LANES(i, T->cfg.nsf[ch]) L->gsf[ch][i] = g[ch][i] + 1;
        } else {
            for (int ch = 0; ch < L->nchan; ch++)
                HMP3_FOR_LANES(i, T->cfg.nsf[ch]) L->gsf[ch][i] = g[ch][i] = imax_(g[ch][i] - 1, L->gmin[ch][i]);
        }
        HMP3_SYNC();
        if (ms) {
            L->hf_quant = 0;
            L->ixmax[0][21] = 0;
            L->gsf_hf = -1;
            long_clear_hf_lines(T, ix, 1);
            if (hf) long_hf_decide(T, L, 0, true);
            long_scale_factors(T, L, true);
            long_quantise(T, L, ix, true);
            L->ixmax[0][21] = 0;
            if (L->hf_quant) long_quantise_hf(T, L, ix, true);
            bits = long_count(T, L, ix, T->cfg.nsf2);
        } else {
            if (hf & 2) {
                long_hf_reset_lr(L);
                long_hf_adjust_lr(T, L);
            }
            long_scale_factors(T, L, false);
            long_quantise(T, L, ix, true);
            if (L->hf_quant) long_quantise_hf(T, L, ix, false);
            bits = long_count(T, L, ix, T->cfg.nsf3);
        }
        if (undo) break;
    }
    return bits;
}

HMP3_FN int long_fewer_bits(const EncTables *T, LongRate *L, const float *xr, QLine *ix, int bits0) {  // :2814-2852
    const int f = (250 * 1024) / (L->active_lines + 10);
    int dN = imax_((f * (bits0 - L->max_target)) >> 10, 40);
    int bits = bits0;
    L->delta_mnr = 0;
    for (int k = 0; k < 10; k++) {
        L->delta_mnr += dN;
        HMP3_SYNC();
        for (int ch = 0; ch < L->nchan; ch++)
            HMP3_FOR_LANES(i, T->cfg.nsf[ch]) L->nt[ch][i] += dN;
        long_seek_actual(T, L, xr);
        long_scale_factors(T, L, false);
        long_quantise(T, L, ix, false);
        bits = long_count(T, L, ix, T->cfg.nsf2);
        if (bits <= L->max_target) break;
        dN = imax_((f * (bits - L->max_target)) >> 10, 40);
    }
    return bits;
}
HMP3_FN int long_cap_bits(const EncTables *T, LongRate *L, QLine *ix, bool per_channel) {  // :2725-2772
    int bits = 0;
    for (int k = 0; k < 100; k++) {
        for (int ch = 0; ch < L->nchan; ch++)
            if (!per_channel || L->huff_bits[ch] > kPart23Max)
                HMP3_FOR_LANES(i, T->cfg.nsf[ch]) L->gsf[ch][i] = imin_(127, L->gsf[ch][i] + 1);
        HMP3_SYNC();
        long_scale_factors(T, L, false);
        long_quantise(T, L, ix, false);
        bits = long_count(T, L, ix, T->cfg.nsf2);
        if (per_channel) {
            if ((L->huff_bits[0] <= kPart23Max) && (L->huff_bits[1] <= kPart23Max)) break;
        } else if (bits <= L->max_bits) break;
    }
    return bits;
}

// the allocation of one long granule; returns the bit count before the budget loops (the CBR
// feedback signal) (bitallo3.cpp:2948-3149)
HMP3_FN int long_allocate(const EncTables *T, LongRate *L, float *xr, QLine *ix, bool ms) {
    const int hf = T->cfg.hf_flag;
    if (hf) {
        if (ms) {
            L->hf_quant = 0;
            L->ixmax[0][21] = L->ixmax[1][21] = 0;
            L->gsf_hf = -1;
        } else long_hf_reset_lr(L);
        long_clear_hf_lines(T, ix, L->nchan);
    }
    long_seek_initial(T, L);
    long_seek_actual(T, L, xr);
    if (ms) {
        if (hf) long_hf_decide(T, L, 0, true);
    } else {
        long_trade_peaks(T, L);
        if (hf & 2) long_hf_adjust_lr(T, L);
    }
    long_scale_factors(T, L, ms);
    long_coarsen_low_bands(T, L, xr);
    long_quantise(T, L, ix, true);
    if (ms) L->ixmax[0][21] = 0;
    if (L->hf_quant) long_quantise_hf(T, L, ix, ms);
    int bits = long_count(T, L, ix, ms ? T->cfg.nsf2 : T->cfg.nsf3);
    const int bits0 = bits;
    if (bits < L->min_target && L->mnr < 2000) bits = long_more_bits(T, L, ix, bits, ms);
    if (ms) {
        L->hf_quant = 0;
        L->ixmax[0][21] = 0;
        L->gsf_hf = -1;
    } else if (hf) long_hf_reset_lr(L);
    const int nclr = ms ? 1 : L->nchan;
    if (bits > L->max_target) {
        long_clear_hf_lines(T, ix, nclr);
        bits = long_fewer_bits(T, L, xr, ix, bits);
    }
    if (bits > L->max_bits) {
        long_clear_hf_lines(T, ix, nclr);
        bits = long_cap_bits(T, L, ix, false);
    }
    if (bits > kPart23Max)
        for (int ch = 0; ch < L->nchan; ch++)
            if (L->huff_bits[ch] > kPart23Max) {
                long_clear_hf_lines(T, ix, nclr);
                bits = long_cap_bits(T, L, ix, true);
                break;
            }
    long_refit_sparse_bands(T, L, xr, ix);
    return bits0;
}

// CBR quality feedback (bitallo3.cpp:2897-2944)
HMP3_FN void long_mnr_feedback(const EncTables *T, LongRate *L, int active_lines, int bits, int block_type) {
    if (block_type == 2) return;
    if (L->calls > 10) {
        const float per_band = 150.0f / (0.20f * (active_lines + 10));
        int mnr = (int)(0.05 * per_band * (bits - L->target));
        int mnr2 = (int)(0.05 * per_band * imax_((bits - L->max_bits), 0));
        int dpool = imax_(L->target - bits, 0);
        int dbits = (((2044 + 8 * 5) - L->pool_bits) >> 4) - dpool;
        dbits = imax_(dbits, 0);
        dbits = imin_(dbits, 200);
        int mnrp = (int)(per_band * dbits);
        int mnr0 = (int)(0.2 * per_band * imax_((L->min_target - bits), 0));
        int dmnr = mnr + mnr2 + mnrp - mnr0;
        int maxd = imax_(L->mnr - T->cfg.initial_mnr, L->target >> 3);
        dmnr = imin_(dmnr, maxd);
        if (L->delta_mnr) dmnr = imax_(dmnr, (L->delta_mnr >> 1));
        L->mnr = L->mnr - dmnr;
        L->mnr = imin_(L->mnr, 2000);
        if (bits > (L->target + 2000)) L->mnr = imin_(L->mnr, T->cfg.initial_mnr);
    }
}

}  // namespace hmp3
