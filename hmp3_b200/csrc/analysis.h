// Phase A of the pipeline: everything upstream of the rate loop, expressed as per-work-item
// routines over flat buffers.  kernels.cu maps CUDA threads onto the work items; the test-only
// host simulator loops over them.  Granule indexing (SURVEY.md Appendix E, mp3enc.cpp:1045-1114):
//   P[j]  = polyphase output of PCM granule j (samples 576j .. 576j+575 plus 480 of history)
//   xr[K] = hybrid(P[K-3], P[K-2], block_type[K])     for encode granule K
//   attack detector of encode granule K looks at P[K-1]
#pragma once
#include "dsp_core.h"
#include "psy_core.h"

namespace hmp3 {

// Per-stream carry of the block-type decision scan (mp3enc.cpp:1325-1486; l3e.h:93-95).
struct SwitchState {
    int attack_hist[2][32];  // millibel energy history per channel, init 9000 (mp3enc.cpp:280-284)
    int short_next;          // short_flag_next of the previous granule
    int block_type;          // block type of the previous granule
};

HMP3_HD void switch_state_init(SwitchState *s) {
    for (int c = 0; c < 2; c++)
        for (int i = 0; i < 32; i++) s->attack_hist[c][i] = 9000;
    s->short_next = 0;
    s->block_type = 0;
}

// Per encode granule decision record.
struct GranuleInfo {
    int block_type, block_type_prev, short_cur, short_next;
};

// One polyphase time slot of PCM granule j (may be negative: all-zero history), channel ch.
// pcm: interleaved int16, num_samples per channel; samples outside [0,num_samples) read as zero.
// out: band-major [32][18] block of this granule-channel; frequency inversion is applied on store.
HMP3_HD void polyphase_item(const EncTables *T, const int16_t *pcm, long num_samples, int nch, int ch, long j, int t,
                            float *out) {
    const long newest = 576 * j + 32 * t + 31;
    auto fetch = [&](int i) -> float {
        long n = newest - i;
        if (n < 0 || n >= num_samples) return 0.0f;
        return (float)pcm[n * nch + ch];
    };
    float col[32];
    polyphase_slot(T, fetch, col, 1);
    const int nsb = T->cfg.nsb_hybrid;
#pragma unroll
    for (int sb = 0; sb < 32; sb++) out[18 * sb + t] = freq_inverted(sb, t, nsb) ? -col[sb] : col[sb];
}

// Same for a stream whose input went through the DC-blocking filter (float samples, `len` per channel).
HMP3_HD void polyphase_item_f(const EncTables *T, const float *pcmf, long len, int nch, int ch, long j, int t, float *out) {
    const long newest = 576 * j + 32 * t + 31;
    auto fetch = [&](int i) -> float {
        long n = newest - i;
        if (n < 0 || n >= len) return 0.0f;
        return pcmf[n * nch + ch];
    };
    float col[32];
    polyphase_slot(T, fetch, col, 1);
    const int nsb = T->cfg.nsb_hybrid;
#pragma unroll
    for (int sb = 0; sb < 32; sb++) out[18 * sb + t] = freq_inverted(sb, t, nsb) ? -col[sb] : col[sb];
}

// Block-type scan step for encode granule K (channels share the decision).  e_new[ch][9] are the
// attack energies of P[K-1].
HMP3_HD GranuleInfo switch_step(const EncTables *T, SwitchState *s, const int *e_new0, const int *e_new1) {
    if (T->cfg.allocator == 1) {  // the CBitAllo1 drivers never select a block type: always long (mp3enc.cpp:1236-1321)
        GranuleInfo z;
        z.block_type = z.block_type_prev = z.short_cur = z.short_next = 0;
        return z;
    }
    const int mpeg2 = (T->cfg.h_id == 0);
    const int nch = T->cfg.nchan;
    int flag = 0;
    for (int c = 0; c < nch; c++) {
        int *h = s->attack_hist[c];
        const int *en = c ? e_new1 : e_new0;
        for (int i = 0; i < 23; i++) h[i] = h[i + 9];
        for (int i = 0; i < 9; i++) h[23 + i] = en[i];
        int m = attack_measure(h, s->short_next, mpeg2);
        if (m > T->cfg.short_block_threshold) flag = 1;
    }
    GranuleInfo g;
    g.short_next = flag;
    g.short_cur = s->short_next;
    g.block_type_prev = s->block_type;
    g.block_type = block_type_rule(g.block_type_prev, g.short_cur, g.short_next);
    s->short_next = flag;
    s->block_type = g.block_type;
    return g;
}

// Hybrid transform of one sub-band of encode granule K (before alias reduction).
// prev/cur: band-major inverted polyphase granules P[K-3], P[K-2]; xr: 576 output lines.
HMP3_HD void hybrid_item(const EncTables *T, const float *prev, const float *cur, int bt, int sb, float *xr) {
    const int nsb = T->cfg.nsb_hybrid;
    if (bt != 2) {
        if (sb < nsb) hybrid_long_band(T, prev + 18 * sb, cur + 18 * sb, bt, xr + 18 * sb);
        else
            for (int k = 0; k < 18; k++) xr[18 * sb + k] = 0.0f;
    } else {
        if (sb < nsb) hybrid_short_band(T, prev + 18 * sb, cur + 18 * sb, xr + 6 * sb);
        else
            for (int w = 0; w < 3; w++)
                for (int k = 0; k < 6; k++) xr[192 * w + 6 * sb + k] = 0.0f;
    }
}
// Alias reduction work item: boundary above sub-band sb (long block types only).
HMP3_HD void alias_item(const EncTables *T, int sb, float *xr) {
    const int nsb = T->cfg.nsb_hybrid;
    if (sb < nsb - 1) alias_boundary(T, xr + 18 * sb, false);
    else if (sb == nsb - 1) alias_boundary(T, xr + 18 * sb, true);
}

}  // namespace hmp3
