// The serial stage as a per-stream STATE MACHINE.  The drivers of rate_driver.h / rate_long.h (rate_run_chunk ->
// encode_one_frame -> encode_frame_mpeg{1,2} -> granule_allocate -> long_allocate and its budget loops) are
// restated here as phases: every phase is a short stretch of code that runs to completion for one stream and
// names the phase that stream needs next; all values that cross a phase boundary live in the stream's RateCtl.
// The leaf routines (noise targets, step search, scale factors, quantiser, region planner, refit, records) are
// the ones of rate_long.h / rate_short.h / rate_driver.h, called in the reference's order, so the bytes are the
// same as the nested drivers' (bitallo3.cpp:484-678, 2948-3149, 2569-2852; mp3enc.cpp:1492-2027, 2106-2593).
//
// Why: the serial stage is bound by instruction supply.  With phases, the warps of an SM can pick
// streams that all need the SAME phase (kernels_rate_ph.cu), so that the SM walks one small piece of code at a
// time and its instruction cache serves all of them.  The host build runs the phases of one stream in sequence.
#pragma once
#include "rate_driver.h"

namespace hmp3 {

// A stream's slice of the chunk in hand (index 0 of the arrays == encode granule K0).
struct RateCtx {
    const EncTables *T;
    RateState *R;
    int K0, NG, ngran, ngran_real;
    const GranuleInfo *gi;
    float *xr;
    const SigMask *sm;
    PrepGranule *prep;
    const signed char *ms;
    PackGc *pack;
    FrameRec *frames;
};

HMP3_HD int ctl_granule(const EncTables *T, const RateCtl *c) {  // absolute index of the granule in hand
    return c->K + (T->cfg.h_id == 1 ? c->igr : c->sub);
}

// ---- RP_FRAME
HMP3_FN int phase_frame(const RateCtx *x) {
    const EncTables *T = x->T;
    RateState *R = x->R;
    RateCtl *c = &R->ctl;
    const EncConfig &C = T->cfg;
    const bool m1 = C.h_id == 1;
    if (c->sub == 0) {  // pair boundary (rate_run_chunk)
        int K = c->K;
        while (K + 1 < x->K0 + x->NG && (K + 1 >= x->ngran || R->finished || K < R->next_granule)) K += 2;
        c->K = K;
        if (K + 1 >= x->K0 + x->NG) return RP_IDLE;
    }
    // top of encode_one_frame
    const int nch = C.nchan;
    c->pad = 0;
    c->mf_bytes = 0;
    if (!C.vbr_flag) {
        R->padcount -= C.pad_remainder;
        if (R->padcount <= 0) {
            R->padcount += C.pad_divisor;
            c->pad = 1;
        }
        c->mf_bytes = C.main_framebytes + c->pad;
    }
    R->byte_pool = (int)(R->mf_tot - R->main_tot);
    if (C.vbr_flag) {
        R->byte_max = C.vbr_main_framebytes[C.ivbr_max] + R->byte_pool;
        R->byte_min = C.vbr_main_framebytes[C.ivbr_min] + R->byte_pool - C.reservoir_back;
    } else {
        R->byte_max = C.main_framebytes + c->pad + R->byte_pool;
        R->byte_min = R->byte_max - C.reservoir_back;
    }
    c->main_data_begin = R->byte_pool;
    c->frame_bits = 0;
    const int o = c->K - x->K0;
    if (m1) {  // top of encode_frame_mpeg1
        c->bit_pool = R->byte_pool << 2;
        const int bit_max = R->byte_max << 2, bit_min = R->byte_min << 2;
        c->sf_bits = nch * C.sf_bit_max;
        c->dba_max = 0;
        if (nch == 2) {
            c->ba_bit_max = bit_max - c->sf_bits;
            c->ba_bit_min = bit_min - c->sf_bits;
            c->dba_max = c->bit_pool >> 2;
            c->ba_min = c->ba_bit_min;
            c->ba_max = c->ba_bit_max + c->dba_max;
            c->target = C.ave_target_bits + C.ave_target_bits;
        } else {
            c->ba_bit_max = bit_max;
            if (c->ba_bit_max > 4095) c->ba_bit_max = 4095;
            c->ba_bit_min = bit_min;
            c->ba_bit_max -= C.sf_bit_max;
            c->ba_bit_min -= C.sf_bit_max;
            c->ba_min = c->ba_bit_min;
            c->ba_max = c->ba_bit_max;
            c->target = C.ave_target_bits;
        }
        GranuleIn g;
        g.info = x->gi[o];
        set_block_info(T, R, 0, &g);
        const int bt0 = g.info.block_type;
        g.info = x->gi[o + 1];
        set_block_info(T, R, 1, &g);
        c->short_frame = (bt0 == 2) | (g.info.block_type == 2);
        c->ms = x->ms[o];
        c->igr = 0;
    } else {  // top of encode_frame_mpeg2
        c->bit_pool = R->byte_pool << 3;
        int bit_max = (R->byte_max << 3), bit_min = (R->byte_min << 3);
        if (nch == 2 && R->byte_pool > 245) bit_min += 40;
        c->ba_bit_max = bit_max;
        if (c->ba_bit_max > 4095) c->ba_bit_max = 4095;
        c->ba_bit_min = bit_min;
        c->ba_bit_max -= nch * C.sf_bit_max;
        c->ba_bit_min -= nch * C.sf_bit_max;
        c->ba_min = c->ba_bit_min;
        c->ba_max = c->ba_bit_max;
        c->target = nch * C.ave_target_bits;
        GranuleIn g;
        g.info = x->gi[o + c->sub];
        set_block_info(T, R, c->sub, &g);
        c->short_frame = 0;
        c->ms = x->ms[o + c->sub];
        c->igr = c->sub;
    }
    return RP_GSTART;
}

// ---- RP_GSTART: granule_allocate up to the step search (bitallo3.cpp:484-560) + head of long_allocate
HMP3_FN int phase_gstart(const RateCtx *x) {
    const EncTables *T = x->T;
    RateState *R = x->R;
    RateCtl *c = &R->ctl;
    const EncConfig &C = T->cfg;
    const bool m1 = C.h_id == 1;
    const int nchan = C.nchan;
    const int igr = c->igr;
    const int o = ctl_granule(T, c) - x->K0;
    const int ms = m1 ? c->ms : (nchan == 2 ? c->ms : C.ms_flag);
    c->ga_ms = ms;
    const int min_bits = c->ba_min, target_bits = c->target, max_bits = c->ba_max, pool_bits = c->bit_pool;
    LongRate *L = &R->L;
    GrSide *gr = R->gr[igr];
    QLine *ix = &R->ix[0][0];
    unsigned *sg = &R->signx[0][0];
    const int bt = gr[0].block_type;
    const int init = C.initial_mnr;
    L->block_type = bt;
    L->calls++;
    L->delta_mnr = 0;
    if (bt == 1) {
        if (L->mnr > init) {
            L->mnr = (L->mnr + init) >> 1;
            L->mnr = imin_(L->mnr, init + 500);
        }
    } else if (bt == 3) {
        L->mnr = (L->mnr + init) >> 1;
        L->mnr = imin_(L->mnr, init + 500);
#if HMP3_COOP
        for (int k = HMP3_LANE; k < nchan * 576; k += HMP3_W) ix[k] = 0;
        HMP3_SYNC();
#else
        for (int k = 0; k < nchan * 576; k++) ix[k] = 0;
#endif
    }
    if (bt == 2) {
        c->gkind = 2;
        return RP_SHORT;
    }
    L->ms = ms;
    L->nchan = nchan;
    L->max_bits = imin_(4000 * nchan, max_bits);
    L->min_target = min_bits < 0 ? 0 : min_bits;
    L->target = target_bits;
    L->pool_bits = pool_bits;
    if (C.vbr_flag == 0) {
        L->pool_fraction = imin_(L->pool_fraction + 50, 614);
        if (bt != 0) L->pool_fraction = 0;
    }
    int tbits = ((L->pool_fraction * L->pool_bits) >> 10);
    if (C.vbr_flag == 0) tbits = imin_(tbits, imax_((2050 - 500) + init - L->mnr, 200));
    L->max_target = imin_(L->max_bits, L->target + tbits);
    if (L->mnr < -200) L->min_target = imax_(L->min_target, (3 * L->target) >> 2);
    L->max_target = imax_(L->min_target, L->max_target);
    L->min_target = imin_(L->min_target, L->max_target - 100);
    const SigMask *sm = x->sm + (long long)o * 72;
    PrepGranule *prep = x->prep + o;
    if (ms) long_startup_ms(T, L, sm, prep, sg);
    else long_startup_lr(T, L, sm, prep, sg);
    if (L->active_lines <= 0) {  // digital silence
        ScaleFac *sf_out = R->sf[igr];
        for (int ch = 0; ch < nchan; ch++) {
            GrSide *g = gr + ch;
            g->global_gain = 0;
            g->window_switching_flag = (bt != 0);
            g->block_type = bt;
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
            for (int j = 0; j < 21; j++) sf_out[ch].l[j] = 0;
        }
        c->gkind = 1;
        return RP_GFIN;
    }
    c->gkind = 0;
    // head of long_allocate
    const int hf = C.hf_flag;
    if (hf) {
        if (ms) {
            L->hf_quant = 0;
            L->ixmax[0][21] = L->ixmax[1][21] = 0;
            L->gsf_hf = -1;
        } else long_hf_reset_lr(L);
        long_clear_hf_lines(T, ix, L->nchan);
    }
    c->loop = RL_NONE;
    return RP_SEEK;
}

// ---- RP_SHORT
HMP3_FN int phase_short(const RateCtx *x) {
    const EncTables *T = x->T;
    RateState *R = x->R;
    RateCtl *c = &R->ctl;
    const EncConfig &C = T->cfg;
    LongRate *L = &R->L;
    const int igr = c->igr;
    const int o = ctl_granule(T, c) - x->K0;
    const int init = C.initial_mnr;
    int mnr0;
    if (C.vbr_flag == 0) {
        mnr0 = L->mnr - (imax_(L->mnr - init, 0) >> 1) - (imax_(L->mnr - init - 400, 0) >> 2);
        mnr0 = imax_(init + 400, mnr0);
    } else mnr0 = init + 400;
    short_granule(T, &R->cold->S, x->xr + (long long)o * 2 * 576, x->sm + (long long)o * 72, C.nchan, c->ba_min, c->target,
                  c->ba_max, c->bit_pool, R->sf[igr], R->gr[igr], &R->ix[0][0], &R->signx[0][0], c->ga_ms, mnr0);
    return RP_GFIN;
}

// ---- RP_SEEK: initial steps (first search of the granule) + the per-band step search
HMP3_FN int phase_seek(const RateCtx *x) {
    RateState *R = x->R;
    RateCtl *c = &R->ctl;
    const int o = ctl_granule(x->T, c) - x->K0;
    if (c->loop == RL_NONE) long_seek_initial(x->T, &R->L);
    long_seek_actual(x->T, &R->L, x->xr + (long long)o * 2 * 576);
    if (c->loop == RL_FEWER) return RP_SF;
    return c->ga_ms ? RP_SF : RP_TRADE;
}

// ---- RP_TRADE: left/right granules, between the step search and the scale factors (bitallo3.cpp:2990-3000)
HMP3_FN int phase_trade(const RateCtx *x) {
    const EncTables *T = x->T;
    LongRate *L = &x->R->L;
    long_trade_peaks(T, L);
    if (T->cfg.hf_flag & 2) long_hf_adjust_lr(T, L);
    return RP_SF;
}

// ---- RP_SF: scale factors -- of the first pass, or of one pass of a budget loop after that loop's move of the
// steps (long_more_bits bitallo3.cpp:2569-2721, long_fewer_bits :2814-2852, long_cap_bits :2725-2772)
HMP3_FN int phase_sf(const RateCtx *x) {
    const EncTables *T = x->T;
    RateState *R = x->R;
    RateCtl *c = &R->ctl;
    LongRate *L = &R->L;
    QLine *ix = &R->ix[0][0];
    const bool ms = c->ga_ms != 0;
    const int hf = T->cfg.hf_flag;
    if (c->loop == RL_NONE) {
        if (ms && hf) long_hf_decide(T, L, 0, true);
        long_scale_factors(T, L, ms);
        return RP_COARSE;
    }
    if (c->loop == RL_MORE) {
        int(*g)[22] = L->gsave;
        HMP3_SYNC();
        if (c->undo) {
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
        } else {
            if (hf & 2) {
                long_hf_reset_lr(L);
                long_hf_adjust_lr(T, L);
            }
            long_scale_factors(T, L, false);
        }
    } else if (c->loop == RL_FEWER) {
        long_scale_factors(T, L, false);
    } else {
        const bool per_channel = c->loop == RL_CAPCH;
        for (int ch = 0; ch < L->nchan; ch++)
            if (!per_channel || L->huff_bits[ch] > kPart23Max)
                HMP3_FOR_LANES(i, T->cfg.nsf[ch]) L->gsf[ch][i] = imin_(127, L->gsf[ch][i] + 1);
        HMP3_SYNC();
        long_scale_factors(T, L, false);
    }
    return RP_QUANT;
}

// ---- RP_COARSE: coarser steps on the low bands while the noise allows (first pass only)
HMP3_FN int phase_coarse(const RateCtx *x) {
    RateState *R = x->R;
    const int o = ctl_granule(x->T, &R->ctl) - x->K0;
    long_coarsen_low_bands(x->T, &R->L, x->xr + (long long)o * 2 * 576);
    return RP_QUANT;
}

HMP3_FN void fewer_step(const EncTables *T, LongRate *L, RateCtl *c) {  // head of one long_fewer_bits pass
    L->delta_mnr += c->dN;
    HMP3_SYNC();
    for (int ch = 0; ch < L->nchan; ch++)
        HMP3_FOR_LANES(i, T->cfg.nsf[ch]) L->nt[ch][i] += c->dN;
    HMP3_SYNC();
}
// ---- RP_QUANT: the quantiser pass
HMP3_FN int phase_quant(const RateCtx *x) {
    const EncTables *T = x->T;
    RateState *R = x->R;
    RateCtl *c = &R->ctl;
    LongRate *L = &R->L;
    QLine *ix = &R->ix[0][0];
    const bool ms = c->ga_ms != 0;
    if (c->loop == RL_NONE || c->loop == RL_MORE) {
        long_quantise(T, L, ix, true);
        if (ms) L->ixmax[0][21] = 0;
        if (L->hf_quant) long_quantise_hf(T, L, ix, ms);
    } else long_quantise(T, L, ix, false);
    return RP_COUNT;
}
// ---- RP_COUNT: region planning + bit count, then what long_allocate and its loops decide from the count
HMP3_FN int phase_count(const RateCtx *x) {
    const EncTables *T = x->T;
    RateState *R = x->R;
    RateCtl *c = &R->ctl;
    LongRate *L = &R->L;
    QLine *ix = &R->ix[0][0];
    const bool ms = c->ga_ms != 0;
    const int hf = T->cfg.hf_flag;
    int bits;
    const QLine *ixc = ix;
#if HMP3_COOP && HMP3_W == 32
    {   // the quantised lines of both channels (2 x 576 int16 = one scratch row) are read pair by pair by the planner
        // and the counters: bring them over in one go (cp.async) instead of waiting for one load per 64 lines
        char *row = (char *)rate_scratch_row();
        const char *src = (const char *)ix;
        for (int o = 16 * HMP3_LANE; o < 2304; o += 16 * 32)
            asm volatile("{ .reg .u64 a; cvta.to.shared.u64 a, %0; cp.async.cg.shared.global [a], [%1], 16; }" ::"l"(row + o), "l"(src + o));
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        HMP3_SYNC();
        ixc = (const QLine *)row;
    }
#endif
    if (c->loop == RL_NONE || c->loop == RL_MORE) bits = long_count(T, L, ixc, ms ? T->cfg.nsf2 : T->cfg.nsf3);
    else bits = long_count(T, L, ixc, T->cfg.nsf2);
    const int nclr = ms ? 1 : L->nchan;
    // the decisions, in the order of long_allocate (bitallo3.cpp:3050-3149)
    switch (c->loop) {
    case RL_NONE:
        c->bits0 = bits;
        if (bits < L->min_target && L->mnr < 2000) {
            c->thres = L->min_target - (L->min_target >> 4);
            if (!(bits > c->thres)) {
                HMP3_SYNC();
                for (int ch = 0; ch < L->nchan; ch++)
                    HMP3_FOR_LANES(i, T->cfg.nsf[ch]) L->gsave[ch][i] = L->gsf[ch][i];
                HMP3_SYNC();
                c->loop = RL_MORE;
                c->pass = 0;
                c->undo = 0;
                c->bits = bits;
                return RP_SF;
            }
        }
        break;
    case RL_MORE:
        if (!c->undo) {
            c->pass++;
            const int undo = (c->pass == 10) || (bits >= c->thres);
            if (!undo || bits > L->max_target) {
                c->undo = undo;
                return RP_SF;
            }
        }
        break;
    case RL_FEWER:
        if (bits > L->max_target && c->pass + 1 < 10) {
            c->pass++;
            c->dN = imax_((c->f * (bits - L->max_target)) >> 10, 40);
            fewer_step(T, L, c);
            return RP_SEEK;
        }
        goto after_fewer;
    case RL_CAP:
        c->pass++;
        if (bits > L->max_bits && c->pass < 100) return RP_SF;
        goto after_cap;
    default:  // RL_CAPCH
        c->pass++;
        if (!((L->huff_bits[0] <= kPart23Max) && (L->huff_bits[1] <= kPart23Max)) && c->pass < 100) return RP_SF;
        return RP_REFIT;
    }
    // after the first count / the more-bits loop
    if (ms) {
        L->hf_quant = 0;
        L->ixmax[0][21] = 0;
        L->gsf_hf = -1;
    } else if (hf) long_hf_reset_lr(L);
    if (bits > L->max_target) {
        long_clear_hf_lines(T, ix, nclr);
        c->loop = RL_FEWER;
        c->pass = 0;
        c->f = (250 * 1024) / (L->active_lines + 10);
        c->dN = imax_((c->f * (bits - L->max_target)) >> 10, 40);
        L->delta_mnr = 0;
        fewer_step(T, L, c);
        return RP_SEEK;
    }
after_fewer:
    if (bits > L->max_bits) {
        long_clear_hf_lines(T, ix, nclr);
        c->loop = RL_CAP;
        c->pass = 0;
        return RP_SF;
    }
after_cap:
    if (bits > kPart23Max)
        for (int ch = 0; ch < L->nchan; ch++)
            if (L->huff_bits[ch] > kPart23Max) {
                long_clear_hf_lines(T, ix, nclr);
                c->loop = RL_CAPCH;
                c->pass = 0;
                return RP_SF;
            }
    return RP_REFIT;
}

// ---- RP_REFIT: tail of long_allocate
HMP3_FN int phase_refit(const RateCtx *x) {
    RateState *R = x->R;
    const int o = ctl_granule(x->T, &R->ctl) - x->K0;
    long_refit_sparse_bands(x->T, &R->L, x->xr + (long long)o * 2 * 576, &R->ix[0][0]);
    return RP_GFIN;
}

// ---- RP_GFIN: tail of granule_allocate, then the frame driver's per-channel accounting and the records
HMP3_FN int phase_gfin(const RateCtx *x) {
    const EncTables *T = x->T;
    RateState *R = x->R;
    RateCtl *c = &R->ctl;
    const EncConfig &C = T->cfg;
    LongRate *L = &R->L;
    const bool m1 = C.h_id == 1;
    const int nch = C.nchan;
    const int igr = c->igr;
    GrSide *gr = R->gr[igr];
    if (c->gkind == 0) {
        const bool ms = c->ga_ms != 0;
        const int bt = L->block_type;
        if (C.vbr_flag == 0) long_mnr_feedback(T, L, L->active_lines, c->bits0, bt);
        ScaleFac *sf_out = R->sf[igr];
        for (int ch = 0; ch < nch; ch++) {
            const int sh = L->sf_scale[ch] == 0 ? 1 : 2;
            for (int i = 0; i < C.nsf[ch]; i++) L->sf[ch][i] >>= sh;
            if (L->preemp[ch])
                for (int i = 11; i < C.nsf[ch]; i++) L->sf[ch][i] -= sf_pre_amount(i);
            for (int i = 0; i < 21; i++) sf_out[ch].l[i] = L->sf[ch][i];
        }
        if (ms) {
            L->G[0] -= 2;
            L->G[1] -= 2;
        }
        for (int ch = 0; ch < nch; ch++) {
            GrSide *g = gr + ch;
            g->global_gain = imin_(L->G[ch] + (4 * 32 + 14), 255);
            g->window_switching_flag = (bt != 0);
            g->block_type = bt;
            g->mixed_block_flag = 0;
            g->preflag = L->preemp[ch];
            g->scalefac_scale = L->sf_scale[ch];
            g->aux_bits = L->huff_bits[ch];
            g->aux_not_null = L->huff_bits[ch];
            plan_to_side(T, &L->plan[ch], g);
        }
    }
    PackGc *pk = x->pack + (long long)(c->K - x->K0) * 2;
    if (m1) {  // encode_frame_mpeg1, after granule_allocate
        for (int ch = 0; ch < nch; ch++) {
            GrSide *g = &R->gr[igr][ch];
            int sfb = 0;
            g->scalefac_compress = 0;
            if (c->short_frame) {
                R->scfsi[ch] = 0;
                if (g->aux_not_null) g->scalefac_compress = plan_sf_mpeg1_plain(&R->sf[igr][ch], g->block_type, &sfb);
            } else {
                g->scalefac_compress =
                    plan_sf_mpeg1_scfsi(&R->sf[igr][ch], R->sf_save[ch], igr, &R->scfsi[ch], g->aux_not_null, &sfb);
            }
            const int bits = g->aux_not_null ? sfb + g->aux_bits : 0;
            if (nch == 2) {
                c->ba_min -= bits;
                c->ba_max -= bits;
            } else {
                c->ba_min += c->ba_bit_min + C.sf_bit_max - bits;
                c->ba_max += c->ba_bit_max + C.sf_bit_max - bits;
            }
            g->part2_3_length = bits;
            c->frame_bits += bits;
            record_gc(R, igr, ch, pk + (igr * nch + ch));
        }
        if (nch == 2) {
            c->ba_min += c->ba_bit_min + c->sf_bits;
            c->ba_max = c->ba_max - c->dba_max;
            c->ba_max += c->ba_bit_max + c->sf_bits;
        }
        if (igr == 0) {
            c->igr = 1;
            return RP_GSTART;
        }
        return RP_FEND;
    }
    // encode_frame_mpeg2, after granule_allocate
    for (int ch = 0; ch < nch; ch++) {
        GrSide *g = &R->gr[igr][ch];
        int bits = 0;
        g->scalefac_compress = 0;
        if (g->aux_not_null) {
            int sfb = 0;
            g->scalefac_compress = plan_sf_mpeg2(&R->sf[igr][ch], R->gr[igr][0].block_type, &sfb);
            bits = sfb + g->aux_bits;
        }
        g->part2_3_length = bits;
        c->frame_bits += bits;
        record_gc(R, igr, ch, pk + 2 * c->sub + ch);
    }
    return RP_FEND;
}

// ---- RP_FEND: bottom of encode_one_frame, then the pair bookkeeping of rate_run_chunk
HMP3_FN int phase_fend(const RateCtx *x) {
    const EncTables *T = x->T;
    RateState *R = x->R;
    RateCtl *c = &R->ctl;
    const EncConfig &C = T->cfg;
    const bool m1 = C.h_id == 1;
    FrameRec *frames = x->frames;
    FrameRec *fr = frames + R->frames;
    const int frame_bits = c->frame_bits;
    const int mode_ext = c->ms + c->ms + C.is_flag;
    int mf_bytes = c->mf_bytes;
    int bytes = (frame_bits + 7) >> 3;
    int ibr = 0;
    fr->main_start = R->mf_tot;
    if (C.vbr_flag) {
        const int bytes2 = bytes - R->byte_pool;
        const int bytes3 = bytes2 + C.vbr_pool_target;
        for (ibr = C.ivbr_min; ibr <= C.ivbr_max; ibr++)
            if (bytes2 <= C.vbr_main_framebytes[ibr]) break;
        bool grow = true;
        if (!m1) {
            const int side_dp = (R->frames - R->frames_done) & 31;
            grow = side_dp < 10;
            if (side_dp > 15) {
                if (side_dp > 24) R->byte_min = C.vbr_main_framebytes[C.ivbr_min] + R->byte_pool;
                else R->byte_min = C.vbr_main_framebytes[C.ivbr_min] + (R->byte_pool >> 4);
            }
        }
        if (grow)
            for (; ibr <= C.ivbr_max; ibr++)
                if (bytes3 < C.vbr_main_framebytes[ibr + 1]) break;
        if (ibr > C.ivbr_max) ibr = C.ivbr_max;
        mf_bytes = C.vbr_main_framebytes[ibr];
    }
    if (bytes < R->byte_min) bytes = R->byte_min;
    fr->mf_bytes = mf_bytes;
    fr->out_off = R->out_tot;
    R->out_tot += (unsigned)(4 + C.side_bytes + mf_bytes);
    fr->data_start = R->main_tot;
    fr->data_bits = frame_bits;
    fr->data_bytes = bytes;
    fr->granule0 = m1 ? c->K : c->K + c->sub;
    fr->ngr = (short)(m1 ? 2 : 1);
    fr->igr0 = (short)(m1 ? 0 : c->sub);
    fr->main_data_begin = (short)c->main_data_begin;
    fr->short_frame = (short)(m1 ? c->short_frame : 0);
    fr->scfsi[0] = (short)R->scfsi[0];
    fr->scfsi[1] = (short)R->scfsi[1];
    frame_header(T, fr->head, c->pad, mode_ext, ibr);
    R->main_tot += bytes;
    R->mf_tot += mf_bytes;
    R->frames++;
    while (R->frames_done < R->frames) {
        const FrameRec *f = frames + R->frames_done;
        if ((long long)R->main_tot - (long long)f->main_start < f->mf_bytes) break;
        R->frames_done++;
    }
    fr->done_after = R->frames_done;
    if (!m1 && c->sub == 0) {  // second frame of an MPEG-2 pair
        c->sub = 1;
        return RP_FRAME;
    }
    // end of the pair (rate_run_chunk)
    const int frames_real = m1 ? x->ngran_real / 2 : x->ngran_real;
    R->next_granule = c->K + 2;
    if (c->K + 2 >= x->ngran_real && R->frames_done >= frames_real) R->finished = 1;
    c->K += 2;
    c->sub = 0;
    return RP_FRAME;
}

HMP3_HD int rate_run_phase(const RateCtx *x, int phase) {
    switch (phase) {
    case RP_FRAME: return phase_frame(x);
    case RP_GSTART: return phase_gstart(x);
    case RP_SHORT: return phase_short(x);
    case RP_SEEK: return phase_seek(x);
    case RP_TRADE: return phase_trade(x);
    case RP_SF: return phase_sf(x);
    case RP_COARSE: return phase_coarse(x);
    case RP_QUANT: return phase_quant(x);
    case RP_COUNT: return phase_count(x);
    case RP_REFIT: return phase_refit(x);
    case RP_GFIN: return phase_gfin(x);
    case RP_FEND: return phase_fend(x);
    default: return RP_IDLE;
    }
}

// A stream enters a chunk at its pair boundary; RP_IDLE when the chunk holds nothing (more) for it.
HMP3_HD void rate_ctl_enter_chunk(const RateCtx *x) {
    RateCtl *c = &x->R->ctl;
    c->phase = RP_FRAME;
    c->K = x->K0;
    c->sub = 0;
}

// One stream, all phases of a chunk in sequence (host build; also the device's one-stream-per-warp form).
HMP3_FN void rate_run_chunk_phased(const EncTables *T, RateState *R, int K0, int NG, int ngran, int ngran_real,
                                   const GranuleInfo *gi, float *xr, const SigMask *sm, PrepGranule *prep,
                                   const signed char *ms, PackGc *pack, FrameRec *frames) {
    RateCtx x;
    x.T = T;
    x.R = R;
    x.K0 = K0;
    x.NG = NG;
    x.ngran = ngran;
    x.ngran_real = ngran_real;
    x.gi = gi;
    x.xr = xr;
    x.sm = sm;
    x.prep = prep;
    x.ms = ms;
    x.pack = pack;
    x.frames = frames;
    rate_ctl_enter_chunk(&x);
    int ph = RP_FRAME;
    while (ph != RP_IDLE) {
        ph = rate_run_phase(&x, ph);
        R->ctl.phase = ph;
    }
}

}  // namespace hmp3
