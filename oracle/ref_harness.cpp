// TEST INFRASTRUCTURE ONLY -- never linked into, loaded by, or called from the product path.
//
// Tap harness around the UNMODIFIED reference encoder (maikmerten/hmp3, Helix 5.2.4).  It is
// compiled by oracle/Makefile together with the reference's own objects into
// oracle/_ref/libhmp3ref.so.  It drives the reference through its public entry points
// (CMp3Enc::L3_audio_encode_init / L3_audio_encode, hmp3/src/pub/mp3enc.h:88-98) and records
// every stage boundary the parity tests compare against:
//   * polyphase output       CMp3Enc::sample[ch][slot][576]        (pub/mp3enc.h:243)
//   * rate-loop inputs       xr, sig_mask, bit budgets             (CBitAllo::BitAllo args, pub/bitallo.h:78-84)
//   * rate-loop outputs      SCALEFACT, GR, ix, signx              (same call)
//   * M/S correlation        CBitAllo::ms_correlation2 return      (pub/bitallo.h:88)
//   * side info / reservoir  CMp3Enc::side_info, byte_pool, ...    (pub/mp3enc.h:260-291)
//   * the emitted bytes
// Private members are reached with `#define private public`; the rate-loop taps use a forwarding
// proxy installed in CMp3Enc::BitAllo after init, so no reference code is altered or restated.
//
// The reference is NOT re-entrant (file-scope statics): one live encoder per process.

// The reference reads at least one heap member before writing it (CBitAllo1::alpha_nmr: bitallo1.cpp:318 after the early
// return of fnc_noise_seek at :1312-1317; found with MALLOC_PERTURB_ -- the dual-channel / intensity-stereo output
// changes with the previous contents of the heap).  In a fresh process (the CLI) such memory is zero; inside a
// long-lived test process it is not.  Every allocation of this library is therefore zero-filled (-Bsymbolic binds the
// reference's `new` to these), which pins the oracle to the fresh-process behaviour.
#include <cstdlib>
#include <new>
void *operator new(std::size_t n) {
    void *p = std::calloc(1, n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void *operator new[](std::size_t n) {
    void *p = std::calloc(1, n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete(void *p) noexcept { std::free(p); }
void operator delete[](void *p) noexcept { std::free(p); }
void operator delete(void *p, std::size_t) noexcept { std::free(p); }
void operator delete[](void *p, std::size_t) noexcept { std::free(p); }

#include <cstring>
#include <cstdlib>
#include <cstdint>

#define private public
#define protected public
#include "mp3enc.h"
#include "bitallo3.h"
#include "bitallos.h"
#undef private
#undef protected

extern "C" {
#include "xhead.h"
#include "srcc.h"
}

namespace {

#pragma pack(push, 4)
struct RefGranule {                 // one granule (all channels), filled across one BitAllo call
    int32_t valid;
    int32_t nchan;
    int32_t ms_flag;                // ms_flag_arg handed to BitAllo
    int32_t min_bits, target_bits, max_bits, bit_pool;
    int32_t block_type, block_type_prev, short_flag_current, short_flag_next;
    int32_t ms_corr;                // last ms_correlation2() result seen for this granule (or 0)
    float   xr[2][576];             // rate-loop input spectrum (before in-place abs / M-S)
    float   sigmask[2][36][2];      // {sig,mask}
    int32_t gr[2][27];              // GR after packing (part2_3_length etc. filled)
    int32_t sf_l[2][23];
    int32_t sf_s[2][3][13];
    int32_t ix[2][576];
    uint8_t signx[2][576];
    int32_t scfsi[2];
    int32_t mode_ext;
};
struct RefCall {                    // one L3_audio_encode call (MPEG-1: 2 granules of one frame;
    int32_t out_bytes;              //  MPEG-2: 2 one-granule frames)
    int32_t byte_pool[2], byte_min[2], byte_max[2];
    int32_t attack_buf[2][32];
    float   sbt[2][2][576];         // [new granule 0/1][ch] polyphase output written by this call
    float   ecsave[2][64];          // psycho pre-echo memory after the call ([ch][0][..])
    RefGranule g[2];
};
#pragma pack(pop)

static_assert(sizeof(GR) == 27 * sizeof(int), "GR layout");

class TapAllo;
CMp3Enc *g_enc = nullptr;
TapAllo *g_tap = nullptr;
RefCall *g_call = nullptr;          // current trace target (may be null)
int g_granule = 0;                  // granule slot inside the current call
int g_ms_corr[2];
int g_ms_n = 0;

class TapAllo : public CBitAllo {
  public:
    CBitAllo *inner;
    explicit TapAllo(CBitAllo *p) : inner(p) {}
    ~TapAllo() override { delete inner; }
    int BitAlloInit(BA_CONTROL &bac) override { return inner->BitAlloInit(bac); }
    void ba_out_stats() override {}
    int ms_correlation2(float x[2][576], int bt) override {
        int r = inner->ms_correlation2(x, bt);
        if (g_ms_n < 2) g_ms_corr[g_ms_n] = r;
        g_ms_n++;
        return r;
    }
    void BitAllo(float xr[][576], SIG_MASK sm[][36], int ch, int nchan, int min_bits, int target_bits,
                 int max_bits, int bit_pool, SCALEFACT sf_out[], GR gr_data[], int ix[][576],
                 unsigned char signx[][576], int ms_flag) override {
        RefGranule *t = (g_call && g_granule < 2) ? &g_call->g[g_granule] : nullptr;
        if (t) {
            memset(t, 0, sizeof(*t));
            t->valid = 1;
            t->nchan = nchan;
            t->ms_flag = ms_flag;
            t->min_bits = min_bits;
            t->target_bits = target_bits;
            t->max_bits = max_bits;
            t->bit_pool = bit_pool;
            t->block_type = gr_data[0].block_type;
            t->block_type_prev = gr_data[0].block_type_prev;
            t->short_flag_current = gr_data[0].short_flag_current;
            t->short_flag_next = gr_data[0].short_flag_next;
            for (int c = 0; c < nchan; c++) {
                memcpy(t->xr[c], xr[c], sizeof(float) * 576);
                for (int i = 0; i < 36; i++) {
                    t->sigmask[c][i][0] = sm[c][i].sig;
                    t->sigmask[c][i][1] = sm[c][i].mask;
                }
            }
        }
        inner->BitAllo(xr, sm, ch, nchan, min_bits, target_bits, max_bits, bit_pool, sf_out, gr_data, ix,
                       signx, ms_flag);
        if (t) {
            for (int c = 0; c < nchan; c++) {
                memcpy(t->sf_l[c], sf_out[c].l, sizeof(int) * 23);
                memcpy(t->sf_s[c], sf_out[c].s, sizeof(int) * 39);
                memcpy(t->ix[c], ix[c], sizeof(int) * 576);
                memcpy(t->signx[c], signx[c], 576);
            }
        }
        g_granule++;
    }
};

}  // namespace

extern "C" {

int ref_trace_sizes(int *granule_bytes, int *call_bytes) {
    *granule_bytes = (int)sizeof(RefGranule);
    *call_bytes = (int)sizeof(RefCall);
    return 0;
}

void ref_close() {
    if (g_enc) {
        delete g_enc;   // deletes the proxy, which deletes the real allocator
        g_enc = nullptr;
        g_tap = nullptr;
    }
}

// E_CONTROL is passed as its raw 45-int image (pub/encapp.h:42-72).
int ref_init(const int *ec_words) {
    ref_close();
    E_CONTROL ec;
    static_assert(sizeof(E_CONTROL) == 45 * sizeof(int), "E_CONTROL layout");
    memcpy(&ec, ec_words, sizeof(ec));
    g_enc = new CMp3Enc;
    int r = g_enc->L3_audio_encode_init(&ec);
    if (r == 0) {
        ref_close();
        return 0;
    }
    g_tap = new TapAllo(g_enc->BitAllo);
    g_enc->BitAllo = g_tap;
    return r;
}

// Integer facts the boundary tests compare (resolved config).
int ref_info(int *out, int n) {
    if (!g_enc) return 0;
    int v[] = {g_enc->nchan, g_enc->h_id, g_enc->sr_index, g_enc->nband, g_enc->band_limit, g_enc->nsb,
               g_enc->nsb_limit, g_enc->nsb_limitMS[0], g_enc->nsb_limitMS[1], g_enc->AveTargetBits,
               g_enc->framebytes, g_enc->main_framebytes, g_enc->side_bytes, g_enc->remainder,
               g_enc->divisor, g_enc->ms_flag, g_enc->is_flag, g_enc->iL3_audio_encode_function,
               g_enc->iencode_function, g_enc->ivbr_min, g_enc->ivbr_max, g_enc->vbr_pool_target,
               g_enc->short_block_threshold, g_enc->h.mode, g_enc->h.br_index, g_enc->totbitrate,
               g_enc->samprate, g_enc->band_limit_stereo, g_enc->sf_bit_max, g_enc->nsf_stereo,
               (int)g_enc->head[0], (int)g_enc->head[1], (int)g_enc->head[2], (int)g_enc->head[3],
               g_enc->ec_global.hf_flag, g_enc->fc2.select};
    int m = (int)(sizeof(v) / sizeof(v[0]));
    for (int i = 0; i < n && i < m; i++) out[i] = v[i];
    return m;
}

// Psychoacoustic tables generated by amod_initLong/Short (amodini2.c:743/587).
void ref_psy_tables(int *nsum_l, int *spd_l, float *w_l, int *nsum_s, int *spd_s, float *w_s) {
    memcpy(nsum_l, g_enc->nsum, sizeof(int) * 68);
    memcpy(spd_l, g_enc->spd_cntl, sizeof(int) * 2 * 65);
    memcpy(w_l, g_enc->w_spd, sizeof(float) * 2200);
    memcpy(nsum_s, g_enc->nsumShort, sizeof(int) * 68);
    memcpy(spd_s, g_enc->spd_cntlShort, sizeof(int) * 2 * 65);
    memcpy(w_s, g_enc->w_spdShort, sizeof(float) * 1000);
}

void ref_vbr_tables(int *main_fb, int *fb) {
    memcpy(main_fb, g_enc->vbr_main_framebytes, sizeof(int) * 16);
    memcpy(fb, g_enc->vbr_framebytes, sizeof(int) * 16);
}

// hybrid / alias / MDCT coefficient tables (hwin.c:49,58; emdct.c:70-76).
extern float win[4][36];
extern float csa[2][8];
struct MdctInit { float *w; float *w2; void *coef; };
MdctInit *mdct_init_addr_18();
MdctInit *mdct_init_addr_6();
void ref_xform_tables(float *win_out, float *csa_out, float *w18, float *w2_9, float *coef94, float *v6,
                      float *v2_3, float *coef87) {
    memcpy(win_out, win, sizeof(float) * 4 * 36);
    memcpy(csa_out, csa, sizeof(float) * 16);
    MdctInit *a = mdct_init_addr_18();
    memcpy(w18, a->w, sizeof(float) * 18);
    memcpy(w2_9, a->w2, sizeof(float) * 9);
    memcpy(coef94, a->coef, sizeof(float) * 36);
    MdctInit *b = mdct_init_addr_6();
    memcpy(v6, b->w, sizeof(float) * 6);
    memcpy(v2_3, b->w2, sizeof(float) * 3);
    memcpy(coef87, b->coef, sizeof(float));
}

// spd_smrLongEcho reads one element of a local array that it has not written when the partition count is odd
// (spdsmr.c:193, 283: stab[npart]); what it finds there is whatever earlier calls left on the stack.  In the CLI that
// is zero.  The harness zeroes the stack below itself before every encode call so that a call made from deep inside
// a Python process sees the same thing.
static void __attribute__((noinline)) scrub_stack() {
    volatile unsigned char pad[192 * 1024];
    for (unsigned i = 0; i < sizeof(pad); i++) pad[i] = 0;
}

// One encode call.  pcm = nchan*1152 floats, interleaved, scaled to +-32768 (pub/mp3enc.h:90-98).
int ref_encode(const float *pcm, unsigned char *out, void *trace) {
    scrub_stack();
    RefCall *t = (RefCall *)trace;
    g_call = t;
    g_granule = 0;
    g_ms_n = 0;
    if (t) memset(t, 0, sizeof(*t));
    int mpeg2 = (g_enc->h_id == 0);
    IN_OUT x = g_enc->L3_audio_encode(const_cast<float *>(pcm), out);
    if (t) {
        t->out_bytes = x.out_bytes;
        memcpy(t->attack_buf, g_enc->attack_buf, sizeof(t->attack_buf));
        int igrx = g_enc->igrx;  // post-call: the two slots written by this call are igrx, igrx+1
        for (int c = 0; c < g_enc->nchan; c++) {
            memcpy(t->sbt[0][c], g_enc->sample[c][(igrx + 0) & 3], sizeof(float) * 576);
            memcpy(t->sbt[1][c], g_enc->sample[c][(igrx + 1) & 3], sizeof(float) * 576);
            memcpy(t->ecsave[c], g_enc->ecsave[c][0], sizeof(float) * 64);
        }
        t->byte_pool[1] = g_enc->byte_pool;
        t->byte_min[1] = g_enc->byte_min;
        t->byte_max[1] = g_enc->byte_max;
        for (int g = 0; g < 2; g++) {
            RefGranule *q = &t->g[g];
            for (int c = 0; c < g_enc->nchan; c++) memcpy(q->gr[c], &g_enc->side_info.gr[g][c], sizeof(GR));
            q->scfsi[0] = g_enc->side_info.scfsi[0];
            q->scfsi[1] = g_enc->side_info.scfsi[1];
            // MPEG-1: one M/S decision per frame from two correlations; MPEG-2: one per granule.
            if (mpeg2) q->ms_corr = (g < g_ms_n) ? g_ms_corr[g] : 0;
            else q->ms_corr = (g_ms_n >= 2) ? g_ms_corr[g] : 0;
            q->mode_ext = g_enc->mode_ext_buf[(g_enc->side_p1 - (mpeg2 ? (2 - g) : 1)) & 31];
        }
    }
    g_call = nullptr;
    return x.out_bytes;
}

unsigned int ref_frames() { return g_enc ? g_enc->L3_audio_encode_get_frames() : 0; }

// One call of CMp3Enc::L3_audio_encode_Packet (mp3enc.cpp:3445-3440): standard bitstream (out may be NULL) plus the
// reformatted packet(s) of this call.  Returns out_bytes.
int ref_encode_packet(const float *pcm, unsigned char *out, unsigned char *packet, int *nbytes_out) {
    scrub_stack();
    g_call = nullptr;
    IN_OUT x = g_enc->L3_audio_encode_Packet(const_cast<float *>(pcm), out, packet, nbytes_out);
    return x.out_bytes;
}
// The info getters after the calls made so far: frames, bytes, bitrate (float), recent bitrate (float).
void ref_getters(int *frames_bytes, float *bitrates) {
    INT_PAIR fb = g_enc->L3_audio_encode_get_frames_bytes();
    frames_bytes[0] = fb.a;
    frames_bytes[1] = fb.b;
    bitrates[0] = g_enc->L3_audio_encode_get_bitrate_float();
    bitrates[1] = g_enc->L3_audio_encode_get_bitrate2_float();
}

// Whole-clip convenience used by tests and by bench.py's CPU baseline: float PCM in, raw MP3 frames out
// (no Xing/Info tag), following the CLI's flush protocol (test/tomp3.cpp:1015-1036): keep feeding zero
// frames until every started frame has been emitted.  Returns bytes written.
long ref_encode_clip(const int *ec_words, const float *pcm, long nsamples_per_ch, unsigned char *out,
                     long out_cap, void *traces, long max_calls, long *ncalls_out) {
    int bytes_in = ref_init(ec_words);
    if (!bytes_in) return -1;
    int nch = g_enc->nchan;
    long per_call = 1152;
    // CLI semantics (test/tomp3.cpp:908-942): 4 x bytes_in_init zero bytes are appended once at EOF and a call is
    // made while bytes_in_init bytes are buffered; bytes_in_init = 1153 sample frames (what Csrc::sr_convert_init
    // returns without rate conversion, srcc.cpp:185-187), each call consumes 1152.
    long ncalls = (nsamples_per_ch + 3 * 1153 + per_call) / per_call;
    int calls_per_frame_mult = (g_enc->h_id == 0) ? 2 : 1;
    float *buf = (float *)calloc((size_t)nch * 1152, sizeof(float));
    unsigned char tmp[16384];
    long total = 0, call = 0;
    RefCall *tr = (RefCall *)traces;
    for (;; call++) {
        if (call >= ncalls && ref_frames() >= (unsigned)(ncalls * calls_per_frame_mult)) break;
        if (call >= ncalls + 64) break;  // safety
        memset(buf, 0, sizeof(float) * nch * 1152);
        if (call < ncalls) {
            long off = call * per_call;
            long n = nsamples_per_ch - off;
            if (n > per_call) n = per_call;
            if (n > 0) memcpy(buf, pcm + off * nch, sizeof(float) * n * nch);
        }
        int ob = ref_encode(buf, tmp, (tr && call < max_calls) ? (void *)&tr[call] : nullptr);
        if (total + ob > out_cap) { free(buf); return -2; }
        memcpy(out + total, tmp, ob);
        total += ob;
    }
    free(buf);
    if (ncalls_out) *ncalls_out = call;
    return total;
}


// The reference's sample-rate converter on its own (Csrc, srcc.cpp / srccf.cpp): `ncalls` calls of sr_convert over the
// raw PCM bytes at `in` (the caller pads the buffer: a call reads more frames than it consumes).  out receives
// 1152 * target_channels floats per call, used[c] the bytes call c consumed.  Returns sr_convert_init's value
// (bytes to buffer per call; <= 0 = refused); *cutoff = the encoder cutoff frequency it reports.
int ref_src_convert(int source, int channels, int bits, int is_float, int target, int target_channels,
                    const unsigned char *in, int ncalls, float *out, int *used, int *cutoff) {
    Csrc c;
    int co = 0;
    const int minb = c.sr_convert_init(source, channels, bits, is_float, target, target_channels, &co);
    if (cutoff) *cutoff = co;
    if (minb <= 0) return minb;
    const int tch = target_channels < channels ? target_channels : channels;
    long off = 0;
    for (int k = 0; k < ncalls; k++) {
        IN_OUT x = c.sr_convert((unsigned char *)in + off, out + (long)k * 1152 * (tch < 1 ? 1 : tch));
        used[k] = x.in_bytes;
        off += x.in_bytes;
    }
    return minb;
}

}  // extern "C"
