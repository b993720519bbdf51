// Phase B of the pipeline: the per-stream serial part.  For every encode granule, in order:
// psychoacoustic stage 2 (pre-echo memory), M/S decision with hysteresis, the rate loop (long or
// short), scale-factor and Huffman packing into the stream's main-data byte stream, and the bit
// reservoir / frame bookkeeping.  Follows CMp3Enc::encode_{jointB,singleB}[_MPEG2] and
// L3_audio_encode_[vbr_]MPEG{1,2} (mp3enc.cpp:1492-2027, 2106-2593), CBitAllo3::BitAllo
// (bitallo3.cpp:484-678) and l3pack.c.  Output is bit-exact with the reference.
#pragma once
#include "analysis.h"
#include "rate_short.h"
// The CBitAllo1 drivers (dual channel, intensity stereo) are compiled into the host build and into their own device
// kernel (k_rate_a1, kernels_rate_a1.cu): the default serial-stage kernel stays exactly the code of the common path.
#include "rate_allo1.h"
#if !HMP3_COOP || defined(HMP3_RATE_ALLOCATOR1)
#define HMP3_WITH_ALLO1 1
#else
#define HMP3_WITH_ALLO1 0
#endif

namespace hmp3 {

// HMP3_RATE_BARRIER (build option, experiments): the warps of a block start every frame together (block-wide
// barrier), so that they walk the code as a convoy and share instruction fetches; every warp of the block then runs
// the loop below with the same trip count, whether or not it has a stream or work left (`live`).
#if HMP3_COOP && defined(HMP3_RATE_BARRIER)
#define HMP3_FRAME_BARRIER() __syncthreads()
#else
#define HMP3_FRAME_BARRIER()
#endif
#if HMP3_COOP && defined(HMP3_RATE_BARRIER) && HMP3_RATE_BARRIER >= 2  // ... and the second granule of the call too
#define HMP3_GRANULE_BARRIER() __syncthreads()
#else
#define HMP3_GRANULE_BARRIER()
#endif
// One emitted frame, as recorded by the serial stage.  The serial stage only does the bit ACCOUNTING; the bits
// themselves (scale factors, Huffman codes, side info) are written afterwards by the packing pass
// (pack_frame: one warp per frame, all frames in parallel) and the final byte layout (header | side | slice of
// the main-data stream) by the assembly pass.
struct FrameRec {
    unsigned main_start;  // offset in the stream's main-data stream where this frame's SLOT begins
    int mf_bytes;         // bytes of main-data slot in this frame
    unsigned out_off;     // byte offset of this frame in the stream's output
    unsigned data_start;  // offset in the main-data stream where this frame's OWN data begins
    int data_bits;        // bits of scale factors + Huffman data of this frame
    int data_bytes;       // bytes the frame's data occupies (rounded up, zero-padded to the reservoir minimum)
    int granule0;         // first encode granule of the frame
    short ngr, igr0;      // granules in the frame (2 MPEG-1, 1 MPEG-2) and granule parity of the first
    short main_data_begin, short_frame;
    short scfsi[2];
    int done_after;       // frames complete (emitted) after the encode call that recorded this frame
    unsigned char head[4];
    unsigned char side[32];
};

// What the packing pass needs of one granule-channel.
struct PackGc {
    short ix[576];          // quantised magnitudes, transmission order
    unsigned sign[18];      // sign bit of line k = bit (k & 31) of word (k >> 5)
    unsigned char sf[64];   // long: l[0..22]; short: s[w][i] at 23 + 13*w + i
    GrSide gr;
};

// MSB-first bit writer appending to a byte buffer (same bit order as l3pack.c:98-153).
struct BitSink {
    unsigned char *p;
    unsigned long long acc;
    int nacc;
    long long total_bits;
};
HMP3_HD void sink_open(BitSink *b, unsigned char *dst) { b->p = dst; b->acc = 0; b->nacc = 0; b->total_bits = 0; }
HMP3_HD void sink_put(BitSink *b, unsigned x, int n) {
    if (n <= 0) return;
    b->acc = (b->acc << n) | (unsigned long long)(x & ((n >= 32) ? 0xFFFFFFFFu : ((1u << n) - 1u)));
    b->nacc += n;
    b->total_bits += n;
    while (b->nacc >= 8) {
        *b->p++ = (unsigned char)(b->acc >> (b->nacc - 8));
        b->nacc -= 8;
    }
}
HMP3_HD long long sink_close(BitSink *b) {  // pad to a byte boundary with zeros; returns bytes written
    if (b->nacc > 0) {
        *b->p++ = (unsigned char)(b->acc << (8 - b->nacc));
        b->nacc = 0;
    }
    return (b->total_bits + 7) >> 3;
}

// ---- the phase machine's per-stream record (rate_phased.h restates the drivers below as phases)
enum RatePhase {
    RP_FRAME = 0,  // pair / frame prologue: reservoir bounds, budget of the frame
    RP_GSTART,     // granule prologue: budget, noise targets (or digital silence)
    RP_SHORT,      // a short-block granule, whole (rare)
    RP_SEEK,       // initial steps + per-band step search
    RP_TRADE,      // left/right granules: peak trading, HF decision
    RP_SF,         // scale factors (after a budget loop's move of the steps, when in one)
    RP_COARSE,     // low-band coarsening (first pass)
    RP_QUANT,      // quantiser pass
    RP_COUNT,      // region planning + bit count, then the budget decision
    RP_REFIT,      // sparse-band refit
    RP_GFIN,       // CBR feedback, coded scale factors, side info, scale-factor plan, records
    RP_FEND,       // frame epilogue: frame size, reservoir, frame record
    RP_IDLE,       // nothing (left) to do in this chunk
    RP_NPHASES
};
enum RateLoop { RL_NONE = 0, RL_MORE, RL_FEWER, RL_CAP, RL_CAPCH };

// Everything of a stream that lives across phase boundaries (in RateState: persistent like the rest).
struct RateCtl {
    int phase;
    int K;       // first encode granule of the pair in hand (a pair = one encode call: 2 granules)
    int sub;     // MPEG-2: frame of the pair (0, 1) = granule parity; MPEG-1: 0
    int igr;     // MPEG-1: granule of the frame in hand
    // ---- frame scope (encode_one_frame / encode_frame_mpeg{1,2})
    int pad, mf_bytes, main_data_begin, frame_bits, short_frame, ms;
    int bit_pool, ba_bit_min, ba_bit_max, ba_min, ba_max, dba_max, target, sf_bits;
    // ---- granule scope (granule_allocate / long_allocate and its loops)
    int gkind;   // 0 long, 1 digital silence, 2 short
    int ga_ms;   // the ms flag the allocator was given
    int loop, pass, undo, bits, bits0, thres, dN, f;
};

// Persistent per-stream state of the serial stage.  The part that every granule touches (RateState, ~7 KB) and the
// part that only short-block granules and the CBitAllo1 configurations touch (RateCold, ~17 KB) live in separate
// arrays: the hot array of a full batch fits in L2 next to the granule inputs streaming through it, and the pipeline
// asks for it to stay there (access-policy window, pipeline.cu).
struct RateCold {
    ShortRate S;
    Allo1 A1;                 // state of the CBitAllo1 configurations (dual channel, intensity stereo)
};
struct RateState {
    LongRate L;
    RateCold *cold;
    alignas(16) QLine ix[2][576];  // quantised lines in transmission order (persist between granules)
    unsigned signx[2][18];    // sign bit of line k of a channel = bit (k & 31) of word (k >> 5); persists like ix
    GrSide gr[2][2];          // [granule][channel]
    ScaleFac sf[2][2];
    int scfsi[2];
    int sf_save[2][21];       // granule-0 scale factors for scfsi (l3pack.c:429)
    // reservoir / frame bookkeeping
    unsigned main_tot, mf_tot, out_tot;
    int padcount;
    int frames;               // frames recorded so far
    int frames_done;          // frames whose main-data slot is completely filled
    int byte_pool, byte_min, byte_max;
    int next_granule;         // encode granule to process next
    int finished;             // all real frames are complete
    RateCtl ctl;              // the stream's place in the phase machine (rate_phased.h)
};

HMP3_FN void rate_state_init(const EncTables *T, RateState *R, RateCold *cold) {
    unsigned char *p = (unsigned char *)R;
    for (unsigned i = 0; i < sizeof(RateState); i++) p[i] = 0;
    R->cold = cold;
    long_rate_init(T, &R->L);
    short_rate_init(&cold->S);
    allo1_init(T, &cold->A1);
    R->padcount = T->cfg.pad_divisor;
}

// ------------------------------------------------------------------ scale-factor packing (l3pack.c:157-933)
HMP3_HD int slen_for(int maxval, int cap) {
    int n = 1, s;
    maxval++;
    for (s = 0; s < cap; s++) {
        if (maxval <= n) break;
        n += n;
    }
    return s;
}
HMP3_HD int sfc_index(int slen1, int slen2) {  // ISO scalefac_compress from (slen1, slen2)
    return kSfcIndex[slen1][slen2];
}
HMP3_HD void sfc_slens(int sfc, int *s1, int *s2) {
    *s1 = kSfcSlen[sfc][0];
    *s2 = kSfcSlen[sfc][1];
}

// Scale-factor coding is split in two: plan_sf_* (serial stage) takes the decisions -- scalefac_compress, scfsi
// -- and returns the number of bits; write_sf_* (packing pass) emits the bits those decisions imply.
HMP3_HD int sf_byte(const unsigned char *sf, int i) { return sf[i]; }                       // long band i
HMP3_HD int sf_byte_s(const unsigned char *sf, int w, int i) { return sf[23 + 13 * w + i]; }  // short window w band i

// MPEG-1, frame contains a short block: no scfsi (l3pack.c:157-281)
HMP3_FN int plan_sf_mpeg1_plain(const ScaleFac *sf, int block_type, int *bits) {
    int m1 = 0, m2 = 0, s1, s2;
    if (block_type == 2) {
        for (int i = 0; i < 6; i++)
            for (int w = 0; w < 3; w++) m1 = imax_(m1, sf->s[w][i]);
        for (int i = 6; i < 12; i++)
            for (int w = 0; w < 3; w++) m2 = imax_(m2, sf->s[w][i]);
        int sfc = sfc_index(slen_for(m1, 4), slen_for(m2, 3));
        sfc_slens(sfc, &s1, &s2);
        *bits = 18 * s1 + 18 * s2;
        return sfc;
    }
    for (int i = 0; i < 11; i++) m1 = imax_(m1, sf->l[i]);
    for (int i = 11; i < 21; i++) m2 = imax_(m2, sf->l[i]);
    int sfc = sfc_index(slen_for(m1, 4), slen_for(m2, 3));
    sfc_slens(sfc, &s1, &s2);
    *bits = 11 * s1 + 10 * s2;
    return sfc;
}
HMP3_FN void write_sf_mpeg1_plain(BitSink *b, const unsigned char *sf, int block_type, int sfc) {
    int s1, s2;
    sfc_slens(sfc, &s1, &s2);
    if (block_type == 2) {
        for (int i = 0; i < 6; i++)
            for (int w = 0; w < 3; w++) sink_put(b, sf_byte_s(sf, w, i), s1);
        for (int i = 6; i < 12; i++)
            for (int w = 0; w < 3; w++) sink_put(b, sf_byte_s(sf, w, i), s2);
        return;
    }
    for (int i = 0; i < 11; i++) sink_put(b, sf_byte(sf, i), s1);
    for (int i = 11; i < 21; i++) sink_put(b, sf_byte(sf, i), s2);
}
// MPEG-1, all-long frame: granule 1 may share scale-factor groups with granule 0 (l3pack.c:421-557)
HMP3_FN int plan_sf_mpeg1_scfsi(const ScaleFac *sf, int *save /*[21]*/, int igr, int *scfsi_out, int not_null,
                                int *bits) {
    int scfsi = 0;
    if (igr == 0) {
        for (int i = 0; i < 21; i++) save[i] = sf->l[i];
    } else {
        for (int g = 0; g < 4; g++) {
            int t = 0;
            for (int i = kSfGroupEdge[g]; i < kSfGroupEdge[g + 1]; i++) t |= (save[i] - sf->l[i]);
            scfsi <<= 1;
            if (t == 0) scfsi |= 1;
        }
    }
    int sfc = 0;
    *bits = 0;
    if (not_null) {
        int m1 = 0, m2 = 0, s1, s2;
        for (int g = 0; g < 4; g++)
            if ((scfsi & (8 >> g)) == 0)
                for (int i = kSfGroupEdge[g]; i < kSfGroupEdge[g + 1]; i++) {
                    if (g < 2) m1 = imax_(m1, sf->l[i]);
                    else m2 = imax_(m2, sf->l[i]);
                }
        sfc = sfc_index(slen_for(m1, 4), slen_for(m2, 3));
        sfc_slens(sfc, &s1, &s2);
        for (int g = 0; g < 4; g++)
            if ((scfsi & (8 >> g)) == 0) *bits += (kSfGroupEdge[g + 1] - kSfGroupEdge[g]) * (g < 2 ? s1 : s2);
    }
    *scfsi_out = scfsi;
    return sfc;
}
HMP3_FN void write_sf_mpeg1_scfsi(BitSink *b, const unsigned char *sf, int scfsi, int sfc) {
    int s1, s2;
    sfc_slens(sfc, &s1, &s2);
    for (int g = 0; g < 4; g++)
        if ((scfsi & (8 >> g)) == 0)
            for (int i = kSfGroupEdge[g]; i < kSfGroupEdge[g + 1]; i++) sink_put(b, sf_byte(sf, i), g < 2 ? s1 : s2);
}
// MPEG-2 (LSF), no intensity stereo: four partitions with their own lengths (l3pack.c:561-933)
HMP3_FN int plan_sf_mpeg2(const ScaleFac *sf, int block_type, int *bits) {
    int m[4] = {0, 0, 0, 0}, sl[4];
    if (block_type == 2) {
        for (int w = 0; w < 3; w++)
            for (int i = 0; i < 12; i++) m[i / 3] = imax_(m[i / 3], sf->s[w][i]);
    } else {
        for (int g = 0; g < 4; g++)
            for (int i = kSfGroupEdge[g]; i < kSfGroupEdge[g + 1]; i++) m[g] = imax_(m[g], sf->l[i]);
    }
    sl[0] = slen_for(m[0], 4);
    sl[1] = slen_for(m[1], 4);
    sl[2] = slen_for(m[2], 3);
    sl[3] = slen_for(m[3], 3);
    if (block_type == 2) *bits = 9 * (sl[0] + sl[1] + sl[2] + sl[3]);
    else *bits = 6 * sl[0] + 5 * (sl[1] + sl[2] + sl[3]);
    return sl[3] + (sl[2] << 2) + ((sl[1] + 5 * sl[0]) << 4);
}
// MPEG-2, right channel of an intensity-stereo frame (long blocks): three partitions of seven bands, an "illegal"
// position (999 from the allocator = no intensity coding in that band) becomes the largest value of its partition's
// field, and a legal position must not collide with it (l3pack.c:576-712).  Rewrites sf->l for the packing pass.
HMP3_FN int plan_sf_mpeg2_is(ScaleFac *sf, int nsf_stereo, int *bits) {
    int m1 = 0, m2 = 0, m3 = 0, ip1 = 0, ip2 = 0, ip3 = 0, is2 = -1, is3 = -1;
    int i;
    for (i = 0; i < 7; i++) {
        if (sf->l[i] >= 999) ip1 = 1;
        else if (sf->l[i] > m1) m1 = sf->l[i];
    }
    for (; i < 14; i++) {
        if (sf->l[i] >= 999) {
            ip2 = 1;
            continue;
        }
        if (sf->l[i] > m2) m2 = sf->l[i];
        if (i < nsf_stereo) continue;
        if (sf->l[i] > is2) is2 = sf->l[i];
    }
    for (; i < 21; i++) {
        if (sf->l[i] >= 999) {
            ip3 = 1;
            continue;
        }
        if (sf->l[i] > m3) m3 = sf->l[i];
        if (i < nsf_stereo) continue;
        if (sf->l[i] > is3) is3 = sf->l[i];
    }
    int s1 = slen_for(m1, 4), s2 = slen_for(m2, 4), s3 = slen_for(m3, 3);
    if (is2 == ((1 << s2) - 1)) s2++;
    if (is3 == ((1 << s3) - 1)) s3++;
    if (ip1)
        for (i = 0; i < 7; i++)
            if (sf->l[i] >= 999) sf->l[i] = (1 << s1) - 1;
    if (ip2)
        for (i = 7; i < 14; i++)
            if (sf->l[i] >= 999) sf->l[i] = (1 << s2) - 1;
    if (ip3)
        for (i = 14; i < 21; i++)
            if (sf->l[i] >= 999) sf->l[i] = (1 << s3) - 1;
    *bits = 7 * (s1 + s2 + s3);
    const int sfc = s3 + 6 * s2 + 36 * s1;
    return sfc + sfc + 1;  // intensity_scale bit
}
HMP3_FN void write_sf_mpeg2_is(BitSink *b, const unsigned char *sf, int sfc) {
    const int v = sfc >> 1;
    const int s1 = v / 36, s2 = (v % 36) / 6, s3 = v % 6;
    for (int i = 0; i < 7; i++) sink_put(b, sf_byte(sf, i), s1);
    for (int i = 7; i < 14; i++) sink_put(b, sf_byte(sf, i), s2);
    for (int i = 14; i < 21; i++) sink_put(b, sf_byte(sf, i), s3);
}
HMP3_FN void write_sf_mpeg2(BitSink *b, const unsigned char *sf, int block_type, int sfc) {
    int sl[4];
    sl[3] = sfc & 3;
    sl[2] = (sfc >> 2) & 3;
    sl[1] = (sfc >> 4) % 5;
    sl[0] = (sfc >> 4) / 5;
    if (block_type == 2) {
        for (int i = 0; i < 12; i++)
            for (int w = 0; w < 3; w++) sink_put(b, sf_byte_s(sf, w, i), sl[i / 3]);
    } else {
        for (int g = 0; g < 4; g++)
            for (int i = kSfGroupEdge[g]; i < kSfGroupEdge[g + 1]; i++) sink_put(b, sf_byte(sf, i), sl[g]);
    }
}

// ------------------------------------------------------------------ Huffman packing (l3pack.c:946-1119)
HMP3_HD unsigned sign_bit(const unsigned *sign, int k) { return (sign[k >> 5] >> (k & 31)) & 1u; }

HMP3_FN void pack_huffman_seq(const EncTables *T, BitSink *b, const GrSide *g, const short *ix, const unsigned *sign) {
    int k0 = 0;  // first line of the current region
    for (int r = 0; r < 3; r++) {
        const int n = g->aux_nreg[r];
        const int t = g->table_select[r];
        const int book = T->huff_sel_book[t];
        if (book != 0) {
            const uint32_t *bk = T->huff_book[book];
            const int lb = T->huff_linbits[t];
            const bool esc = t >= 16;
            for (int j = 0; j < n; j++) {
                const int k = k0 + 2 * j;
                int x = ix[k], y = ix[k + 1];
                if (esc) {
                    if (x > 15) x = 15;
                    if (y > 15) y = 15;
                }
                const uint32_t e = bk[(x & 15) * 16 + (y & 15)];
                sink_put(b, e & 0xFFFFFFu, (int)(e >> 24));
                if (esc && x >= 15) sink_put(b, (unsigned)(ix[k] - 15), lb);
                if (x) sink_put(b, sign_bit(sign, k), 1);
                if (esc && y >= 15) sink_put(b, (unsigned)(ix[k + 1] - 15), lb);
                if (y) sink_put(b, sign_bit(sign, k + 1), 1);
            }
        }
        k0 += 2 * n;
    }
    const int nq = g->aux_nquads;
    const bool tabB = g->count1table_select == 1;
    for (int j = 0; j < nq; j++) {
        const int k = k0 + 4 * j;
        const unsigned x = (unsigned)((ix[k] << 3) + (ix[k + 1] << 2) + (ix[k + 2] << 1) + ix[k + 3]) & 15u;
        if (tabB) sink_put(b, x ^ 15u, 4);
        else sink_put(b, kQuadCodeA[x], kQuadLenA[x]);
        for (int q = 0; q < 4; q++)
            if (x & (8u >> q)) sink_put(b, sign_bit(sign, k + q), 1);
    }
}

#if HMP3_COOP
// (the packing pass runs one FULL warp per frame: its primitives are warp-wide, unlike the serial stage's)
#define PK_LANE ((int)(threadIdx.x & 31u))
#define PK_SYNC() __syncwarp()
// Warp-parallel Huffman packing: each lane builds the bit field of one pair (or quad), a warp scan of the
// field lengths gives its position, and the fields are OR-ed into a per-warp shared-memory bit buffer that
// is then streamed out in whole bytes.  Produces exactly the bits of the sequential writer above.
constexpr int kPackWords = 168;  // 5376 bits: part2_3_length can never exceed 4095 + the scale factors
__device__ __forceinline__ void bitbuf_or(unsigned *W, int pos, unsigned long long v, int len) {
    if (len <= 0) return;
    const unsigned long long V = v << (64 - len);
    const unsigned hi = (unsigned)(V >> 32), lo = (unsigned)V;
    const int w = pos >> 5, o = pos & 31;
    const unsigned w0 = hi >> o;
    const unsigned w1 = funnel_r(lo, hi, (unsigned)o);
    const unsigned w2 = funnel_r(0u, lo, (unsigned)o);
    if (w0) smem_or(W + w, w0);
    if (w1) smem_or(W + w + 1, w1);
    if (w2) smem_or(W + w + 2, w2);
}
__device__ __forceinline__ int warp_scan_excl(int v, int *total) {
    const int lane = PK_LANE;
    int s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = wshfl_up(s, d);
        if (lane >= d) s += t;
    }
    *total = wshfl(s, 31);
    return s - v;
}
HMP3_FN void pack_huffman(const EncTables *T, BitSink *b, const GrSide *g, const short *ix, const unsigned *sign) {
    __shared__ unsigned s_bits[kPackWarpsPerBlock][kPackWords];
    unsigned *W = s_bits[(threadIdx.x >> 5) % kPackWarpsPerBlock];
    const int lane = PK_LANE;
    PK_SYNC();
    const int p0 = b->nacc;  // bits pending in the writer (< 8)
    if (p0 + g->aux_bits > kPackWords * 32 - 160) {  // cannot happen within the part2_3 limits; stay correct anyway
        pack_huffman_seq(T, b, g, ix, sign);
        return;
    }
    const int nwords = imin_(kPackWords, ((p0 + g->aux_bits + 31) >> 5) + 3);
    for (int k = lane; k < nwords; k += 32) W[k] = 0;
    PK_SYNC();
    if (lane == 0 && p0 > 0) W[0] = ((unsigned)b->acc & ((1u << p0) - 1u)) << (32 - p0);
    PK_SYNC();
    int pos = p0;
    int k0 = 0;
    for (int r = 0; r < 3; r++) {
        const int n = g->aux_nreg[r];
        const int t = g->table_select[r];
        const int book = T->huff_sel_book[t];
        if (book != 0) {
            const uint32_t *bk = T->huff_book[book];
            const int lb = T->huff_linbits[t];
            const bool esc = t >= 16;
            for (int j0 = 0; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                unsigned long long v = 0;
                int len = 0;
                if (j < n) {
                    const int k = k0 + 2 * j;
                    const int x0 = ix[k], y0 = ix[k + 1];
                    int x = x0, y = y0;
                    if (esc) {
                        if (x > 15) x = 15;
                        if (y > 15) y = 15;
                    }
                    const uint32_t e = bk[(x & 15) * 16 + (y & 15)];
                    v = e & 0xFFFFFFu;
                    len = (int)(e >> 24);
                    if (esc && x >= 15 && lb > 0) { v = (v << lb) | ((unsigned)(x0 - 15) & ((1u << lb) - 1u)); len += lb; }
                    if (x) { v = (v << 1) | sign_bit(sign, k); len++; }
                    if (esc && y >= 15 && lb > 0) { v = (v << lb) | ((unsigned)(y0 - 15) & ((1u << lb) - 1u)); len += lb; }
                    if (y) { v = (v << 1) | sign_bit(sign, k + 1); len++; }
                }
                int tot;
                const int off = warp_scan_excl(len, &tot);
                bitbuf_or(W, pos + off, v, len);
                pos += tot;
            }
        }
        k0 += 2 * n;
    }
    {
        const int nq = g->aux_nquads;
        const bool tabB = g->count1table_select == 1;
        for (int j0 = 0; j0 < nq; j0 += 32) {
            const int j = j0 + lane;
            unsigned long long v = 0;
            int len = 0;
            if (j < nq) {
                const int k = k0 + 4 * j;
                const unsigned x = (unsigned)((ix[k] << 3) + (ix[k + 1] << 2) + (ix[k + 2] << 1) + ix[k + 3]) & 15u;
                if (tabB) { v = x ^ 15u; len = 4; }
                else { v = kQuadCodeA[x]; len = kQuadLenA[x]; }
                for (int q = 0; q < 4; q++)
                    if (x & (8u >> q)) { v = (v << 1) | sign_bit(sign, k + q); len++; }
            }
            int tot;
            const int off = warp_scan_excl(len, &tot);
            bitbuf_or(W, pos + off, v, len);
            pos += tot;
        }
    }
    PK_SYNC();
    const int nbytes = pos >> 3, rem = pos & 7;
    unsigned char *dst = b->p;
    for (int k = lane; k < nbytes; k += 32) dst[k] = (unsigned char)(W[k >> 2] >> (24 - 8 * (k & 3)));
    const unsigned last = (W[nbytes >> 2] >> (24 - 8 * (nbytes & 3))) & 0xffu;
    PK_SYNC();
    b->p = dst + nbytes;
    b->acc = rem ? (unsigned long long)(last >> (8 - rem)) : 0ull;
    b->nacc = rem;
    b->total_bits += pos - p0;
    PK_SYNC();
}

// One granule-channel of a frame into the frame's shared bit buffer W at bit position pos0: scale factors (one field
// per lane, in transmission order, as write_sf_* emit them) and then the Huffman data exactly as pack_huffman lays it
// out.  Returns the bits written.  The four granule-channels of a frame are packed by four warps at once: where
// each one starts follows from the part2_3_length of those before it, the fields are OR-ed in with shared-memory
// atomics, so neighbours may share a word.  k = index of the granule-channel in the frame.
HMP3_FN int pack_gc_bits(const EncTables *T, const FrameRec *fr, const PackGc *p, int k, unsigned *W, int pos0) {
    const GrSide *g = &p->gr;
    const int lane = PK_LANE;
    const int nch = T->cfg.nchan;
    const bool m1 = T->cfg.h_id == 1;
    int pos = pos0;
    {
        const int sfc = g->scalefac_compress;
        const bool shortb = g->block_type == 2;
        const bool is_right = !m1 && T->cfg.is_flag && (k % nch) == 1;
        const bool scfsi_mode = m1 && !fr->short_frame;
        const int scfsi = (scfsi_mode && (k / nch)) ? fr->scfsi[k % nch] : 0;
        int s1 = 0, s2 = 0, s3 = 0, sl0 = 0, sl1 = 0, sl2 = 0, sl3 = 0;
        if (m1) sfc_slens(sfc, &s1, &s2);
        else if (is_right) {
            const int v = sfc >> 1;
            s1 = v / 36;
            s2 = (v % 36) / 6;
            s3 = v % 6;
        } else {
            sl3 = sfc & 3;
            sl2 = (sfc >> 2) & 3;
            sl1 = (sfc >> 4) % 5;
            sl0 = (sfc >> 4) / 5;
        }
        for (int h = 0; h < 2; h++) {  // 21 fields (long) or 36 (short, in (band, window) order): two rounds of 32
            const int t = lane + 32 * h;
            unsigned v = 0;
            int l = 0;
            if (shortb && !is_right) {
                if (t < 36) {
                    const int i = t / 3, w = t - 3 * i;
                    v = (unsigned)sf_byte_s(p->sf, w, i);
                    const int q = i / 3;
                    l = m1 ? (i < 6 ? s1 : s2) : (q == 0 ? sl0 : (q == 1 ? sl1 : (q == 2 ? sl2 : sl3)));
                }
            } else if (t < 21) {
                v = (unsigned)sf_byte(p->sf, t);
                const int grp = t < 6 ? 0 : (t < 11 ? 1 : (t < 16 ? 2 : 3));  // kSfGroupEdge = 0, 6, 11, 16, 21
                if (is_right) l = t < 7 ? s1 : (t < 14 ? s2 : s3);
                else if (!m1) l = grp == 0 ? sl0 : (grp == 1 ? sl1 : (grp == 2 ? sl2 : sl3));
                else if (scfsi_mode) l = (scfsi & (8 >> grp)) ? 0 : (grp < 2 ? s1 : s2);
                else l = t < 11 ? s1 : s2;
            }
            if (l > 0) v &= (1u << l) - 1u;
            else v = 0;
            int tot;
            const int off = warp_scan_excl(l, &tot);
            bitbuf_or(W, pos + off, v, l);
            pos += tot;
            if (h == 0 && !shortb) break;  // long blocks: 21 fields, one round
        }
    }
    int k0 = 0;
    for (int r = 0; r < 3; r++) {
        const int n = g->aux_nreg[r];
        const int t = g->table_select[r];
        const int book = T->huff_sel_book[t];
        if (book != 0) {
            const uint32_t *bk = T->huff_book[book];
            const int lb = T->huff_linbits[t];
            const bool esc = t >= 16;
            const short *ix = p->ix;
            for (int j0 = 0; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                unsigned long long v = 0;
                int len = 0;
                if (j < n) {
                    const int kk = k0 + 2 * j;
                    const int x0 = ix[kk], y0 = ix[kk + 1];
                    int x = x0, y = y0;
                    if (esc) {
                        if (x > 15) x = 15;
                        if (y > 15) y = 15;
                    }
                    const uint32_t e = bk[(x & 15) * 16 + (y & 15)];
                    v = e & 0xFFFFFFu;
                    len = (int)(e >> 24);
                    if (esc && x >= 15 && lb > 0) { v = (v << lb) | ((unsigned)(x0 - 15) & ((1u << lb) - 1u)); len += lb; }
                    if (x) { v = (v << 1) | sign_bit(p->sign, kk); len++; }
                    if (esc && y >= 15 && lb > 0) { v = (v << lb) | ((unsigned)(y0 - 15) & ((1u << lb) - 1u)); len += lb; }
                    if (y) { v = (v << 1) | sign_bit(p->sign, kk + 1); len++; }
                }
                int tot;
                const int off = warp_scan_excl(len, &tot);
                bitbuf_or(W, pos + off, v, len);
                pos += tot;
            }
        }
        k0 += 2 * n;
    }
    {
        const int nq = g->aux_nquads;
        const bool tabB = g->count1table_select == 1;
        const short *ix = p->ix;
        for (int j0 = 0; j0 < nq; j0 += 32) {
            const int j = j0 + lane;
            unsigned long long v = 0;
            int len = 0;
            if (j < nq) {
                const int kk = k0 + 4 * j;
                const unsigned x = (unsigned)((ix[kk] << 3) + (ix[kk + 1] << 2) + (ix[kk + 2] << 1) + ix[kk + 3]) & 15u;
                if (tabB) { v = x ^ 15u; len = 4; }
                else { v = kQuadCodeA[x]; len = kQuadLenA[x]; }
                for (int q = 0; q < 4; q++)
                    if (x & (8u >> q)) { v = (v << 1) | sign_bit(p->sign, kk + q); len++; }
            }
            int tot;
            const int off = warp_scan_excl(len, &tot);
            bitbuf_or(W, pos + off, v, len);
            pos += tot;
        }
    }
    return pos - pos0;
}
#else
HMP3_FN void pack_huffman(const EncTables *T, BitSink *b, const GrSide *g, const short *ix, const unsigned *sign) {
    pack_huffman_seq(T, b, g, ix, sign);
}
// (device only; the host build packs a frame with the sequential writer)
HMP3_FN int pack_gc_bits(const EncTables *, const FrameRec *, const PackGc *, int, unsigned *, int) { return 0; }
#endif

// ------------------------------------------------------------------ side information (l3pack.c:1123-1244)
// Bit writer of the side information: the reference's 32-bit accumulator with lazy byte flush and NO masking of
// the value (l3pack.c:98-153).  The masking matters: part2_3_length is written in 12 bits but can exceed 4095 at
// the highest bitrates, and the reference then ORs the excess bit into the last bit of the previous field when
// that bit is still in the accumulator.  Reproduced as is so that the bytes stay identical.
struct SideBits {
    unsigned char *buf;
    unsigned bitbuf;
    int room;
};
HMP3_HD void side_put(SideBits *w, unsigned x, int nbits) {
    if (w->room < nbits) {
        while (w->room < 24) {
            *w->buf++ = (unsigned char)(w->bitbuf >> (24 - w->room));
            w->room += 8;
        }
    }
    w->bitbuf = (w->bitbuf << nbits) | x;
    w->room -= nbits;
}
HMP3_HD void side_flush(SideBits *w) {
    while (w->room < 24) {
        *w->buf++ = (unsigned char)(w->bitbuf >> (24 - w->room));
        w->room += 8;
    }
    if (w->room < 32) *w->buf++ = (unsigned char)(w->bitbuf << (w->room - 24));
    w->room = 32;
}

// gc = the frame's granule-channel records in (granule, channel) order
HMP3_FN void pack_side(const EncTables *T, const FrameRec *fr, const PackGc *gc, unsigned char *out) {
    SideBits b;
    b.buf = out;
    b.bitbuf = 0;
    b.room = 32;
    const int nch = T->cfg.nchan;
    const bool m1 = T->cfg.h_id == 1;
    if (m1) {
        side_put(&b, (unsigned)fr->main_data_begin, 9);
        side_put(&b, 0, T->cfg.h_mode == 3 ? 5 : 3);
        for (int ch = 0; ch < nch; ch++) side_put(&b, (unsigned)fr->scfsi[ch], 4);
    } else {
        side_put(&b, (unsigned)fr->main_data_begin, 8);
        side_put(&b, 0, T->cfg.h_mode == 3 ? 1 : 2);
    }
    for (int k = 0; k < fr->ngr * nch; k++) {
        const GrSide *g = &gc[k].gr;
        side_put(&b, (unsigned)g->part2_3_length, 12);
        side_put(&b, (unsigned)g->big_values, 9);
        side_put(&b, (unsigned)g->global_gain, 8);
        side_put(&b, (unsigned)g->scalefac_compress, m1 ? 4 : 9);
        side_put(&b, (unsigned)g->window_switching_flag, 1);
        if (g->window_switching_flag) {
            side_put(&b, (unsigned)g->block_type, 2);
            side_put(&b, (unsigned)g->mixed_block_flag, 1);
            side_put(&b, (unsigned)g->table_select[0], 5);
            side_put(&b, (unsigned)g->table_select[1], 5);
            side_put(&b, (unsigned)g->subblock_gain[0], 3);
            side_put(&b, (unsigned)g->subblock_gain[1], 3);
            side_put(&b, (unsigned)g->subblock_gain[2], 3);
        } else {
            side_put(&b, (unsigned)g->table_select[0], 5);
            side_put(&b, (unsigned)g->table_select[1], 5);
            side_put(&b, (unsigned)g->table_select[2], 5);
            side_put(&b, (unsigned)g->region0_count, 4);
            side_put(&b, (unsigned)g->region1_count, 3);
        }
        if (m1) side_put(&b, (unsigned)g->preflag, 1);
        side_put(&b, (unsigned)g->scalefac_scale, 1);
        side_put(&b, (unsigned)g->count1table_select, 1);
    }
    side_flush(&b);
}

// The packing pass for one frame (one warp on the device): scale factors + Huffman data of every granule-
// channel into the stream's main-data stream at the position the serial stage reserved, zero padding up to
// the reservoir minimum, then the side information.  Returns 0, or 1 if the bits written differ from the
// serial stage's accounting (which would be a bug: the counter and the packer use the same tables).
HMP3_FN int pack_frame(const EncTables *T, FrameRec *fr, const PackGc *gc, unsigned char *main) {
    const int nch = T->cfg.nchan;
    const bool m1 = T->cfg.h_id == 1;
    BitSink b;
    sink_open(&b, main + fr->data_start);
    int bad = 0;
    for (int k = 0; k < fr->ngr * nch; k++) {
        const PackGc *p = gc + k;
        const GrSide *g = &p->gr;
        const long long start = b.total_bits;
        if (g->aux_not_null) {
            if (!m1 && T->cfg.is_flag && (k % nch) == 1) write_sf_mpeg2_is(&b, p->sf, g->scalefac_compress);
            else if (!m1) write_sf_mpeg2(&b, p->sf, g->block_type, g->scalefac_compress);
            else if (fr->short_frame) write_sf_mpeg1_plain(&b, p->sf, g->block_type, g->scalefac_compress);
            else write_sf_mpeg1_scfsi(&b, p->sf, (k / nch) ? fr->scfsi[k % nch] : 0, g->scalefac_compress);
            pack_huffman(T, &b, g, p->ix, p->sign);
        }
        if ((int)(b.total_bits - start) != g->part2_3_length) bad = 1;
    }
    const int bytes = (int)sink_close(&b);
#if HMP3_COOP
    for (int k = bytes + PK_LANE; k < fr->data_bytes; k += 32) main[fr->data_start + k] = 0;
#else
    for (int k = bytes; k < fr->data_bytes; k++) main[fr->data_start + k] = 0;
#endif
    pack_side(T, fr, gc, fr->side);
    return bad;
}

// ------------------------------------------------------------------ the rate loop of one granule
// CBitAllo3::BitAllo (bitallo3.cpp:484-678).  xr is this granule's spectrum [2][576], consumed in place.
HMP3_FN void granule_allocate(const EncTables *T, RateState *R, float *xr, const SigMask *sm, PrepGranule *prep, int igr,
                              int nchan, int min_bits, int target_bits, int max_bits, int pool_bits, int ms) {
    LongRate *L = &R->L;
    GrSide *gr = R->gr[igr];
    ScaleFac *sf_out = R->sf[igr];
    QLine *ix = &R->ix[0][0];
    unsigned *sg = &R->signx[0][0];
    const int bt = gr[0].block_type;
    const int init = T->cfg.initial_mnr;
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
        int mnr0;
        if (T->cfg.vbr_flag == 0) {
            mnr0 = L->mnr - (imax_(L->mnr - init, 0) >> 1) - (imax_(L->mnr - init - 400, 0) >> 2);
            mnr0 = imax_(init + 400, mnr0);
        } else mnr0 = init + 400;
        short_granule(T, &R->cold->S, xr, sm, nchan, min_bits, target_bits, max_bits, pool_bits, sf_out, gr, ix, sg, ms, mnr0);
        return;  // (the CBR feedback is a no-op for short blocks, bitallo3.cpp:2905-2909)
    }
    L->ms = ms;
    L->nchan = nchan;
    L->max_bits = imin_(4000 * nchan, max_bits);
    L->min_target = min_bits < 0 ? 0 : min_bits;
    L->target = target_bits;
    L->pool_bits = pool_bits;
    if (T->cfg.vbr_flag == 0) {
        L->pool_fraction = imin_(L->pool_fraction + 50, 614);
        if (bt != 0) L->pool_fraction = 0;
    }
    int tbits = ((L->pool_fraction * L->pool_bits) >> 10);
    if (T->cfg.vbr_flag == 0) tbits = imin_(tbits, imax_((2050 - 500) + init - L->mnr, 200));
    L->max_target = imin_(L->max_bits, L->target + tbits);
    if (L->mnr < -200) L->min_target = imax_(L->min_target, (3 * L->target) >> 2);
    L->max_target = imax_(L->min_target, L->max_target);
    L->min_target = imin_(L->min_target, L->max_target - 100);
    if (ms) long_startup_ms(T, L, sm, prep, sg);
    else long_startup_lr(T, L, sm, prep, sg);
    if (L->active_lines <= 0) {  // digital silence
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
        return;
    }
    const int fb = long_allocate(T, L, xr, ix, ms != 0);
    if (T->cfg.vbr_flag == 0) long_mnr_feedback(T, L, L->active_lines, fb, bt);
    // scale factors on the coded grid (bitallo3.cpp:763-812)
    for (int ch = 0; ch < nchan; ch++) {
        const int sh = L->sf_scale[ch] == 0 ? 1 : 2;
        for (int i = 0; i < T->cfg.nsf[ch]; i++) L->sf[ch][i] >>= sh;
        if (L->preemp[ch])
            for (int i = 11; i < T->cfg.nsf[ch]; i++) L->sf[ch][i] -= sf_pre_amount(i);
        for (int i = 0; i < 21; i++) sf_out[ch].l[i] = L->sf[ch][i];
    }
    if (ms) {  // the mid/side rotation carried no 1/sqrt(2)
        L->G[0] -= 2;
        L->G[1] -= 2;
    }
    for (int ch = 0; ch < nchan; ch++) {
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

// Per-granule inputs produced by Phase A.
struct GranuleIn {
    GranuleInfo info;
    float *xr;            // [2][576], consumed in place
    const SigMask *sm;    // [2][36] psychoacoustic sig/mask of this granule (stage 2 done by the parallel pass)
    PrepGranule *prep;    // long blocks: prepared energies / |x|^(3/4) / step bounds
    int ms;               // M/S decision of the granule's frame (hysteresis applied by the scan pass)
};

HMP3_HD void set_block_info(const EncTables *T, RateState *R, int igr, const GranuleIn *g) {
    // mp3enc.cpp:1429-1438: both channels carry the same decision; the aux fields live in channel 0
    R->gr[igr][0].short_flag_next = g->info.short_next;
    R->gr[igr][0].short_flag_current = g->info.short_cur;
    R->gr[igr][0].block_type_prev = g->info.block_type_prev;
    R->gr[igr][0].block_type = R->gr[igr][1].block_type = g->info.block_type;
}

// Record what the packing pass needs of granule-channel (igr, ch): quantised lines, signs, scale factors, side info.
HMP3_FN void record_gc(const RateState *R, int igr, int ch, PackGc *out) {
    const QLine *ix = R->ix[ch];
    const unsigned *sg = R->signx[ch];
#if HMP3_COOP
    const int lane = HMP3_LANE;
    HMP3_SYNC();
    // only the coded extent is ever read by the packing pass: the big-value pairs and the count1 quads
    const GrSide *g = &R->gr[igr][ch];
    const int extent = g->aux_not_null ? 2 * (g->aux_nreg[0] + g->aux_nreg[1] + g->aux_nreg[2]) + 4 * g->aux_nquads : 0;
    const int nwords = imin_(18, (extent + 31) >> 5);
    {   // the coded lines, two per 32-bit word (both arrays are 4-byte aligned), and their sign words
        const unsigned *src = (const unsigned *)ix;
        unsigned *dst = (unsigned *)out->ix;
        for (int k = lane; k < 16 * nwords; k += HMP3_W) dst[k] = src[k];
        for (int w = lane; w < nwords; w += HMP3_W) out->sign[w] = sg[w];
    }
    const ScaleFac *sf = &R->sf[igr][ch];
    for (int k = lane; k < 23; k += HMP3_W) out->sf[k] = (unsigned char)sf->l[k];
    for (int k = lane; k < 39; k += HMP3_W) out->sf[23 + k] = (unsigned char)sf->s[k / 13][k % 13];
    const int *gs = (const int *)&R->gr[igr][ch];
    int *gd = (int *)&out->gr;
    for (int k = lane; k < (int)(sizeof(GrSide) / sizeof(int)); k += HMP3_W) gd[k] = gs[k];
    HMP3_SYNC();
#else
    for (int k = 0; k < 576; k++) out->ix[k] = ix[k];
    for (int w = 0; w < 18; w++) out->sign[w] = sg[w];
    const ScaleFac *sf = &R->sf[igr][ch];
    for (int i = 0; i < 23; i++) out->sf[i] = (unsigned char)sf->l[i];
    for (int w = 0; w < 3; w++)
        for (int i = 0; i < 13; i++) out->sf[23 + 13 * w + i] = (unsigned char)sf->s[w][i];
    out->gr = R->gr[igr][ch];
#endif
}

// One encode call worth of granules for MPEG-1 (two granules of one frame; mp3enc.cpp:1492-1749).
// Returns the ms flag of the frame; *frame_bits receives the bits of the frame's main data.
HMP3_FN int encode_frame_mpeg1(const EncTables *T, RateState *R, GranuleIn *g0, GranuleIn *g1, PackGc *pack,
                               int *frame_bits) {
    const EncConfig &C = T->cfg;
    const int nch = C.nchan;
    GranuleIn *gs[2] = {g0, g1};
    const int bit_pool = R->byte_pool << 2;
    const int bit_max = R->byte_max << 2, bit_min = R->byte_min << 2;
    int ba_bit_max, ba_bit_min, ba_min, ba_max, dba_max = 0, target;
    const int sf_bits = nch * C.sf_bit_max;
    if (nch == 2) {
        ba_bit_max = bit_max - sf_bits;
        ba_bit_min = bit_min - sf_bits;
        dba_max = bit_pool >> 2;
        ba_min = ba_bit_min;
        ba_max = ba_bit_max + dba_max;
        target = C.ave_target_bits + C.ave_target_bits;
    } else {
        ba_bit_max = bit_max;
        if (ba_bit_max > 4095) ba_bit_max = 4095;
        ba_bit_min = bit_min;
        ba_bit_max -= C.sf_bit_max;
        ba_bit_min -= C.sf_bit_max;
        ba_min = ba_bit_min;
        ba_max = ba_bit_max;
        target = C.ave_target_bits;
    }
    set_block_info(T, R, 0, g0);
    set_block_info(T, R, 1, g1);
    const int short_frame = (g0->info.block_type == 2) | (g1->info.block_type == 2);
    const int ms = g0->ms;  // frame decision (both granules carry it)
    int total = 0;
    for (int igr = 0; igr < 2; igr++) {
        if (igr) HMP3_GRANULE_BARRIER();
        granule_allocate(T, R, gs[igr]->xr, gs[igr]->sm, gs[igr]->prep, igr, nch, ba_min, target, ba_max, bit_pool, ms);
        for (int ch = 0; ch < nch; ch++) {
            GrSide *g = &R->gr[igr][ch];
            int sfb = 0;
            g->scalefac_compress = 0;
            if (short_frame) {
                R->scfsi[ch] = 0;
                if (g->aux_not_null) g->scalefac_compress = plan_sf_mpeg1_plain(&R->sf[igr][ch], g->block_type, &sfb);
            } else {
                g->scalefac_compress =
                    plan_sf_mpeg1_scfsi(&R->sf[igr][ch], R->sf_save[ch], igr, &R->scfsi[ch], g->aux_not_null, &sfb);
            }
            // the Huffman bits are exactly the counted ones (the counter and the packer share their tables)
            const int bits = g->aux_not_null ? sfb + g->aux_bits : 0;
            if (nch == 2) {
                ba_min -= bits;
                ba_max -= bits;
            } else {
                ba_min += ba_bit_min + C.sf_bit_max - bits;
                ba_max += ba_bit_max + C.sf_bit_max - bits;
            }
            g->part2_3_length = bits;
            total += bits;
            record_gc(R, igr, ch, pack + (igr * nch + ch));
        }
        if (nch == 2) {
            ba_min += ba_bit_min + sf_bits;
            ba_max = ba_max - dba_max;
            ba_max += ba_bit_max + sf_bits;
        }
    }
    *frame_bits = total;
    return ms;
}

// One MPEG-2 frame = one granule (mp3enc.cpp:1832-2027)
HMP3_FN int encode_frame_mpeg2(const EncTables *T, RateState *R, int igr, GranuleIn *g0, PackGc *pack, int *frame_bits) {
    const EncConfig &C = T->cfg;
    const int nch = C.nchan;
    const int bit_pool = R->byte_pool << 3;
    int bit_max = (R->byte_max << 3), bit_min = (R->byte_min << 3);
    if (nch == 2 && R->byte_pool > 245) bit_min += 40;
    int ba_bit_max = bit_max;
    if (ba_bit_max > 4095) ba_bit_max = 4095;
    int ba_bit_min = bit_min;
    ba_bit_max -= nch * C.sf_bit_max;
    ba_bit_min -= nch * C.sf_bit_max;
    set_block_info(T, R, igr, g0);
    const int ms = g0->ms;
    granule_allocate(T, R, g0->xr, g0->sm, g0->prep, igr, nch, ba_bit_min, nch * C.ave_target_bits, ba_bit_max, bit_pool,
                     nch == 2 ? ms : C.ms_flag);
    int total = 0;
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
        total += bits;
        record_gc(R, igr, ch, pack + ch);
    }
    *frame_bits = total;
    return ms;
}


#if HMP3_WITH_ALLO1
// ------------------------------------------------------------------ the CBitAllo1 drivers
// encode_jointA / encode_singleA and their MPEG-2 forms (mp3enc.cpp:1236-1321, 1601-1671, 1753-1828, 1911-1973):
// long blocks only, M/S decided per frame (MPEG-1) or granule (MPEG-2) from the allocator's own correlation measure
// without hysteresis, dual channel allocated one channel at a time.
HMP3_FN void allo1_io(RateState *R, GranuleIn *g, int ch0, int nch, Allo1Io *io) {
    float(*x34)[576] = (float(*)[576]) & R->cold->S.x34[0][0][0];  // the short-block work area is free: no short blocks here
    for (int c = 0; c < nch; c++) {
        const int ch = ch0 + c;
        io->xr[c] = g->xr + 576 * ch;
        io->x34[c] = x34[ch];
        io->ix[c] = &R->ix[ch][0];
        io->sign[c] = &R->signx[ch][0];
        io->sm[c] = g->sm + 36 * ch;
    }
}
HMP3_FN int encode_frame_a1(const EncTables *T, RateState *R, int igr_arg, GranuleIn *g0, GranuleIn *g1, PackGc *pack,
                            int *frame_bits) {
    const EncConfig &C = T->cfg;
    const int nch = C.nchan;
    const bool m1 = C.h_id == 1, joint = C.h_mode == 1;
    Allo1 *A = &R->cold->A1;
    GranuleIn *gs[2] = {g0, g1};
    const int ngr = m1 ? 2 : 1;
    int bit_max, bit_min;
    if (m1) {
        const int sh = (joint || nch != 2) ? 2 : 1;
        bit_max = R->byte_max << sh;
        bit_min = R->byte_min << sh;
    } else {
        const int sh = (joint || nch != 2) ? 3 : 2;
        bit_max = R->byte_max << sh;
        bit_min = R->byte_min << sh;
        if (joint && R->byte_pool > 245) bit_min += 40;
    }
    int ba_bit_max = bit_max > 4095 ? 4095 : bit_max, ba_bit_min = bit_min;
    const int sf_bits = joint ? 2 * C.sf_bit_max : C.sf_bit_max;
    ba_bit_max -= sf_bits;
    ba_bit_min -= sf_bits;
    int ba_min = ba_bit_min, ba_max = ba_bit_max;
    int ms = 0;
    if (joint && C.ms_flag) {
        int m = allo1_ms_measure(T, g0->xr, g0->xr + 576);
        if (m1) m += allo1_ms_measure(T, g1->xr, g1->xr + 576);
        if (m >= 0) ms = 1;
    }
    int total = 0;
    for (int q = 0; q < ngr; q++) {
        const int igr = m1 ? q : igr_arg;
        GranuleIn *g = gs[q];
        set_block_info(T, R, igr, g);
        Allo1Io io;
        if (joint) {
            allo1_io(R, g, 0, 2, &io);
            allo1_granule(T, A, &io, 0, 2, ba_min, 2 * C.ave_target_bits, ba_max, &R->sf[igr][0], &R->gr[igr][0], ms);
        }
        for (int ch = 0; ch < nch; ch++) {
            GrSide *gr = &R->gr[igr][ch];
            if (!joint) {
                allo1_io(R, g, ch, 1, &io);
                allo1_granule(T, A, &io, ch, 1, ba_min, C.ave_target_bits, ba_max, &R->sf[igr][ch], gr, C.ms_flag);
            }
            int bits = 0, sfb = 0;
            gr->scalefac_compress = 0;
            if (m1 && joint) {
                gr->scalefac_compress =
                    plan_sf_mpeg1_scfsi(&R->sf[igr][ch], R->sf_save[ch], igr, &R->scfsi[ch], gr->aux_not_null, &sfb);
                if (gr->aux_not_null) bits = sfb + gr->aux_bits;
            } else if (m1) {
                if (gr->aux_bits) {
                    gr->scalefac_compress = plan_sf_mpeg1_plain(&R->sf[igr][ch], gr->block_type, &sfb);
                    bits = sfb + gr->aux_bits;
                }
            } else if (joint ? gr->aux_not_null : gr->aux_bits) {
                if (joint && (ch & C.is_flag)) gr->scalefac_compress = plan_sf_mpeg2_is(&R->sf[igr][ch], C.nsf_stereo, &sfb);
                else gr->scalefac_compress = plan_sf_mpeg2(&R->sf[igr][ch], R->gr[igr][0].block_type, &sfb);
                bits = sfb + gr->aux_bits;
            }
            if (joint) {
                if (m1) {
                    ba_min -= bits;
                    ba_max -= bits;
                }
            } else {
                ba_min += ba_bit_min + C.sf_bit_max - bits;
                ba_max += ba_bit_max + C.sf_bit_max - bits;
            }
            gr->part2_3_length = bits;
            total += bits;
            record_gc(R, igr, ch, pack + (q * nch + ch));
        }
        if (joint && m1) {
            ba_min += ba_bit_min + sf_bits;
            ba_max += ba_bit_max + sf_bits;
        }
    }
    *frame_bits = total;
    return ms;
}
#endif

// ------------------------------------------------------------------ frame driver + reservoir
// Records one frame: header, its main-data slot and where its own data goes in the stream's main-data
// stream (mp3enc.cpp:2106-2593).
HMP3_HD void frame_header(const EncTables *T, unsigned char *h, int pad, int mode_ext, int br_index) {
    h[0] = T->cfg.head[0];
    h[1] = T->cfg.head[1];
    h[2] = T->cfg.head[2];
    h[3] = T->cfg.head[3];
    if (T->cfg.vbr_flag) h[2] = (unsigned char)((h[2] & 0x0F) | (br_index << 4));
    else if (pad) h[2] |= 2;
    h[3] = (unsigned char)((h[3] & 0xCF) | (mode_ext << 4));
}

// Encode the granules of one frame (MPEG-1: g0,g1; MPEG-2: g0 with granule parity igr).  `pack` receives the
// frame's granule-channel records, K is the frame's first encode granule.
HMP3_FN void encode_one_frame(const EncTables *T, RateState *R, FrameRec *frames, int igr, GranuleIn *g0,
                              GranuleIn *g1, PackGc *pack, int K) {
    const EncConfig &C = T->cfg;
    const bool m1 = C.h_id == 1;
    int pad = 0;
    int mf_bytes = 0;
    if (!C.vbr_flag) {
        R->padcount -= C.pad_remainder;
        if (R->padcount <= 0) {
            R->padcount += C.pad_divisor;
            pad = 1;
        }
        mf_bytes = C.main_framebytes + pad;
    }
    FrameRec *fr = frames + R->frames;
    fr->main_start = R->mf_tot;
    R->byte_pool = (int)(R->mf_tot - R->main_tot);
    if (C.vbr_flag) {
        R->byte_max = C.vbr_main_framebytes[C.ivbr_max] + R->byte_pool;
        R->byte_min = C.vbr_main_framebytes[C.ivbr_min] + R->byte_pool - C.reservoir_back;
    } else {
        R->byte_max = C.main_framebytes + pad + R->byte_pool;
        R->byte_min = R->byte_max - C.reservoir_back;
    }
    const int main_data_begin = R->byte_pool;
    int frame_bits = 0;
#if HMP3_WITH_ALLO1
    const int ms = C.allocator == 1 ? encode_frame_a1(T, R, igr, g0, g1, pack, &frame_bits)
                   : (m1 ? encode_frame_mpeg1(T, R, g0, g1, pack, &frame_bits) : encode_frame_mpeg2(T, R, igr, g0, pack, &frame_bits));
#else
    const int ms = m1 ? encode_frame_mpeg1(T, R, g0, g1, pack, &frame_bits) : encode_frame_mpeg2(T, R, igr, g0, pack, &frame_bits);
#endif
    const int mode_ext = ms + ms + C.is_flag;
    int bytes = (frame_bits + 7) >> 3;
    int ibr = 0;
    if (C.vbr_flag) {
        const int bytes2 = bytes - R->byte_pool;
        const int bytes3 = bytes2 + C.vbr_pool_target;
        for (ibr = C.ivbr_min; ibr <= C.ivbr_max; ibr++)
            if (bytes2 <= C.vbr_main_framebytes[ibr]) break;
        bool grow = true;
        if (!m1) {  // MPEG-2: keep the number of frames spanned by the reservoir bounded (mp3enc.cpp:2390-2413)
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
    if (bytes < R->byte_min) bytes = R->byte_min;  // the packing pass zero-fills up to here
    fr->mf_bytes = mf_bytes;
    fr->out_off = R->out_tot;
    R->out_tot += (unsigned)(4 + C.side_bytes + mf_bytes);
    fr->data_start = R->main_tot;
    fr->data_bits = frame_bits;
    fr->data_bytes = bytes;
    fr->granule0 = K;
    fr->ngr = (short)(m1 ? 2 : 1);
    fr->igr0 = (short)igr;
    fr->main_data_begin = (short)main_data_begin;
    fr->short_frame = (short)(m1 ? ((g0->info.block_type == 2) | (g1->info.block_type == 2)) : 0);
    if (C.allocator == 1 && m1 && C.h_mode != 1) fr->short_frame = 1;  // encode_singleA codes the scale factors without scfsi
    fr->scfsi[0] = (short)R->scfsi[0];
    fr->scfsi[1] = (short)R->scfsi[1];
    frame_header(T, fr->head, pad, mode_ext, ibr);
    R->main_tot += bytes;
    R->mf_tot += mf_bytes;
    R->frames++;
    // frames whose slot is now completely covered by produced main data
    while (R->frames_done < R->frames) {
        const FrameRec *f = frames + R->frames_done;
        if ((long long)R->main_tot - (long long)f->main_start < f->mf_bytes) break;
        R->frames_done++;
    }
    fr->done_after = R->frames_done;
}

}  // namespace hmp3

namespace hmp3 {

// Run the serial stage of one stream over the encode granules [K0, K0+NG) that Phase A has prepared.
// gi/xr/sm/prep/ms/pack point at this stream's slice of the chunk buffers (index 0 == granule K0; pack holds
// two records per granule).  ngran_real = granules of real encode calls (2 per call); after them the stream
// keeps consuming zero-PCM granules until every real frame's main-data slot is filled (the CLI's tail flush,
// test/tomp3.cpp:1015-1036), checked at call boundaries.
HMP3_FN void rate_run_chunk(const EncTables *T, RateState *R, int K0, int NG, int ngran, int ngran_real,
                            const GranuleInfo *gi, float *xr, const SigMask *sm, PrepGranule *prep,
                            const signed char *ms, PackGc *pack, FrameRec *frames, bool live = true) {
    const bool m1 = live && T->cfg.h_id == 1;
    const int frames_real = m1 ? ngran_real / 2 : ngran_real;
    for (int K = K0; K + 1 < K0 + NG; K += 2) {
        HMP3_FRAME_BARRIER();
        if (!live || K + 1 >= ngran || R->finished || K < R->next_granule) {
            HMP3_GRANULE_BARRIER();
            continue;
        }
        GranuleIn g[2];
        for (int q = 0; q < 2; q++) {
            const int o = K - K0 + q;
            g[q].info = gi[o];
            g[q].xr = xr + (long long)o * 2 * 576;
            g[q].sm = sm + (long long)o * 72;
            g[q].prep = prep + o;
            g[q].ms = ms[o];
        }
        PackGc *pk = pack + (long long)(K - K0) * 2;
        if (m1) encode_one_frame(T, R, frames, 0, &g[0], &g[1], pk, K);
        else {
            encode_one_frame(T, R, frames, 0, &g[0], nullptr, pk, K);
            HMP3_GRANULE_BARRIER();
            encode_one_frame(T, R, frames, 1, &g[1], nullptr, pk + 2, K + 1);
        }
        R->next_granule = K + 2;
        if (K + 2 >= ngran_real && R->frames_done >= frames_real) R->finished = 1;
    }
}

// Final byte layout of frame f of a stream: header | side info | slice of the main-data stream.
HMP3_HD int frame_bytes(const EncTables *T, const FrameRec *f) { return 4 + T->cfg.side_bytes + f->mf_bytes; }

}  // namespace hmp3
