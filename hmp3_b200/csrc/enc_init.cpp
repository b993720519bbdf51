// Host-side resolution of a control block into the constants and tables the kernels read.
// Pure integer/float host logic (double-precision libm at init, rounded to float tables) following the
// reference's init rules; every block cites the rule it reproduces.  Compared table-by-table with the
// oracle in tests/test_cpu_parity.py (resolved values, golden tables) and tests/test_cabi.py.
#include "enc_init.h"

#include <math.h>
#include <algorithm>
#include <stdlib.h>
#include <string.h>

#define HMP3_TABLE_QUAL const
#include "tables_data.h"

namespace hmp3 {

namespace {

// ISO 11172-3 / 13818-3 scale-factor band edges [h_id][sr_index] (reference copy: l3init.c:54-98)
struct BandEdges { int l[23]; int s[14]; };
const BandEdges kBandEdges[2][3] = {
    {{{0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 116, 140, 168, 200, 238, 284, 336, 396, 464, 522, 576},
      {0, 4, 8, 12, 18, 24, 32, 42, 56, 74, 100, 132, 174, 192}},
     {{0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 114, 136, 162, 194, 232, 278, 332, 394, 464, 540, 576},
      {0, 4, 8, 12, 18, 26, 36, 48, 62, 80, 104, 136, 180, 192}},
     {{0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 116, 140, 168, 200, 238, 284, 336, 396, 464, 522, 576},
      {0, 4, 8, 12, 18, 26, 36, 48, 62, 80, 104, 134, 174, 192}}},
    {{{0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 52, 62, 74, 90, 110, 134, 162, 196, 238, 288, 342, 418, 576},
      {0, 4, 8, 12, 16, 22, 30, 40, 52, 66, 84, 106, 136, 192}},
     {{0, 4, 8, 12, 16, 20, 24, 30, 36, 42, 50, 60, 72, 88, 106, 128, 156, 190, 230, 276, 330, 384, 576},
      {0, 4, 8, 12, 16, 22, 28, 38, 50, 64, 80, 100, 126, 192}},
     {{0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 54, 66, 82, 102, 126, 156, 194, 240, 296, 364, 448, 550, 576},
      {0, 4, 8, 12, 16, 22, 30, 42, 58, 78, 104, 138, 180, 192}}}};

const int kRates[8] = {22050, 24000, 16000, 1, 44100, 48000, 32000, 1};  // setup.c:48-49
const int kBitrates[2][16] = {{0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160, -1},
                              {0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, -1}};

inline int imin(int a, int b) { return a < b ? a : b; }
inline int imax(int a, int b) { return a > b ? a : b; }

float bits_f(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// number of long / short bands that end at or below `limit` lines (l3init.c:397-414, 434-451)
int sfb_limit_long(const BandEdges &e, int limit) {
    int i;
    for (i = 0; i < 23; i++)
        if (limit <= e.l[i]) break;
    return i > 21 ? 21 : i;
}
int sfb_limit_short(const BandEdges &e, int limit) {
    int i;
    for (i = 0; i < 14; i++)
        if (limit <= e.s[i]) break;
    return i > 12 ? 12 : i;
}

// nearest long-band edge frequency (l3init.c:140-171)
int nearest_band_freq(int sr_index, int h_id, int freq) {
    int samprate = kRates[4 * h_id + sr_index];
    float a = samprate / (2.0f * 576.0f);
    int best = 999999, fout = freq;
    for (int i = 0; i < 21; i++) {
        int f = (int)(a * kBandEdges[h_id][sr_index].l[i + 1] + 0.5f);
        int d = abs(f - freq);
        if (d < best) { best = d; fout = f; }
    }
    return fout;
}

// CBR bandwidth rule (mp3enc.cpp:899-935)
int cbr_freq_limit(int samprate, int totbitrate, int mode) {
    static const float factor[4] = {1.1f, 1.333f, 1.0f, 1.0f};
    if (samprate < 8000) samprate = 8000;
    float chan_bitrate = (float)totbitrate;
    if (mode != 3) chan_bitrate = (float)(0.5 * chan_bitrate);
    chan_bitrate = factor[mode] * chan_bitrate;
    int flimit;
    if (samprate < 32000) {
        if (chan_bitrate <= 32.0f) flimit = (int)(752.0 + 203.0 * chan_bitrate);
        else if (chan_bitrate <= 42.7f) flimit = (int)(-2967.0 + 327.0 * chan_bitrate);
        else flimit = 11000;
    } else {
        flimit = (int)(187.97 * chan_bitrate);
    }
    return flimit;
}

// ---------------------------------------------------------------- psychoacoustic tables (amodini2.c)
float hz_to_bark(float f) {  // amodini2.c:354-366
    float t = (1.0f / 1000.0f) * f;
    float tt = (1.0f / 7.5f) * t;
    tt = tt * tt;
    return (float)(13.0 * atan((double)(0.76f * t)) + 3.5 * atan((double)tt));
}
float interp_xy(const float xy[][2], float x) {  // amodini2.c:369-385
    int i;
    for (i = 1; i < 100; i++)
        if (x <= xy[i][0]) break;
    return xy[i - 1][1] + (x - xy[i - 1][0]) * ((xy[i][1] - xy[i - 1][1])) / (xy[i][0] - xy[i - 1][0]);
}
// Schroeder spreading (Painter & Spanias form), short blocks (amodini2.c:183-203)
float spread_plain(float bz0, float bz) {
    const double a = 0.2302585093;
    double tx = (bz0 - bz) * 1.00;
    tx += 0.474;
    double ty = 15.811389 + 7.5 * tx - 17.5 * sqrt(1.0 + tx * tx);
    float s;
    if (ty <= -60.0) s = 0.0;
    else s = exp(ty * a);
    return s;
}
// long-block variant with frequency dependent slopes (amodini2.c:207-251)
float spread_sloped(float bz0, float bz) {
    const double a = 0.2302585093;
    double t1 = 1.2, t2 = 1.2;
    double dt = (0.5 / 7.0) * (7.0 - bz0);
    if (dt < 0.0) dt = 0.0;
    t1 = t1 + dt;
    t2 = t2 + dt;
    dt = bz0 - 22.5;
    if (dt < 0.0) dt = 0.0;
    t2 = t2 + dt;
    double tx = (bz0 - bz);
    if (tx > 0.0) tx = t1 * tx;
    else tx = t2 * tx;
    tx += 0.474;
    double ty = 15.811389 + 7.5 * tx - 17.5 * sqrt(1.0 + tx * tx);
    float s;
    if (ty <= -60.0) s = 0.0;
    else s = exp(ty * a);
    return s;
}

typedef float (*SpreadFn)(float, float);

// Spreading rows: for masker partition i, the run of maskee partitions whose weight exceeds 1e-6
// (amodini2.c:388-402, 434-458, 461-579).  scale_row = fixed 0.35 factor for short blocks.
int build_spreading(SpreadFn fn, const float *bval, float *snr_factor, int npart, bool short_rule, int *cnt,
                    int *off, int *w0, float *w, int *npart_out) {
    float s[64];
    float thres = 1.0e-6;
    for (int i = 0; i < 64; i++) cnt[i] = off[i] = w0[i] = 0;
    int ntot = 0, i;
    for (i = 0; i < npart; i++) {
        for (int j = 0; j < 64; j++) s[j] = 0.0;
        for (int j = 0; j < npart; j++) s[j] = fn(bval[i], bval[j]);
        // keep the first run above threshold, zero everything else (amodini2.c:434-458)
        int j = 0;
        for (; j < npart; j++) {
            if (s[j] > thres) break;
            s[j] = 0.0;
        }
        for (; j < npart; j++)
            if (s[j] <= thres) break;
        for (; j < npart; j++) s[j] = 0.0;
        for (j = 0; j < npart; j++)
            if (s[j] != 0.0f) break;
        int nj = j;
        if (nj >= npart) break;
        int count = 0;
        w0[i] = ntot;
        for (; j < npart; j++) {
            if (s[j] == 0.0) break;
            count++;
            if (short_rule) {
                float r_norm = 0.35f;
                w[ntot] = r_norm * snr_factor[i] * s[j];
            } else {
                w[ntot] = snr_factor[i] * s[j];
            }
            ntot++;
        }
        cnt[i] = count;
        off[i] = nj;
        if (short_rule) snr_factor[i] = 0.35f * snr_factor[i];
    }
    *npart_out = i;
    return ntot;
}

void build_psy_short(EncTables &T, int sr_index, int nsb, int h_id) {  // amodini2.c:587-739
    static const float db_snr[][2] = {{0.0f, 12.0f},   {861.0f, 10.0f},  {2584.0f, 8.0f},  {5857.0f, 7.0f},
                                      {9302.0f, 5.0f}, {13092.0f, 4.0f}, {15500.0f, 3.0f}, {99999.0f, -2.0f}};
    const int maxpart = 32;
    int part[32];
    float snr_factor[32], bval[32];
    int f_select = sr_index & 3;
    if (f_select == 3) f_select = 0;
    for (int i = 0; i < maxpart; i++) part[i] = 192;
    int nb[14];
    for (int i = 0; i < 14; i++) nb[i] = 0;
    for (int i = 0; i < 13; i++) nb[i] = T.nBand_s[i];
    int t = 0;
    for (int i = 0; i < 14; i++) {  // model partition = half a coder band
        part[2 * i] = t;
        int m = nb[i] / 2;
        t += m;
        part[2 * i + 1] = t;
        m = nb[i] - m;
        t += m;
    }
    int nbin = 6 * nsb, i;
    for (i = 0; i < maxpart; i++)
        if (part[i] >= nbin) break;
    int npart = imin(i, 2 * 12);
    float x = 0.5f * kRates[4 * h_id + f_select] / 192;
    for (i = 0; i < maxpart - 1; i++) {
        float freq = x * 0.5f * (part[i] + part[i + 1]);
        if (h_id == 1) snr_factor[i] = 0.7 * pow(10.0, -0.1 * interp_xy(db_snr, freq));
        else snr_factor[i] = 2.8 * pow(10.0, -0.1 * interp_xy(db_snr, freq));
        bval[i] = hz_to_bark(freq);
    }
    snr_factor[i] = 1.0f;
    bval[i] = bval[i - 1];
    memset(T.w_spd_s, 0, sizeof(T.w_spd_s));
    int np2;
    build_spreading(spread_plain, bval, snr_factor, npart, true, T.spd_cnt_s, T.spd_off_s, T.spd_w0_s, T.w_spd_s,
                    &np2);
    T.psy_npart_s = np2;
    for (i = 0; i < 64; i++) T.psy_nsum_s[i] = 0;
    for (i = 0; i < npart; i++) T.psy_nsum_s[i] = part[i + 1] - part[i];
    // the energy map runs over npart (nsum[66]); the spreading loop over cntl[64].count (== np2)
    T.psy_start_s[0] = 0;
    for (i = 0; i < 64; i++) T.psy_start_s[i + 1] = T.psy_start_s[i] + T.psy_nsum_s[i];
    T.cfg.nsf_s[0] = T.cfg.nsf_s[0];  // (set elsewhere)
    // note: emap count
    T.psy_nsum_s[63] = 0;
    T.psy_emap_n_s = npart;
}

void build_psy_long(EncTables &T, int sr_index, int nsb, int h_id) {  // amodini2.c:743-940
    static const float db_snr[][2] = {
        {0, 0.0f},     {38, 0.0f},    {115, 0.0f},   {191, 0.0f},   {268, 0.0f},   {345, 0.0f},   {421, 0.0f},
        {498, 0.0f},   {574, 0.0f},   {651, 1.0f},   {727, 1.0f},   {804, 2.5f},   {880, 2.5f},   {976, 1.5f},
        {1091, 1.5f},  {1206, 2.0f},  {1321, 2.0f},  {1455, 2.0f},  {1608, 3.0f},  {1761, 3.0f},  {1914, 3.0f},
        {2086, 3.0f},  {2278, 3.0f},  {2488, 1.0f},  {2718, 1.0f},  {2986, 0.0f},  {3292, 0.0f},  {3637, 0.0f},
        {4020, 0.0f},  {4441, 0.0f},  {4900, 0.0f},  {5398, 0.0f},  {5934, 0.0f},  {6527, 0.0f},  {7178, 0.0f},
        {7905, 0.0f},  {8709, 0.0f},  {9589, 0.0f},  {10546, 0.0f}, {11542, 0.0f}, {12575, 0.0f}, {13820, -2.0f},
        {15274, -2.0f}, {99999, 0.0f}};
    static const float abs_thres[][2] = {{0.0f, 5.0f},    {350.0f, 0.03f},  {2584.0f, 0.01f}, {5857.0f, 0.01f},
                                         {9302.0f, 0.03f}, {13092.0f, 0.5f}, {15500.0f, 5.0f}, {99999.0f, 100.0f}};
    const int maxpart = 64;
    int part[64];
    float snr_factor[64], bval[64], athres[64];
    memset(athres, 0, sizeof(athres));
    int f_select = sr_index & 3;
    if (f_select == 3) f_select = 0;
    for (int i = 0; i < maxpart; i++) part[i] = 576;
    int t = 0;
    for (int i = 0; i < 22; i++) {
        part[2 * i] = t;
        int m = T.nBand_l_iso[i] / 2;
        t += m;
        part[2 * i + 1] = t;
        m = T.nBand_l_iso[i] - m;
        t += m;
    }
    int nbin = 18 * nsb, i;
    for (i = 0; i < maxpart; i++)
        if (part[i] >= nbin) break;
    int npart = imin(i, 2 * 21);
    float x = 0.5f * kRates[4 * h_id + f_select] / 576;
    for (i = 0; i < maxpart - 1; i++) {
        float freq = x * 0.5f * (part[i] + part[i + 1]);
        snr_factor[i] = pow(10.0, -0.1 * interp_xy(db_snr, freq));
        bval[i] = hz_to_bark(freq);
        athres[i] = interp_xy(abs_thres, freq) * (part[i + 1] - part[i]);
    }
    snr_factor[i] = 1.0f;
    bval[i] = bval[i - 1];
    memset(T.w_spd_l, 0, sizeof(T.w_spd_l));
    int np2;
    int ntot = build_spreading(spread_sloped, bval, snr_factor, npart, false, T.spd_cnt_l, T.spd_off_l, T.spd_w0_l,
                               T.w_spd_l + 128, &np2);
    T.psy_npart_l = np2;
    for (i = 0; i < 64; i++) T.spd_w0_l[i] += 128;
    // weights enter a non-linear (x^0.3) summation (amodini2.c:897-903)
    for (i = 128; i < ntot + 128; i++)
        if (T.w_spd_l[i] > 0.0f) T.w_spd_l[i] = pow(T.w_spd_l[i], 0.30);
    for (i = 0; i < 64; i++) T.w_spd_l[i] = athres[i];
    for (i = 0; i < 64; i++) T.psy_nsum_l[i] = 0;
    for (i = 0; i < npart; i++) T.psy_nsum_l[i] = part[i + 1] - part[i];
    T.psy_start_l[0] = 0;
    for (i = 0; i < 64; i++) T.psy_start_l[i + 1] = T.psy_start_l[i] + T.psy_nsum_l[i];
    T.psy_emap_n_l = npart;
}

// ---------------------------------------------------------------- Huffman bit-count LUTs
// The counter sums, per (x,y) pair, the code length + sign bits + linbits of two candidate tables at
// once in the low/high halves of one word; which candidates depends on the largest value in the
// region (reference: SelectTable/CountCase, cnttab.h:38-62, 797-899; bitalloc.cpp:310-404).
// We derive those words from the ISO code books instead of transcribing them.
struct CountClassDef { int maxval; int tab[4]; int tmax; };
const CountClassDef kClasses[kCountClasses] = {
    {0, {0, 0, 0, 0}, 0},          {1, {1, 3, 0, 0}, 1},        {2, {2, 3, 0, 0}, 2},
    {3, {5, 6, 0, 0}, 3},          {5, {7, 8, 9, 12}, 5},       {7, {10, 11, 12, 15}, 7},
    {15, {13, 15, 0, 0}, 15},      {16, {16, 24, 0, 0}, 16},    {18, {17, 24, 0, 0}, 18},
    {22, {18, 24, 0, 0}, 22},      {30, {19, 24, 0, 0}, 30},    {46, {25, 20, 0, 0}, 46},
    {78, {20, 26, 0, 0}, 78},      {142, {27, 21, 0, 0}, 142},  {270, {21, 28, 0, 0}, 270},
    {526, {29, 22, 0, 0}, 526},    {1038, {22, 30, 0, 0}, 1038}, {2062, {30, 23, 0, 0}, 2062},
    {0x7fffffff, {31, 23, 0, 0}, 8206}, {0, {0, 0, 0, 0}, 0}};

int pair_bits(const EncTables &T, int table, int x, int y) {
    if (table == 0) return 0;
    int book = T.huff_sel_book[table];
    int dim = kHuffBookDim[book];
    if (x >= dim || y >= dim) return 0;
    int len = (int)(T.huff_book[book][x * 16 + y] >> 24);
    if (len == 0) return 0;  // not a code word of this book
    int lb = T.huff_linbits[table];
    return len + (x != 0) + (y != 0) + (x == 15 ? lb : 0) + (y == 15 ? lb : 0);
}

void build_count_luts(EncTables &T) {
    for (int c = 0; c < kCountClasses; c++) {
        const CountClassDef &d = kClasses[c];
        for (int k = 0; k < 4; k++) T.cnt_tables[c][k] = d.tab[k];
        T.cnt_tmax[c] = d.tmax;
        T.cnt_ncand[c] = d.tab[0] == 0 ? 0 : (d.tab[2] ? 4 : 2);
        for (int x = 0; x < 16; x++)
            for (int y = 0; y < 16; y++) {
                uint32_t w0 = (uint32_t)pair_bits(T, d.tab[0], x, y) | ((uint32_t)pair_bits(T, d.tab[1], x, y) << 16);
                uint32_t w1 = (uint32_t)pair_bits(T, d.tab[2], x, y) | ((uint32_t)pair_bits(T, d.tab[3], x, y) << 16);
                T.cnt_lut[c][x * 16 + y][0] = w0;
                T.cnt_lut[c][x * 16 + y][1] = w1;
            }
    }
    for (int m = 0; m < 24; m++) {
        int c = 0;
        while (kClasses[c].maxval < m) c++;
        T.cnt_class_of_max[m] = c;
    }
}

}  // namespace

void control_defaults(hmp3_control *ec) {  // test/tomp3.cpp:357-387
    memset(ec, 0, sizeof(*ec));
    ec->mode = 1;
    ec->bitrate = -1;
    ec->samprate = 44100;
    ec->nsbstereo = -1;
    ec->filter_select = -1;
    ec->nsb_limit = -1;
    ec->freq_limit = 24000;
    ec->cr_bit = 1;
    ec->original = 1;
    ec->layer = 3;
    ec->hf_flag = 0;
    ec->vbr_flag = 1;
    ec->vbr_mnr = 50;
    ec->vbr_br_limit = 160;
    ec->chan_add_f0 = 24000;
    ec->chan_add_f1 = 24000;
    ec->sparse_scale = -1;
    ec->vbr_delta_mnr = 0;
    ec->cpu_select = 0;
    ec->quick = -1;
    ec->test1 = -1;
    ec->test2 = 0;
    ec->test3 = 0;
    ec->short_block_threshold = 700;
}

// L3_audio_encode_init (mp3enc.cpp:220-870) + BitAlloInit (bitallo3.cpp:288-480, bitallos.cpp:128-198).
// Returns bytes_in (nchan*4*1152) or 0 on rejection; *unsupported is set when the control selects
// the intensity-stereo / dual-channel allocator that is outside the built path.
// One command-line option with the reference CLI's meaning (test/tomp3.cpp:390-558).  Options that do not
// touch E_CONTROL (-D -EC -X -IL -A -P -Z, file names) are accepted and ignored.  Returns 0, or -1 if the
// option is not one the reference knows (the reference prints its usage for those).
int control_apply_option(hmp3_control *ec, const char *opt) {
    if (!opt || opt[0] != '-') return -1;
    auto lc = [](char c) { return (c >= 'A' && c <= 'Z') ? (char)(c + 32) : c; };
    const char c1 = lc(opt[1]), c2 = opt[1] ? lc(opt[2]) : 0, c3 = (opt[1] && opt[2]) ? lc(opt[3]) : 0;
    switch (c1) {
        case 'e': return 0;                                    // -EC: display only
        case 'h':
            if (c2 == 'f') { ec->hf_flag = 1 | atoi(opt + 3); return 0; }
            return -1;                                         // -h: usage
        case 'q': ec->quick = atoi(opt + 2); return 0;
        case 'd': return 0;
        case 'u': ec->cpu_select = atoi(opt + 2); return 0;
        case 'x': return 0;                                    // Xing/Info header: host post-pass, not this path
        case 'b': ec->bitrate = atoi(opt + 2); break;
        case 'c': ec->cr_bit = atoi(opt + 2); return 0;
        case 'o': ec->original = atoi(opt + 2); return 0;
        case 'p': return 0;
        case 'm': ec->mode = atoi(opt + 2); return 0;
        case 'n': ec->nsbstereo = atoi(opt + 2); return 0;
        case 's':
            if (c2 == 'b') { if (c3 == 't') ec->short_block_threshold = atoi(opt + 4); }
            else ec->filter_select = atoi(opt + 2);
            return 0;
        case 'f': ec->freq_limit = atoi(opt + 2); return 0;
        case 'z': return 0;
        case 't':
            if (c2 == 'x') ec->test1 = atoi(opt + 3);
            else ec->vbr_delta_mnr = atoi(opt + 2);
            return 0;
        case 'i':
            if (c2 != 'l') ec->chan_add_f0 = atoi(opt + 2);
            return 0;
        case 'j': ec->chan_add_f1 = atoi(opt + 2); return 0;
        case 'v': ec->vbr_flag = 1; ec->vbr_mnr = atoi(opt + 2); break;
        case 'l': ec->vbr_br_limit = atoi(opt + 2); return 0;
        case 'a': return 0;                                    // mpeg_select is an init argument, not E_CONTROL
        case 'w': return 0;                                    // mnr_adjust file: parsed but unused by the live path
        default: return -1;
    }
    // the CLI derives vbr_flag from the final bitrate after all options (test/tomp3.cpp:563-566)
    ec->vbr_flag = (ec->bitrate < 0) ? 1 : 0;
    return 0;
}

void fixed_polyphase_tables(float *polyA, float *polyB, float *dct32) {
    for (int i = 0; i < 256; i++) {
        polyA[i] = bits_f(kPolyA_bits[i]);
        polyB[i] = bits_f(kPolyB_bits[i]);
    }
    const double pi = 4.0 * atan(1.0);  // 32-point DCT twiddles (sbt.c:113-131), as in build_tables
    int n = 16, kk = 0;
    for (int i = 0; i < 5; i++, n = n / 2)
        for (int p = 0; p < n; p++, kk++) {
            const double t = (pi / (4 * n)) * (2 * p + 1);
            dct32[kk] = (float)(2.0 * cos(t));
        }
}

int build_tables(const hmp3_control *ec_arg, EncTables *Tp, int *unsupported) {
    EncTables &T = *Tp;
    memset(&T, 0, sizeof(T));
    EncConfig &C = T.cfg;
    if (unsupported) *unsupported = 0;
    hmp3_control ec = *ec_arg;

    // ---- defaults and clamps (mp3enc.cpp:289-378)
    if (ec.mode < 0) ec.mode = 1;
    if (ec.mode > 3) ec.mode = 3;
    if (ec.bitrate < 0) {
        ec.bitrate = 64;
        if (ec.samprate < 32000) ec.bitrate = 32;
    }
    if (ec.mode == 2) ec.vbr_flag = 0;
    if (ec.vbr_mnr < 0) ec.vbr_mnr = 0;
    if (ec.vbr_mnr > 150) ec.vbr_mnr = 150;
    if (ec.mode != 1) ec.nsbstereo = 0;
    if (ec.vbr_flag) ec.nsbstereo = 0;
    if (ec.mode == 2) ec.hf_flag = 0;
    if (ec.vbr_flag == 0) {
        if (ec.bitrate < 96) ec.hf_flag = 0;
    } else {
        if (ec.vbr_mnr < 80) ec.hf_flag = 0;
    }
    if (ec.samprate < 44100) ec.hf_flag = 0;
    if (ec.filter_select < 0) ec.filter_select = 0;
    if ((ec.vbr_flag == 0) && (ec.samprate > 24000) && (ec.bitrate < 48)) return 0;
    ec.cr_bit &= 1;
    ec.original &= 1;
    if (ec.samprate > 32000) {
        if (ec.bitrate < 24) ec.bitrate = 24;
    } else if (ec.samprate > 24000) {
        if (ec.bitrate < 16) ec.bitrate = 16;
    } else if (ec.samprate > 16000) {
        if (ec.bitrate < 12) ec.bitrate = 12;
    } else {
        if (ec.bitrate < 8) ec.bitrate = 8;
    }
    C.short_block_threshold = ec.short_block_threshold;

    // ---- header (setup.c:191-289)
    int h_option = 4 - ec.layer;
    if (h_option > 3) h_option = 3;
    if (h_option < 1) h_option = 1;
    int k = 0, dmin = 99999;
    for (int i = 0; i < 8; i++) {
        int d = abs(ec.samprate - kRates[i]);
        if (d < dmin) { dmin = d; k = i; }
    }
    int h_id = k >> 2, sr_index = k & 3;
    int h_mode = ec.mode;
    int h_mode_ext = 0;
    if (h_mode == 1) h_mode_ext = ec.nsbstereo / 4 - 1;
    if (h_mode_ext < 0) {
        h_mode_ext = 0;
        if (h_id == 0) h_mode_ext = 1;
    }
    if (h_mode_ext > 3) h_mode_ext = 3;
    int bitrate = ec.bitrate;
    static const int min_br[4][2] = {{4, 8}, {4, 8}, {4, 8}, {16, 32}};
    static const int max_br[4][2] = {{160, 320}, {160, 320}, {160, 384}, {256, 448}};
    if (bitrate < min_br[h_option][h_id]) bitrate = min_br[h_option][h_id];
    if (ec.mode != 3) bitrate = 2 * bitrate;
    if (bitrate > max_br[h_option][h_id]) bitrate = max_br[h_option][h_id];
    int br_index = 0;
    if (h_option != 1) return 0;  // Layer III only (mp3enc.cpp:388-389)
    for (int i = 1;; i++) {
        if (kBitrates[h_id][i] < 0) break;
        if (kBitrates[h_id][i] == bitrate) br_index = i;
    }
    int totbitrate = bitrate;
    // 4 header bytes: sync(12) id(1) layer(2) prot(1) | br(4) sr(2) pad(1) priv(1) | mode(2) ext(2) cr(1) orig(1) emph(2)
    {
        uint32_t h = 0xFFFu;
        h = (h << 1) | (uint32_t)h_id;
        h = (h << 2) | (uint32_t)h_option;
        h = (h << 1) | 1u;
        h = (h << 4) | (uint32_t)br_index;
        h = (h << 2) | (uint32_t)sr_index;
        h = (h << 1) | 0u;
        h = (h << 1) | 0u;
        h = (h << 2) | (uint32_t)h_mode;
        h = (h << 2) | (uint32_t)h_mode_ext;
        h = (h << 1) | (uint32_t)ec.cr_bit;
        h = (h << 1) | (uint32_t)ec.original;
        h = (h << 2) | 0u;
        C.head[0] = (unsigned char)(h >> 24);
        C.head[1] = (unsigned char)(h >> 16);
        C.head[2] = (unsigned char)(h >> 8);
        C.head[3] = (unsigned char)h;
    }
    C.h_mode = h_mode;
    C.h_id = h_id;
    C.sr_index = sr_index;
    C.br_index = br_index;
    C.totbitrate = totbitrate;
    int monodual = (h_mode == 3) ? 0 : 1;
    C.nchan = monodual + 1;
    C.mono = !monodual;
    const BandEdges &E = kBandEdges[h_id][sr_index];
    C.nband = E.l[21];
    C.nsb = (C.nband + 17) / 18;

    int nsbstereo;
    if (h_id == 0) {
        nsbstereo = 7 * totbitrate / 16 - 7;
        nsbstereo = imin(nsbstereo, 32);
        nsbstereo = imax(nsbstereo, 3);
        if (totbitrate >= 48) nsbstereo = 32;
    } else {
        nsbstereo = 12 * totbitrate / 32 - 20;
        nsbstereo = imin(nsbstereo, 32);
        nsbstereo = imax(nsbstereo, 3);
        if (totbitrate >= 96) nsbstereo = 32;
    }
    if (ec.vbr_flag) nsbstereo = 32;
    if (ec.nsbstereo > 0) {
        nsbstereo = ec.nsbstereo;
        if (nsbstereo < 3) nsbstereo = 3;
        if (nsbstereo > 32) nsbstereo = 32;
    }
    if (nsbstereo > C.nsb) nsbstereo = C.nsb;
    int samprate = kRates[4 * h_id + sr_index];
    C.samprate = samprate;
    C.pad_divisor = samprate;

    // ---- frame geometry (mp3enc.cpp:447-486)
    C.sf_bit_max = 3 * (6 * 4 + 6 * 3);
    if (h_id == 1) {
        C.frame_driver = FD_CBR_MPEG1;
        C.framebytes = 144000 * totbitrate / samprate;
        C.pad_remainder = (144000 * totbitrate) % samprate;
        C.side_bytes = (h_mode == 3) ? 17 : 32;
        C.main_framebytes = C.framebytes - 4 - C.side_bytes;
        C.ave_target_bits = 8 * C.main_framebytes / 2;
        if (h_mode != 3) C.ave_target_bits >>= 1;
        C.ave_target_bits -= C.sf_bit_max;
        C.reservoir_back = 511;
        C.granules_per_frame = 2;
    } else {
        C.frame_driver = FD_CBR_MPEG2;
        C.framebytes = (144000 / 2) * totbitrate / samprate;
        C.pad_remainder = ((144000 / 2) * totbitrate) % samprate;
        C.side_bytes = (h_mode == 3) ? 9 : 17;
        C.main_framebytes = C.framebytes - 4 - C.side_bytes;
        C.ave_target_bits = 8 * C.main_framebytes;
        if (h_mode != 3) C.ave_target_bits >>= 1;
        C.ave_target_bits -= C.sf_bit_max;
        C.reservoir_back = 255;
        C.granules_per_frame = 1;
    }

    // ---- bandwidth (mp3enc.cpp:488-590)
    int nsb_user_flag = 0, nsb_limit_user1 = 32, nsb_limit_user2 = 32;
    if (ec.nsb_limit > 0) {
        nsb_limit_user1 = imin(ec.nsb_limit, 32);
        nsb_limit_user1 = imax(ec.nsb_limit, (64 * 1000 + samprate / 2) / samprate);
        nsb_user_flag = 1;
    }
    if (ec.freq_limit < 24000) {
        nsb_limit_user2 = (64 * imax(ec.freq_limit, 1000) + samprate / 2) / samprate;
        nsb_user_flag = 1;
    }
    int nsb_limit_user = imin(nsb_limit_user1, nsb_limit_user2);
    int freq_limit;
    if (ec.vbr_flag) {
        if (h_id == 1) {
            freq_limit = 12000 + 80 * ec.vbr_mnr;
            if (ec.vbr_mnr <= 5) freq_limit = 12000;
            freq_limit = imin(freq_limit, ((int)((0.96f * 0.5f) * samprate)));
        } else {
            freq_limit = 7500 + 50 * ec.vbr_mnr;
            if (ec.vbr_mnr <= 5) freq_limit = 7500;
            freq_limit = imin(freq_limit, ((int)((0.96f * 0.5f) * samprate)));
            freq_limit = nearest_band_freq(sr_index, h_id, freq_limit);
            int tmp = (64 * freq_limit + (samprate / 2)) / samprate;
            freq_limit = (tmp * samprate) / 64;
        }
    } else {
        freq_limit = cbr_freq_limit(samprate, totbitrate, h_mode);
        if (h_id == 0) {
            freq_limit = nearest_band_freq(sr_index, h_id, freq_limit);
            int tmp = (64 * freq_limit + (samprate / 2)) / samprate;
            freq_limit = (tmp * samprate) / 64;
        }
    }
    int nsb_limit;
    if (nsb_user_flag) nsb_limit = nsb_limit_user;
    else nsb_limit = (64 * imax(freq_limit, 1000) + samprate / 2) / samprate;
    nsb_limit = imin(C.nsb, nsb_limit);
    C.nsb_limit = nsb_limit;
    C.nsb_hybrid = C.nsb_limit_ms1 = nsb_limit;
    if (nsb_limit < C.nsb) ec.hf_flag = 0;
    if (ec.hf_flag) {
        C.nsb_hybrid = 29;
        if (nsb_user_flag) C.nsb_hybrid = imin(nsb_limit_user, 29);
    }
    if (ec.hf_flag & 2) {
        C.nsb_limit_ms1 = 29;
        if (nsb_user_flag) C.nsb_limit_ms1 = imin(nsb_limit_user, 29);
    }
    C.band_limit = 18 * nsb_limit;
    if (C.band_limit > C.nband) C.band_limit = C.nband;
    int nsbstereo_limit = imin(nsbstereo, nsb_limit);
    if (h_mode == 1) C.band_limit_stereo = 18 * nsbstereo_limit;
    else C.band_limit_stereo = C.band_limit;
    if (C.band_limit_stereo > C.band_limit) C.band_limit_stereo = C.band_limit;

    // ---- input filter (filter2.c:60-76)
    C.dc_alpha = (float)(0.001 * 44100.0 / samprate);
    C.filter_select = ec.filter_select > 1 ? 1 : ec.filter_select;

    // ---- stereo processing mode (mp3enc.cpp:623-637)
    C.ms_flag = C.is_flag = 0;
    if (h_mode == 1) {
        if (nsbstereo_limit < nsb_limit) C.is_flag = 1;
        C.ms_flag = 1;
    }
    if (C.is_flag) ec.vbr_flag = 0;
    C.vbr_flag = ec.vbr_flag;

    // ---- VBR frame-size ladder (mp3enc.cpp:964-1041)
    C.ivbr_min = 1;
    C.ivbr_max = 14;
    C.vbr_pool_target = 255;
    if (ec.vbr_flag) {
        int max_tot = (ec.mode == 3 ? 1 : 2) * ec.vbr_br_limit;
        int i;
        if (h_id == 1) {
            for (i = 1; i < 15; i++) C.vbr_main_framebytes[i] = 144000 * kBitrates[1][i] / samprate - 4 - C.side_bytes;
            C.vbr_main_framebytes[15] = 9999999;
            C.vbr_pool_target = 256;
            for (i = 14; i >= 2; i--) {
                if (max_tot >= kBitrates[1][i]) break;
                C.vbr_pool_target = (C.vbr_pool_target + 511) >> 1;
            }
            C.ivbr_max = i;
            C.ave_target_bits = (8 * C.vbr_main_framebytes[C.ivbr_max] / (2 * C.nchan)) - C.sf_bit_max;
            C.frame_driver = FD_VBR_MPEG1;
        } else {
            for (i = 1; i < 15; i++) C.vbr_main_framebytes[i] = 72000 * kBitrates[0][i] / samprate - 4 - C.side_bytes;
            C.vbr_main_framebytes[15] = 9999999;
            C.vbr_pool_target = 128;
            for (i = 14; i >= 2; i--) {
                if (max_tot >= kBitrates[0][i]) break;
                C.vbr_pool_target = (C.vbr_pool_target + 255) >> 1;
            }
            C.ivbr_max = i;
            C.ave_target_bits = (8 * C.vbr_main_framebytes[C.ivbr_max] / (C.nchan)) - C.sf_bit_max;
            C.frame_driver = FD_VBR_MPEG2;
        }
    }

    // ---- quality target (mp3enc.cpp:659-685)
    int initialMNR;
    if (ec.vbr_flag) {
        initialMNR = 10 * ec.vbr_mnr;
        if (initialMNR < 210) initialMNR = 210;
        if (initialMNR > 1500) initialMNR = 1500;
    } else {
        int tmp = totbitrate / C.nchan;
        if (h_id == 1) initialMNR = 125 * (tmp - 32) / 8;
        else initialMNR = 10 * ((30 * tmp) / 8 - 70);
        if (initialMNR < 0) initialMNR = 0;
        if (initialMNR > 1000) initialMNR = 1000;
    }
    ec.vbr_delta_mnr = imin(ec.vbr_delta_mnr, 50);
    ec.vbr_delta_mnr = imax(ec.vbr_delta_mnr, -40);

    // ---- allocator selection (mp3enc.cpp:696-766): dual channel and intensity stereo take CBitAllo1
    C.allocator = (h_mode == 2 || (h_mode == 1 && C.is_flag)) ? 1 : 0;
    C.hf_flag = ec.hf_flag;
    int mnr_bias = 10 * ec.vbr_delta_mnr;
    C.nt_flatten = ec.test1 < 0 ? 6 : ec.test1;

    // ---- scale-factor band tables (bitallo3.cpp:326-361)
    for (int i = 0; i < 22; i++) T.nBand_l[i] = T.nBand_l_iso[i] = E.l[i + 1] - E.l[i];
    for (int i = 0; i < 13; i++) T.nBand_s[i] = E.s[i + 1] - E.s[i];
    C.nsf[0] = C.nsf2[0] = C.nsf3[0] = sfb_limit_long(E, C.band_limit);
    C.nsf[1] = C.nsf2[1] = C.nsf3[1] = sfb_limit_long(E, C.band_limit_stereo);
    if (C.hf_flag) {
        C.nsf2[0] = 22;
        T.nBand_l[21] = 100;
    }
    if (C.hf_flag & 2) C.nsf3[0] = C.nsf3[1] = 22;
    {
        int kk = 0, i;
        for (i = 0; i < 22; i++) { T.startBand_l[i] = kk; kk += T.nBand_l[i]; }
        T.startBand_l[22] = kk;
        T.startBand_l[23] = 576;
        for (i = 0; i < 576; i++) T.line_band_l[i] = 22;
        for (i = 0; i < 22; i++)
            for (int k = T.startBand_l[i]; k < T.startBand_l[i + 1] && k < 576; k++) T.line_band_l[k] = (unsigned char)i;
        for (int k = 0; k < 576; k++) {
            const int b = T.line_band_l[k], k0 = k & ~31;
            const int s = b < 22 ? T.startBand_l[b] : T.startBand_l[22], e = b < 22 ? T.startBand_l[b + 1] : 576;
            const int lo = std::max(s - k0, 0), hi = std::min(e - k0, 32);
            T.line_seg_l[k] = (hi - lo >= 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
            T.line_segflag_l[k] = (unsigned char)(((k - k0) == lo ? 1 : 0) | (e <= k0 + 32 ? 2 : 0));
        }
        kk = 0;
        for (i = 0; i < 13; i++) { T.startBand_s[i] = kk; kk += T.nBand_s[i]; }
        T.startBand_s[13] = kk;
    }
    for (int c = 0; c < 2; c++) {
        C.nbmax[c] = C.nbmax2[c] = C.nbmax3[c] = T.startBand_l[C.nsf[c]];
    }
    if (C.hf_flag) C.nbmax2[0] = T.startBand_l[C.nsf2[0]];
    if (C.hf_flag & 2) {
        C.nbmax3[0] = T.startBand_l[C.nsf3[0]];
        C.nbmax3[1] = T.startBand_l[C.nsf3[1]];
    }
    // short-block band limits (bitallos.cpp:146-158)
    C.nsf_s[0] = sfb_limit_short(E, C.band_limit / 3 - 10);
    C.nsf_s[1] = sfb_limit_short(E, C.band_limit_stereo / 3 - 10);
    C.nbmax_s[0] = T.startBand_s[C.nsf_s[0]];
    C.nbmax_s[1] = T.startBand_s[C.nsf_s[1]];
    C.nsf_stereo = C.nsf[1];
    C.ill_is_pos = h_id ? 7 : 999;

    // ---- quantiser step tables (bitallo3.cpp:366-380); offset 8 = the allocator's gain origin
    for (int i = 0; i < 128; i++) {
        T.gain[i] = (float)(pow(2.0, 0.25 * (i - 8)));
        T.igain34[i] = (float)(1.0 / pow((double)T.gain[i], (double)(3.0 / 4.0)));
    }
    for (int i = 0; i < 256; i++) T.ix43[i] = (float)(i * pow((double)i, (double)(1.0 / 3.0)));
    for (int i = 0; i < 21; i++) {
        float db = (float)(10.0 * log10((double)(float)(double)T.nBand_l[i]));
        T.log_cbw_l[i] = (int)(100.0f * db);
    }
    T.log_cbw_l[21] = 0;
    for (int i = 0; i < 12; i++) {
        float db = (float)(10.0 * log10((double)(float)T.nBand_s[i]));
        T.log_cbw_s[i] = (int)(100.0f * db);
    }
    // ---- CBitAllo1::BitAlloInit (bitallo1.cpp:107-203): estimators, intensity positions, constants
    {
        T.a1_bits[0] = 0;
        for (int ixm = 1; ixm < 256; ixm++)
            T.a1_bits[ixm] = (int)(16 * (1.4427 * log((double)(ixm + 1)) + (ixm - 0.6) / ixm));
        double cum = 0.0f;
        for (int ix = 0; ix < 256; ix++) {
            double t = ix + 0.5;
            const double xh = t * pow(t, 1.0 / 3.0);
            t = ix;
            const double x0 = t * pow(t, 1.0 / 3.0);
            t = ix - 0.5;
            const double xl = t * pow(fabs(t), 1.0 / 3.0);
            const double dh = xh - x0, dl = xl - x0;
            const double eps = (dh * dh * dh - dl * dl * dl) / (3.0 * (xh - xl));
            cum += eps;
            const double ave = cum / (ix + 1);
            T.a1_f_ix[ix] = (float)eps;
            T.a1_f_ixmax[ix] = (float)(10.0 * log10(ave));
        }
        cum = 0.0;
        for (int i = 0; i < 256; i++) {
            const int ix = 32 * i + 16;
            double t = ix + 0.5;
            const double xh = t * pow(t, 1.0 / 3.0);
            t = ix;
            const double x0 = t * pow(t, 1.0 / 3.0);
            t = ix - 0.5;
            const double xl = t * pow(fabs(t), 1.0 / 3.0);
            const double dh = xh - x0, dl = xl - x0;
            const double eps = (dh * dh * dh - dl * dl * dl) / (3.0 * (xh - xl));
            cum += eps;
            const double ave = cum / (i + 1);
            T.a1_f_big_ix[i] = (float)(eps);
            T.a1_f_big_ixmax[i] = (float)(10.0 * log10(ave));
        }
        if (h_id) {
            const double pi = 4.0 * atan(1.0);
            for (int i = 0; i < 34; i++) T.a1_is_pos[i] = (int)(((12.0 / pi) * atan(sqrt(i / 32.0))) + .25);
        } else {
            for (int i = 0; i < 34; i++) {
                int k = (int)(-log((i + .0001) / 32.0) / log(2.0) + 0.5);
                if (k < 0) k = 0;
                if (k > 3) k = 3;
                T.a1_is_pos[i] = k + k;
            }
        }
        for (int i = 0; i < 21; i++) T.a1_log_cbw[i] = (float)(10.0 * log10((double)T.nBand_l_iso[i]));
        static const float sparse1[21] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.10f, 0.10f, 0.10f, 0.10f, 0.10f, 0.10f,
                                          0.20f, 0.30f, 0.40f, 0.50f, 0.60f, 0.70f, 0.80f, 0.90f, 1.0f, 1.5f};
        static const float sparse2[21] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.10f, 0.10f, 0.10f, 0.10f, 0.10f, 0.10f,
                                          0.20f, 0.30f, 0.40f, 0.50f, 0.50f, 0.60f, 0.70f, 0.80f, 0.90f};
        for (int i = 0; i < 21; i++) T.a1_sparse[i] = h_id ? sparse1[i] : sparse2[i];
        T.a1_gz_con1 = (float)(16.0 / (3.0 * log(2.0)));
        T.a1_gz_con2 = (float)(1 - (16.0 / (3.0 * log(2.0))) * log(.5946) + 8);
        T.a1_gz_con0 = (float)(exp((0.99 - T.a1_gz_con2) / T.a1_gz_con1));
        T.a1_con707 = (float)(1.0 / sqrt(2.0));
    }
    // noise-target taper (bitallo3.cpp:411-446); quick = -1 (default) is "truthy" => no taper
    for (int i = 0; i < 22; i++) T.taperNT[i] = 0;
    if (!ec.quick) {
        static const int gold[22] = {-5, 0, 0, 0, 0, 0, 0, 0, 0, 3, 5, 5, 5, 5, 3, 0, 0, 0, -1, -8, -10, 0};
        for (int i = 11; i < 22; i++) T.taperNT[i] = 100 + imin(150, 20 * (i - 11));
        if (ec.vbr_flag)
            for (int i = 11; i < 22; i++) T.taperNT[i] = imin(T.taperNT[i], initialMNR);
        for (int i = 0; i < 21; i++) T.taperNT[i] -= 10 * gold[i];
    }
    if (h_id == 1) C.initial_mnr = initialMNR + mnr_bias;
    else C.initial_mnr = initialMNR + mnr_bias - 300;
    for (int i = 0; i < 22; i++)
        if (T.nBand_l[i] != 0) T.rnBand_l[i] = (1.0f / T.nBand_l[i]);
    // RD-tuned rounding offsets of the "B" quantiser (l3math.c:81-114)
    {
        static const float r[32] = {0.09460f, 0.02799f, 0.01671f, 0.01192f, 0.00927f, 0.00758f, 0.00641f, 0.00556f,
                                    0.00490f, 0.00439f, 0.00397f, 0.00362f, 0.00333f, 0.00309f, 0.00287f, 0.00269f,
                                    0.00253f, 0.00238f, 0.00225f, 0.00214f, 0.00203f, 0.00194f, 0.00185f, 0.00177f,
                                    0.00170f, 0.00163f, 0.00157f, 0.00152f, 0.00146f, 0.00141f, 0.00136f, 0.00132f};
        for (int i = 0; i < 32; i++) T.quantB_round[i] = r[i] - 0.4375f;
    }

    // ---- fixed LUTs
    for (int i = 0; i < 32; i++)
        for (int j = 0; j < 8; j++) {
            T.polyA[i][j] = bits_f(kPolyA_bits[i * 8 + j]);
            T.polyB[i][j] = bits_f(kPolyB_bits[i * 8 + j]);
        }
    for (int i = 0; i < 32; i++) { T.polyIa[i] = kPolyIa[i]; T.polyIb[i] = kPolyIb[i]; }
    for (int i = 0; i < 256; i++) {
        T.logmb[i] = kLogMb[i];
        T.exp_hi[i] = bits_f(kExpHi_bits[i]);
        T.exp_lo[i] = bits_f(kExpLo_bits[i]);
        T.p34_exp[i] = bits_f(kP34Exp_bits[i]);
    }
    for (int i = 0; i < 84; i++) T.logsub[i] = kLogSub[i];
    for (int i = 0; i < 32; i++) T.p34_seg[i] = bits_f(kP34Seg_bits[i]);
    for (int b = 0; b < HMP3_NBOOKS; b++)
        for (int i = 0; i < 256; i++) T.huff_book[b][i] = kHuffBook[b * 256 + i];
    for (int i = 0; i < 32; i++) { T.huff_sel_book[i] = kHuffSelBook[i]; T.huff_linbits[i] = kHuffLinbits[i]; }
    build_count_luts(T);

    // ---- 32-point DCT twiddles (sbt.c:113-131)
    {
        double pi = 4.0 * atan(1.0);
        int n = 16, kk = 0;
        for (int i = 0; i < 5; i++, n = n / 2)
            for (int p = 0; p < n; p++, kk++) {
                double t = (pi / (4 * n)) * (2 * p + 1);
                T.dct32[kk] = (float)(2.0 * cos(t));
            }
    }
    // ---- alias-reduction butterflies (l3init.c:130-132, 186-192)
    {
        static const float Ci[8] = {-0.6f, -0.535f, -0.33f, -0.185f, -0.095f, -0.041f, -0.0142f, -0.0037f};
        for (int i = 0; i < 8; i++) {
            T.csa[0][i] = (float)(1.0 / sqrt(1.0 + Ci[i] * Ci[i]));
            T.csa[1][i] = (float)(Ci[i] / sqrt(1.0 + Ci[i] * Ci[i]));
        }
    }
    // ---- MDCT twiddles (l3init.c:290-345)
    {
        double pi = 4.0 * atan(1.0);
        int n = 18;
        double t = pi / (4 * n);
        for (int p = 0; p < n; p++) T.m18_w[p] = (float)(2.0 * cos(t * (2 * p + 1)));
        for (int p = 0; p < 9; p++) T.m18_w2[p] = (float)2.0 * cos(2 * t * (2 * p + 1));
        t = pi / (2 * n);
        for (int kq = 0; kq < 9; kq++)
            for (int p = 0; p < 4; p++) T.m18_c[kq][p] = (float)cos(t * (2 * kq) * (2 * p + 1));
        n = 6;
        t = pi / (4 * n);
        for (int p = 0; p < n; p++) T.m6_v[p] = (float)2.0 * cos(t * (2 * p + 1));
        for (int p = 0; p < 3; p++) T.m6_v2[p] = (float)2.0 * cos(2 * t * (2 * p + 1));
        t = pi / (2 * n);
        T.m6_c = (float)cos(t * (2 * 1) * (2 * 0 + 1));
        for (int p = 0; p < 6; p++) T.m6_v[p] = T.m6_v[p] / 2.0f;
        T.m6_c = (float)2.0 * (T.m6_c);
    }
    // ---- hybrid windows by block type, sign-folded and scaled (l3init.c:208-280)
    {
        double pi = 4.0 * atan(1.0);
        float (*w)[36] = T.win;
        for (int i = 0; i < 36; i++) w[0][i] = (float)sin(pi / 36 * (i + 0.5));
        for (int i = 0; i < 18; i++) w[1][i] = (float)sin(pi / 36 * (i + 0.5));
        for (int i = 18; i < 24; i++) w[1][i] = 1.0F;
        for (int i = 24; i < 30; i++) w[1][i] = (float)sin(pi / 12 * (i + 0.5 - 18));
        for (int i = 30; i < 36; i++) w[1][i] = 0.0F;
        for (int i = 0; i < 6; i++) w[3][i] = 0.0F;
        for (int i = 6; i < 12; i++) w[3][i] = (float)sin(pi / 12 * (i + 0.5 - 6));
        for (int i = 12; i < 18; i++) w[3][i] = 1.0F;
        for (int i = 18; i < 36; i++) w[3][i] = (float)sin(pi / 36 * (i + 0.5));
        for (int i = 0; i < 12; i++) w[2][i] = (float)sin(pi / 12 * (i + 0.5));
        for (int i = 12; i < 36; i++) w[2][i] = 0.0F;
        for (int j = 0; j < 4; j++) {
            if (j == 2) continue;
            for (int i = 9; i < 36; i++) w[j][i] = -w[j][i];
        }
        for (int i = 3; i < 12; i++) w[2][i] = -w[2][i];
        for (int j = 0; j < 4; j++) {
            if (j == 2) continue;
            for (int i = 0; i < 36; i++) w[j][i] = (1.0f / 9.0f) * w[j][i];
        }
        for (int i = 0; i < 36; i++) w[2][i] = (1.0f / 3.0f) * w[2][i];
    }
    // ---- psychoacoustic model tables
    build_psy_long(T, sr_index, nsb_limit, h_id);
    build_psy_short(T, sr_index, nsb_limit, h_id);

    // ---- info block (mp3enc.cpp:841-866)
    C.info_nsbstereo = C.is_flag ? nsbstereo : 32;
    C.info_freq_limit = ec.hf_flag ? ec.freq_limit : nsb_limit * (samprate / 64);
    C.vbr_mnr = ec.vbr_mnr;
    C.vbr_delta_mnr = ec.vbr_delta_mnr;
    C.hf_flag_user = ec.hf_flag;
    {
        hmp3_control g = ec;
        g.mode = C.h_mode;
        g.bitrate = C.totbitrate;
        if (C.h_mode != 3) g.bitrate /= 2;
        g.samprate = samprate;
        g.nsbstereo = C.info_nsbstereo;
        g.freq_limit = C.info_freq_limit;
        g.nsb_limit = nsb_limit;
        g.layer = 3;
        static_assert(sizeof(hmp3_control) == sizeof(C.info_ec), "E_CONTROL image size");
        memcpy(C.info_ec, &g, sizeof(g));
    }
    return C.nchan * 4 * 1152;
}

}  // namespace hmp3
