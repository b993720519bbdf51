// Resolved per-configuration constants and look-up tables of the Layer III encode path.
// One EncTables object is built on the host per distinct control block (enc_init.cpp), uploaded
// once to HBM, and read by every kernel through a const pointer.  Layout is ours; the numbers
// follow the reference's init (file:line cited at each generator in enc_init.cpp).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && defined(HMP3_NO_FORCE_INLINE)
#define HMP3_HD __host__ __device__ inline
#define HMP3_FN static __host__ __device__ __noinline__
#elif defined(__CUDACC__)
#define HMP3_HD __host__ __device__ __forceinline__
#define HMP3_FN static __host__ __device__ __noinline__
#else
#define HMP3_HD inline
#define HMP3_FN static inline
#endif

// Device code of the serial stage runs one stream per GROUP of HMP3_W lanes: scalar control flow is executed
// uniformly by the lanes of a group and the per-line / per-band loops are split over them (HMP3_COOP sections).
// The product is built with HMP3_W = 32 (a warp per stream).  HMP3_W = 16 puts two streams in a warp so that they can
// share instruction fetches while they run the same code (the serial stage is instruction-fetch bound, DESIGN.md
// 7.2); they diverge too often for that to pay below ~9500 streams per GPU, so it is a build option only
// (HMP3_RATE_W=16).  Every warp-level primitive below names only the lanes of its own group.
// The host build (test-only simulator) runs the plain sequential loops.
#if defined(__CUDA_ARCH__)
#define HMP3_COOP 1
#ifndef HMP3_W
#define HMP3_W 32
#endif
#define HMP3_LANE ((int)(threadIdx.x & (unsigned)(HMP3_W - 1)))
#define HMP3_GSHIFT (threadIdx.x & 31u & ~(unsigned)(HMP3_W - 1))
#define HMP3_GMASK ((HMP3_W == 32) ? 0xffffffffu : (((1u << (HMP3_W & 31)) - 1u) << HMP3_GSHIFT))
// (inline PTX, not __syncwarp(): the serial stage is compiled with the front end at -O1, which leaves the CUDA header
// wrappers as real function calls -- lanes that arrive at such a call in separate groups synchronise inside it and
// return in separate groups again, whereas the barrier instruction in line reconverges them here)
#define HMP3_SYNC() asm volatile("bar.warp.sync %0;" ::"r"(HMP3_GMASK) : "memory")
// independent items i = 0..n-1 dealt over the lanes of the group (sequential on the host)
#define HMP3_FOR_LANES(i, n) for (int i = HMP3_LANE; i < (n); i += HMP3_W)
#else
#define HMP3_COOP 0
#define HMP3_W 1
#define HMP3_LANE 0
#define HMP3_SYNC()
#define HMP3_FOR_LANES(i, n) for (int i = 0; i < (n); i++)
#endif

#if !HMP3_COOP
namespace hmp3 {
// reductions over the lanes of a group: the host build has one "lane"
static inline int gsum(int v) { return v; }
static inline int gmax(int v) { return v; }
static inline int gor(int v) { return v; }
}
#endif

// Small constant look-up tables live at namespace scope in device memory (a function-local array would be
// rebuilt on the stack at every call and bloat the code).
#if defined(__CUDACC__)
#define HMP3_CONST_TABLE __device__ const
#else
#define HMP3_CONST_TABLE static const
#endif

namespace hmp3 {

HMP3_CONST_TABLE unsigned char kPretab[22] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 2, 0};  // ISO pretab
HMP3_CONST_TABLE unsigned char kQuadLenA[16] = {1, 4, 4, 5, 4, 6, 5, 6, 4, 5, 5, 6, 5, 6, 6, 6};   // count1 table A lengths
HMP3_CONST_TABLE unsigned char kQuadCodeA[16] = {1, 5, 4, 5, 6, 5, 4, 4, 7, 3, 6, 0, 7, 2, 3, 1};  // ... and codes
HMP3_CONST_TABLE unsigned char kRegion0[24] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 6, 6, 6, 7, 7, 7};
HMP3_CONST_TABLE unsigned char kRegion1[24] = {1, 1, 1, 1, 1, 2, 2, 2, 3, 3, 4, 5, 5, 5, 5, 6, 6, 7, 7, 7, 8, 8, 8, 8};
HMP3_CONST_TABLE unsigned char kSfcIndex[5][4] = {{0, 1, 2, 3}, {5, 5, 6, 7}, {8, 8, 9, 10}, {4, 11, 12, 13}, {14, 14, 14, 15}};
HMP3_CONST_TABLE unsigned char kSfcSlen[16][2] = {{0, 0}, {0, 1}, {0, 2}, {0, 3}, {3, 0}, {1, 1}, {1, 2}, {1, 3},
                                                  {2, 1}, {2, 2}, {2, 3}, {3, 1}, {3, 2}, {3, 3}, {4, 2}, {4, 3}};
HMP3_CONST_TABLE int kSfGroupEdge[5] = {0, 6, 11, 16, 21};
HMP3_CONST_TABLE unsigned char kPeakSnap[16] = {0, 1, 2, 3, 3, 5, 5, 7, 7, 7, 7, 15, 15, 15, 15, 15};

#if HMP3_COOP
// Warp intrinsics as inline PTX: the serial-stage translation unit is compiled at a low front-end optimisation
// level (for code size), where the CUDA header wrappers of these would become real function calls.
__device__ __forceinline__ float wshfl(float v, int src) {
    float r;
    asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=f"(r) : "f"(v), "r"(src));
    return r;
}
__device__ __forceinline__ int wshfl(int v, int src) {
    int r;
    asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(r) : "r"(v), "r"(src));
    return r;
}
__device__ __forceinline__ int wshfl_up(int v, int d) {
    int r;
    asm volatile("shfl.sync.up.b32 %0, %1, %2, 0x0, 0xffffffff;" : "=r"(r) : "r"(v), "r"(d));
    return r;
}
__device__ __forceinline__ unsigned wsum(unsigned v) {
    unsigned r;
    asm volatile("redux.sync.add.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ int wsum(int v) { return (int)wsum((unsigned)v); }
__device__ __forceinline__ int wmax(int v) {
    int r;
    asm volatile("redux.sync.max.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
// acc + v[0] + v[1] + ... + v[m-1] in lane order (m <= 32, uniform), v[k] = lane k's value; every lane gets the
// result.  Shuffles are issued four at a time so their latencies overlap; the additions stay strictly ordered.
__device__ __forceinline__ float wsum_ordered(float acc, float v, int m) {
    int k = 0;
    for (; k + 4 <= m; k += 4) {
        const float a0 = wshfl(v, k), a1 = wshfl(v, k + 1), a2 = wshfl(v, k + 2), a3 = wshfl(v, k + 3);
        acc += a0;
        acc += a1;
        acc += a2;
        acc += a3;
    }
    for (; k < m; k++) acc += wshfl(v, k);
    return acc;
}
__device__ __forceinline__ unsigned wballot(int pred) {
    unsigned r;
    asm volatile("{ .reg .pred p; setp.ne.s32 p, %1, 0; vote.sync.ballot.b32 %0, p, 0xffffffff; }" : "=r"(r) : "r"(pred));
    return r;
}
__device__ __forceinline__ void smem_or(unsigned *p, unsigned v) {
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned funnel_r(unsigned lo, unsigned hi, unsigned sh) {
    unsigned r;
    asm("shf.r.clamp.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(sh));
    return r;
}
// ---- the same primitives restricted to the calling lane's GROUP of HMP3_W lanes (serial stage)
__device__ __forceinline__ float gshfl(float v, int src) {
    float r;
    asm volatile("shfl.sync.idx.b32 %0, %1, %2, %3, %4;" : "=f"(r) : "f"(v), "r"(src), "r"(((32 - HMP3_W) << 8) | 0x1f), "r"(HMP3_GMASK));
    return r;
}
__device__ __forceinline__ int gshfl(int v, int src) {
    int r;
    asm volatile("shfl.sync.idx.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(v), "r"(src), "r"(((32 - HMP3_W) << 8) | 0x1f), "r"(HMP3_GMASK));
    return r;
}
__device__ __forceinline__ unsigned gsum(unsigned v) {
    unsigned r;
    asm volatile("redux.sync.add.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(HMP3_GMASK));
    return r;
}
__device__ __forceinline__ int gsum(int v) { return (int)gsum((unsigned)v); }
__device__ __forceinline__ int gmax(int v) {
    int r;
    asm volatile("redux.sync.max.s32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(HMP3_GMASK));
    return r;
}
__device__ __forceinline__ int gor(int v) {
    int r;
    asm volatile("redux.sync.or.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(HMP3_GMASK));
    return r;
}
__device__ __forceinline__ unsigned gballot(int pred) {  // bit i = lane i of the group
    unsigned r;
    asm volatile("{ .reg .pred p; setp.ne.s32 p, %1, 0; vote.sync.ballot.b32 %0, p, %2; }" : "=r"(r) : "r"(pred), "r"(HMP3_GMASK));
    return r >> HMP3_GSHIFT;
}
// acc + v[0] + ... + v[m-1] in lane order (m <= HMP3_W, uniform in the group); every lane gets the result
__device__ __forceinline__ float gsum_ordered(float acc, float v, int m) {
    int k = 0;
    for (; k + 4 <= m; k += 4) {
        const float a0 = gshfl(v, k), a1 = gshfl(v, k + 1), a2 = gshfl(v, k + 2), a3 = gshfl(v, k + 3);
        acc += a0;
        acc += a1;
        acc += a2;
        acc += a3;
    }
    for (; k < m; k++) acc += gshfl(v, k);
    return acc;
}
#endif

// warps per thread block of the serial-stage kernel (each warp carries 32 / HMP3_W streams)
#ifndef HMP3_RATE_WARPS
#define HMP3_RATE_WARPS 4
#endif
constexpr int kRateWarpsPerBlock = HMP3_RATE_WARPS;
// warps (= frames) per thread block of the packing kernel
constexpr int kPackWarpsPerBlock = 4;

enum FrameDriver { FD_VBR_MPEG1 = 0, FD_CBR_MPEG1 = 1, FD_VBR_MPEG2 = 2, FD_CBR_MPEG2 = 3 };

// Number of candidate-table classes of the Huffman bit counter (see build_count_luts()).
constexpr int kCountClasses = 20;

struct EncConfig {
    // stream shape
    int nchan, h_id, sr_index, samprate, h_mode, totbitrate, br_index;
    int nband, nsb, nsb_limit, nsb_hybrid /* nsb_limitMS[0] */, nsb_limit_ms1, band_limit, band_limit_stereo;
    // frame geometry / reservoir
    int frame_driver, framebytes, main_framebytes, side_bytes, pad_remainder, pad_divisor;
    int ave_target_bits, sf_bit_max, reservoir_back /* 511 MPEG-1, 255 MPEG-2 */;
    int vbr_flag, ivbr_min, ivbr_max, vbr_pool_target;
    int vbr_main_framebytes[16];
    int ms_flag, is_flag, hf_flag, nsf_stereo;
    int short_block_threshold, filter_select, mono;
    float dc_alpha;
    // rate-loop constants (CBitAllo3::BitAlloInit / CBitAlloShort::BitAlloInit)
    int initial_mnr, initial_mnr_short, nt_flatten /* test1 */;
    int nsf[2], nsf2[2], nsf3[2], nbmax[2], nbmax2[2], nbmax3[2];
    int nsf_s[2], nbmax_s[2];
    int ill_is_pos;
    int allocator;           // 0 = CBitAllo3 + CBitAlloShort; 1 = CBitAllo1 (dual channel, intensity stereo: long blocks only)
    unsigned char head[4];
    int granules_per_frame;  // 2 MPEG-1, 1 MPEG-2
    // info (ec_global, mp3enc.cpp:841-866)
    int info_freq_limit, info_nsbstereo, vbr_mnr, vbr_delta_mnr, hf_flag_user;
    int info_ec[45];  // image of the effective E_CONTROL (ec_global)
};

struct EncTables {
    EncConfig cfg;
    // ---- polyphase (window folded into 32 sums + 32-point fast DCT)
    float polyA[32][8], polyB[32][8];
    int polyIa[32], polyIb[32];
    float dct32[31];
    // ---- hybrid windows / MDCT / alias
    float win[4][36];
    float csa[2][8];
    float m18_w[18], m18_w2[9], m18_c[9][4];
    float m6_v[6], m6_v2[3], m6_c;
    // ---- integer-indexed log / exp / pow(3/4)
    int logmb[256];
    float exp_hi[256], exp_lo[256];
    int logsub[84];
    float p34_exp[256], p34_seg[32];
    // ---- psychoacoustic partitions and spreading
    int psy_npart_l, psy_npart_s;      // partitions the spreading loop visits (cntl[64].count)
    int psy_emap_n_l, psy_emap_n_s;    // partitions the energy map fills (nsum[66])
    int psy_nsum_l[64], psy_nsum_s[64];
    int psy_start_l[65], psy_start_s[65];
    int spd_cnt_l[64], spd_off_l[64], spd_w0_l[64];  // w0 = first weight index of partition i
    int spd_cnt_s[64], spd_off_s[64], spd_w0_s[64];
    float w_spd_l[2200], w_spd_s[1000];
    // ---- scale-factor bands
    int nBand_l[22], startBand_l[24], nBand_s[13], startBand_s[14];
    int nBand_l_iso[22];               // ISO widths (nBand_l[21] becomes 100 when -HF is active)
    int log_cbw_l[22], log_cbw_s[16];
    float rnBand_l[22];
    int taperNT[22];
    // ---- quantiser
    float gain[128], igain34[128], ix43[256];
    float quantB_round[32];
    // ---- Huffman
    uint32_t huff_book[16][256];  // (len << 24) | code, [book][x*16+y]
    int huff_sel_book[32], huff_linbits[32];
    // bit-count LUTs: per class two packed words per (x,y): lo16/hi16 = candidate tables 0/1 and 2/3
    uint32_t cnt_lut[kCountClasses][256][2];
    int cnt_class_of_max[24];          // ixmax 0..22 -> class (23.. handled by thresholds)
    int cnt_tables[kCountClasses][4];  // candidate Huffman table numbers (0 = none)
    int cnt_tmax[kCountClasses];       // largest value the class can code
    int cnt_ncand[kCountClasses];      // 2 or 4 (0 for the null class)
    // ---- line -> scale-factor band (long blocks), 22 = above the last band
    unsigned char line_band_l[576];
    // for the warp-wide quantiser pass (chunks of 32 lines, lane = line & 31): the lanes of a line's chunk that hold
    // lines of the same band, and flags: 1 = first such lane of the chunk, 2 = the band ends inside this chunk
    uint32_t line_seg_l[576];
    unsigned char line_segflag_l[576];
    // ---- CBitAllo1 (bitallo1.cpp:107-203, 441-542): estimators and constants of the allocator-1 configurations
    int a1_bits[256];                              // look_bits: estimated bits (x16) per line at band maximum ixmax
    float a1_f_ix[256], a1_f_ixmax[256];           // quantisation-noise estimators per value / per band maximum
    float a1_f_big_ix[256], a1_f_big_ixmax[256];   // ... for values above 255, in steps of 32
    int a1_is_pos[34];                             // intensity position from the channel energy ratio
    float a1_log_cbw[21];                          // 10 log10(band width)
    float a1_sparse[21];                           // Ssb: side-channel sparsing thresholds
    float a1_gz_con0, a1_gz_con1, a1_gz_con2, a1_con707;
};

// ------------------------------------------------------------------ scalar table functions
HMP3_HD unsigned f2u(float x) {
    union { float f; unsigned u; } v;
    v.f = x;
    return v.u;
}
HMP3_HD float u2f(unsigned x) {
    union { float f; unsigned u; } v;
    v.u = x;
    return v.f;
}

// millibel log: table on the top 8 mantissa bits + 301 per exponent step (l3math.c:227-243, IEEE_FLOAT branch)
HMP3_HD int mb_log(const EncTables *T, float x) {
    unsigned u = f2u(x);
    return T->logmb[(u >> 15) & 255] + 301 * (int)(u >> 23);
}
// millibel antilog (l3math.c:341-357, IEEE_FLOAT branch)
HMP3_HD float mb_exp(const EncTables *T, int x) {
    float t = T->exp_lo[(unsigned)x & 0xff] * T->exp_hi[((unsigned)x & 0xff00) >> 8];
    if (x > 32000) return 1.0E32f;
    if (x < -32000) return 1.0E-32f;
    return t;
}
// round half away from zero (l3math.c:360-364)
// The serial-stage translation unit is built without loop unrolling (code size), so the few hot sequential loops
// are unrolled by hand: the loads of four elements are issued together, the arithmetic keeps its order.
HMP3_HD float sum_seq(const float *v, int n, float acc) {  // acc + v[0] + v[1] + ... in that order
    int k = 0;
    for (; k + 4 <= n; k += 4) {
        const float a0 = v[k], a1 = v[k + 1], a2 = v[k + 2], a3 = v[k + 3];
        acc += a0;
        acc += a1;
        acc += a2;
        acc += a3;
    }
    for (; k < n; k++) acc += v[k];
    return acc;
}
HMP3_HD int round_away(float x) { return (int)(x + ((f2u(x) >> 31) ? -0.5f : 0.5f)); }
// 2*antilog(a) - antilog(b) in millibels (l3math.c:367-382)
HMP3_HD int mb_logsub(const EncTables *T, int a, int b) {
    int k = (a - b) >> 4;
    if (k > 83) k = 83;
    return a + T->logsub[k];
}
// |x|^(3/4), piecewise linear in the mantissa (pow34.c:131-156, IEEE_FLOAT branch); x >= 0
HMP3_HD float pow34(const EncTables *T, float x) {
    unsigned u = f2u(x);
    float mant = u2f((u & 0x7FFFFFu) | (127u << 23));
    unsigned e = u >> 19;
    unsigned e2 = (e >> 4) & 255u;
    e &= 15u;
    return (mant * T->p34_seg[2 * e + 1] + T->p34_seg[2 * e]) * T->p34_exp[e2];
}

}  // namespace hmp3
