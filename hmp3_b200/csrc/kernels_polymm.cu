// K1, contraction form (north_star (1)): the polyphase analysis filterbank as the dense product
//     S[slot][sb] = sum_{n=0}^{511} x_slot[n] * M[sb][n]
// on the 5th-generation tensor cores (tcgen05.mma, kind::tf32, FP32 accumulation in tensor memory) with the 3xTF32
// split  x*w ~= xh*wh + xh*wl + xl*wh  (xh, wh = the operands rounded to TF32, xl, wl = what is left).
//
// NOT the default: its rounding differs from the reference's fast algorithm (window fold + 32-point fast DCT in a
// fixed order), so the bitstream is no longer byte-identical.  It exists so that the question "does the tensor-core
// contraction beat the exact FP32 SIMT kernel" is answered by measurement (select it with
// HMP3_POLY_MODE=3xtf32 | tf32).  Replaces, in that mode, k_polyphase of kernels_analysis.cuh: same inputs (int16
// PCM), same output (cb.P, sub-band major, frequency inversion applied).
//
// The contraction is never materialised as an im2col matrix.  With the staged PCM of a run of slots written as rows of
// 32 samples, X2[r][k] = pcm[32 r + k], the 512-sample window of slot t is rows t .. t+15 of X2, so
//     S[t][sb] = sum_{c=0}^{15} sum_{k=0}^{31} X2[t + c][k] * W_c[k][sb],      W_c[k][sb] = M[sb][511 - 32 c - k]
// i.e. sixteen [128 x 32] x [32 x 32] products whose A operands are the SAME shared-memory matrix shifted down by c
// rows.  In the no-swizzle K-major operand layout (8-row x 16-byte core matrices stored back to back, stride between
// 8-row groups = 128 bytes) the address of (row, 16-byte chunk) is linear in the row, so "shifted by c rows" is just a
// start address 16 c bytes further on: one copy of the samples feeds all sixteen products.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#define HMP3_W 32
#include "analysis.h"
#include "batch_types.h"
#include "dsp_core.h"

namespace hmp3 {

constexpr int kMmSlots = 126;                  // valid slots per tile: 7 granules of one channel (M = 128, two rows idle)
constexpr int kMmRows = 144;                   // rows of X2 staged per tile (128 + 15, padded to a multiple of 8)
constexpr int kMmASize = 8 * kMmRows * 16;     // bytes of one A part: [16-byte chunk 0..7][row][4 floats]
constexpr int kMmBBlock = 8 * 32 * 16;         // bytes of one W_c part: [chunk 0..7][sb 0..31][4 floats]
constexpr int kMmBSize = 16 * kMmBBlock;       // all sixteen blocks
constexpr int kMmSmem = 2 * kMmBSize + 2 * kMmASize + 64;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): start address, leading (K) byte
// offset between the two 16-byte chunks of a K=8 step, stride byte offset between 8-row groups, version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N = 32, M = 128
constexpr uint32_t kMmIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(a), "l"(b), "r"(kMmIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ float tf32_hi(float x) {  // nearest TF32 (10 mantissa bits), ties away from zero
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// One CTA per SM, persistent over tiles.  Tile = (stream, channel, run of 7 polyphase granules).
// parts: 3 = xh*wh + xh*wl + xl*wh (3xTF32), 1 = xh*wh only (plain TF32)
__global__ void __launch_bounds__(128, 1)
    k_polyphase_mm(const EncTables *tabs, const StreamDev *st, const int16_t *pcm, const float *wmat, ChunkBufs cb, int K0,
                   int nstreams, int runs_per_stream, int parts) {
    extern __shared__ __align__(128) unsigned char smem[];
    float *sBh = (float *)smem, *sBl = (float *)(smem + kMmBSize);
    float *sAh = (float *)(smem + 2 * kMmBSize), *sAl = (float *)(smem + 2 * kMmBSize + kMmASize);
    uint64_t *bar = (uint64_t *)(smem + 2 * kMmBSize + 2 * kMmASize);
    uint32_t *tmem_slot = (uint32_t *)(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int G = cb.NG + 3;

    // the coefficient blocks, already split and laid out by the host (build_polymm_matrix)
    for (int i = tid; i < 2 * kMmBSize / 16; i += 128) ((float4 *)smem)[i] = ((const float4 *)wmat)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {  // 32 columns of tensor memory: the 128 x 32 FP32 accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;
    const uint32_t aH = smem_u32(sAh), aL = smem_u32(sAl), bH = smem_u32(sBh), bL = smem_u32(sBl);
    uint32_t phase = 0;

    const long long ntiles = (long long)nstreams * 2 * runs_per_stream;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int run = (int)(tile % runs_per_stream);
        const int ch = (int)((tile / runs_per_stream) & 1);
        const int s = (int)(tile / (2LL * runs_per_stream));
        const StreamDev sd = st[s];
        const int jj0 = run * 7;                        // first polyphase granule of the run within the chunk
        const long long j0 = (long long)K0 - 3 + jj0;   // ... its absolute index (negative: zero history)
        if (ch >= sd.nch || jj0 >= G || j0 >= sd.ngran) continue;   // uniform for the block
        const EncTables *T = tabs + sd.cfg;
        // ---- stage the samples as rows of 32: row r = samples n0 + 32 r .., split into TF32 high and low parts
        const long long n0 = 576 * j0 - 480;
        const int16_t *src = pcm + sd.pcm_off;
        for (int q = tid; q < kMmRows * 8; q += 128) {  // q = (row, 16-byte chunk): one float4 of each part
            const int r = q >> 3, c = q & 7;
            float h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const long long n = n0 + 32 * r + 4 * c + e;
                const float x = (n >= 0 && n < sd.nsamples) ? (float)src[n * sd.nch + ch] : 0.0f;
                h[e] = tf32_hi(x);
                l[e] = x - h[e];
            }
            ((float4 *)sAh)[c * kMmRows + r] = make_float4(h[0], h[1], h[2], h[3]);
            ((float4 *)sAl)[c * kMmRows + r] = make_float4(l[0], l[1], l[2], l[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> the tensor core's reads
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
        // ---- one thread issues the whole contraction: 16 row shifts x 4 K-steps x `parts` products
        if (tid == 0) {
            uint32_t acc = 0;
            for (int c = 0; c < 16; c++) {
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {
                    const uint32_t ao = 2 * ks * (kMmRows * 16) + 16 * c;          // chunk pair of the K-step, shifted c rows
                    const uint32_t bo = c * kMmBBlock + 2 * ks * (32 * 16);
                    const uint64_t dAh = umma_desc(aH + ao, kMmRows * 16, 128), dBh = umma_desc(bH + bo, 32 * 16, 128);
                    umma_tf32(tmem, dAh, dBh, acc);
                    acc = 1;
                    if (parts == 3) {
                        umma_tf32(tmem, dAh, umma_desc(bL + bo, 32 * 16, 128), 1);
                        umma_tf32(tmem, umma_desc(aL + ao, kMmRows * 16, 128), dBh, 1);
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                         : "memory");
        }
        // ---- everyone waits for the accumulator, then each thread takes its slot's 32 sub-band values
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "POLY_WAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra POLY_DONE_%=;\n\t"
            "bra POLY_WAIT_%=;\n\t"
            "POLY_DONE_%=:\n\t}\n" ::"r"(smem_u32(bar)),
            "r"(phase)
            : "memory");
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;");
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
            "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(tmem + ((uint32_t)(32 * warp) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int slot = tid;  // row of the accumulator = slot of the run
        if (slot < kMmSlots) {
            const int jr = slot / 18, t = slot - 18 * jr;
            const int jj = jj0 + jr;
            if (jj < G && j0 + jr < sd.ngran) {
                float *out = cb.P + (((long long)s * G + jj) * 2 + ch) * 576;
                const int nsb = T->cfg.nsb_hybrid;
#pragma unroll
                for (int sb = 0; sb < 32; sb++) {
                    const float y = __uint_as_float(v[sb]);
                    out[18 * sb + t] = freq_inverted(sb, t, nsb) ? -y : y;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();  // the accumulator and the sample tile are free again
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}

void launch_polyphase_mm(const EncTables *tabs, const StreamDev *st, const int16_t *pcm, const float *wmat, ChunkBufs cb,
                         int K0, int n, int parts, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_polyphase_mm, cudaFuncAttributeMaxDynamicSharedMemorySize, kMmSmem);
        configured = true;
    }
    const int G = cb.NG + 3;
    const int runs = (G + 6) / 7;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (long long)n * 2 * runs;
    const int grid = (int)(tiles < sms ? tiles : sms);
    k_polyphase_mm<<<grid, 128, kMmSmem, stream>>>(tabs, st, pcm, wmat, cb, K0, n, runs, parts);
}

// Host: the dense operator of the reference's polyphase (window fold + fast 32-point DCT, dsp_core.h polyphase_slot),
// probed with unit impulses, split into TF32 high / low parts and laid out as the kernel reads it:
// out[part][c 0..15][chunk 0..7][sb 0..31][4], value = M[sb][511 - 32 c - (4 chunk + e)].
void build_polymm_matrix(const EncTables *T, float *out /* 2 * 16 * 8 * 32 * 4 floats */) {
    static float M[32][512];
    for (int n = 0; n < 512; n++) {
        float col[32];
        auto fetch = [&](int i) -> float { return i == n ? 1.0f : 0.0f; };
        polyphase_slot(T, fetch, col, 1);
        for (int sb = 0; sb < 32; sb++) M[sb][n] = col[sb];
    }
    const size_t part = (size_t)16 * 8 * 32 * 4;
    for (int c = 0; c < 16; c++)
        for (int ck = 0; ck < 8; ck++)
            for (int sb = 0; sb < 32; sb++)
                for (int e = 0; e < 4; e++) {
                    const float w = M[sb][511 - 32 * c - (4 * ck + e)];
                    uint32_t u;
                    memcpy(&u, &w, 4);
                    u = (u + 0x1000u) & 0xFFFFE000u;
                    float h;
                    memcpy(&h, &u, 4);
                    const size_t idx = (((size_t)c * 8 + ck) * 32 + sb) * 4 + e;
                    out[idx] = h;
                    out[part + idx] = w - h;
                }
}
size_t polymm_matrix_floats() { return (size_t)2 * 16 * 8 * 32 * 4; }

}  // namespace hmp3
