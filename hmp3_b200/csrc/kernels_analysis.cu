// Phase A kernels + launchers (see kernels_analysis.cuh).
#define HMP3_W 32  // cooperative sections in this translation unit are warp-wide
#include "kernels_analysis.cuh"

#include <atomic>
#include <mutex>
#include <stdlib.h>
#include "enc_init.h"
namespace hmp3 {
static inline unsigned blocks_for(long long items, int bs) { return (unsigned)((items + bs - 1) / bs); }

// HMP3_PHASEA_CARVEOUT=pct (experiments): shared-memory carve-out preference of the Phase A kernels.  With the serial
// stage's own preference (40) their blocks can share an SM with resident serial-stage blocks.
static void phase_a_configure() {
    static std::atomic<unsigned long long> configured{0};
    int dev = 0;
    cudaGetDevice(&dev);
    if ((configured.fetch_or(1ull << (dev & 63)) >> (dev & 63)) & 1ull) return;
    const char *e = getenv("HMP3_PHASEA_CARVEOUT");
    if (!e) return;
    const int pct = atoi(e);
    cudaFuncSetAttribute(k_polyphase, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_attack, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_hybrid, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_psy_stage1, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_psy_stage2, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_prepare, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

// the configuration-independent polyphase tables -> constant memory, once per device
static void polyphase_constants() {
    static std::mutex mu;
    static unsigned long long done = 0;  // one bit per device
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if ((done >> (dev & 63)) & 1ull) return;
    done |= 1ull << (dev & 63);
    float A[256], B[256], D[31];
    fixed_polyphase_tables(A, B, D);
    cudaMemcpyToSymbol(c_polyA, A, sizeof(A));
    cudaMemcpyToSymbol(c_polyB, B, sizeof(B));
    cudaMemcpyToSymbol(c_dct32, D, sizeof(D));
}
void launch_polyphase(const EncTables *tabs, const StreamDev *st, const int16_t *pcm, const float *pcmf, ChunkBufs cb,
                      int K0, int n, cudaStream_t stream) {
    const int G = cb.NG + 3;
    phase_a_configure();
    polyphase_constants();
    k_polyphase<<<dim3((unsigned)((G + kPolyRun - 1) / kPolyRun), (unsigned)n), 256, 0, stream>>>(tabs, st, pcm, pcmf, cb,
                                                                                                 K0, n);
}
void launch_dc_filter(const EncTables *tabs, const StreamDev *st, const int16_t *pcm, float *pcmf, float *dc,
                      long long lo, long long hi, int n, cudaStream_t stream) {
    k_dc_filter<<<blocks_for(2LL * n, 64), 64, 0, stream>>>(tabs, st, pcm, pcmf, dc, lo, hi, n);
}
void launch_attack(const EncTables *tabs, const StreamDev *st, ChunkBufs cb, int K0, int n, cudaStream_t stream) {
    const long long G = cb.NG + 3;
    k_attack<<<blocks_for((long long)n * G * 2 * 9, 128), 128, 0, stream>>>(tabs, st, cb, K0, n);
}
void launch_switch_scan(const EncTables *tabs, const StreamDev *st, SwitchState *sw, ChunkBufs cb, int K0, int n,
                        cudaStream_t stream) {
    k_switch_scan<<<blocks_for(n, 32), 32, 0, stream>>>(tabs, st, sw, cb, K0, n);
}
void launch_hybrid(const EncTables *tabs, const StreamDev *st, ChunkBufs cb, int K0, int n, cudaStream_t stream) {
    k_hybrid<<<blocks_for((long long)n * cb.NG * 2 * 32, 128), 128, 0, stream>>>(tabs, st, cb, K0, n);
}
void launch_psy_stage1(const EncTables *tabs, const StreamDev *st, ChunkBufs cb, int K0, int n, cudaStream_t stream) {
    k_psy_stage1<<<blocks_for((long long)n * cb.NG * 32, 128), 128, 0, stream>>>(tabs, st, cb, K0, n);
}
void launch_ms_scan(const EncTables *tabs, const StreamDev *st, int *msmem, ChunkBufs cb, int K0, int n,
                    cudaStream_t stream) {
    k_ms_scan<<<blocks_for(n, 64), 64, 0, stream>>>(tabs, st, msmem, cb, K0, n);
}
void launch_psy_stage2(const EncTables *tabs, const StreamDev *st, PsyState *psy, ChunkBufs cb, int K0, int n,
                       cudaStream_t stream) {
    k_psy_stage2<<<blocks_for(2LL * n * 32, 128), 128, 0, stream>>>(tabs, st, psy, cb, K0, n);
}
void launch_prepare(const EncTables *tabs, const StreamDev *st, ChunkBufs cb, int K0, int n, cudaStream_t stream) {
    k_prepare<<<blocks_for((long long)n * cb.NG * 32, 128), 128, 0, stream>>>(tabs, st, cb, K0, n);
}

// FP32 issue-rate microbenchmark (SURVEY 8d / BASELINE.md 3): the ceilings the exact-order kernels are quoted against.
// mode 0: dependent FFMA chains, 8 per thread (2 flop per lane per issue); mode 1: alternating FMUL / FADD chains, the
// instruction mix of code compiled --fmad=false (1 flop per lane per issue).
template <int MODE>
__global__ void __launch_bounds__(256) k_fp32_peak(float *out, int iters, float a, float b) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = (float)(threadIdx.x + j) * 1.0e-3f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (MODE == 0) v[j] = __fmaf_rn(v[j], a, b);
            else v[j] = __fadd_rn(__fmul_rn(v[j], a), b);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; j++) s += v[j];
    if (s == 123.456f) out[0] = s;  // never true: keeps the chains alive
}
int fp32_peak(int device, float *ffma_tflops, float *nonfused_tflops) {
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, device) != cudaSuccess) return -1;
    float *d = nullptr;
    if (cudaMalloc(&d, 16) != cudaSuccess) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 1 << 15, blocks = pr.multiProcessorCount * 8;
    float best[2] = {0, 0};
    for (int mode = 0; mode < 2; mode++)
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k_fp32_peak<0><<<blocks, 256>>>(d, iters, 0.999f, 1.0e-3f);
            else k_fp32_peak<1><<<blocks, 256>>>(d, iters, 0.999f, 1.0e-3f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double flop = (double)blocks * 256 * iters * 8 * 2;  // mul + add per element either way
            const float tf = (float)(flop / (ms * 1e-3) / 1e12);
            if (rep > 0 && tf > best[mode]) best[mode] = tf;
        }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (ffma_tflops) *ffma_tflops = best[0];
    if (nonfused_tflops) *nonfused_tflops = best[1];
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
void launch_prepare_init(int *msmem, PsyState *psy, int n, cudaStream_t stream) {
    k_prepare_init<<<blocks_for(2LL * n, 128), 128, 0, stream>>>(msmem, psy, n);
}
size_t sizeof_prep_granule() { return sizeof(PrepGranule); }
size_t sizeof_psy_state() { return sizeof(PsyState); }
}  // namespace hmp3
