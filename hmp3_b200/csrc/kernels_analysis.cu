// Phase A kernels + launchers (see kernels_analysis.cuh).
#define HMP3_W 32  // cooperative sections in this translation unit are warp-wide
#include "kernels_analysis.cuh"

namespace hmp3 {
static inline unsigned blocks_for(long long items, int bs) { return (unsigned)((items + bs - 1) / bs); }

void launch_polyphase(const EncTables *tabs, const StreamDev *st, const int16_t *pcm, const float *pcmf, ChunkBufs cb,
                      int K0, int n, cudaStream_t stream) {
    const int G = cb.NG + 3;
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
void launch_prepare(const EncTables *tabs, const StreamDev *st, int *msmem, PsyState *psy, ChunkBufs cb, int K0, int n,
                    cudaStream_t stream) {
    k_ms_scan<<<blocks_for(n, 64), 64, 0, stream>>>(tabs, st, msmem, cb, K0, n);
    k_psy_stage2<<<blocks_for(2LL * n * 32, 128), 128, 0, stream>>>(tabs, st, psy, cb, K0, n);
    k_prepare<<<blocks_for((long long)n * cb.NG * 32, 128), 128, 0, stream>>>(tabs, st, cb, K0, n);
}
void launch_prepare_init(int *msmem, PsyState *psy, int n, cudaStream_t stream) {
    k_prepare_init<<<blocks_for(2LL * n, 128), 128, 0, stream>>>(msmem, psy, n);
}
size_t sizeof_prep_granule() { return sizeof(PrepGranule); }
size_t sizeof_psy_state() { return sizeof(PsyState); }
}  // namespace hmp3
