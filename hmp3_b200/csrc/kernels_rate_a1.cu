// The serial-stage kernel of the CBitAllo1 configurations (dual channel, intensity stereo): the same driver code as
// k_rate with the allocator-1 branch compiled in (rate_driver.h, rate_allo1.h).  A separate kernel so that the code of
// the common path's k_rate -- whose speed is set by its instruction footprint -- is untouched by it.
#define HMP3_RATE_ALLOCATOR1 1
#include "kernels_rate.cuh"

#ifndef HMP3_W_HOST_VALUE
#define HMP3_W_HOST_VALUE 32
#endif

namespace hmp3 {
void launch_rate_a1(const EncTables *tabs, const StreamDev *st, const StreamOut *so, RateState *rs, ChunkBufs cb,
                    unsigned char *main_buf, FrameRec *frames, int K0, int n, cudaStream_t stream) {
    constexpr size_t smem = sizeof(float) * 576 * kRateWarpsPerBlock * (32 / HMP3_W_HOST_VALUE);
    const long long threads = (long long)n * HMP3_W_HOST_VALUE;
    const int bs = 32 * kRateWarpsPerBlock;
    k_rate_a1<<<(unsigned)((threads + bs - 1) / bs), bs, smem, stream>>>(tabs, st, so, rs, cb, main_buf, frames, K0, n, nullptr);
}
}  // namespace hmp3
