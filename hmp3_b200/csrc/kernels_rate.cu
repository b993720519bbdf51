// Serial-stage kernels + launchers (see kernels_rate.cuh).  This translation unit is compiled for code
// SIZE (the rate loop is tens of thousands of instructions of branchy scalar code executed once per
// granule: instruction fetch, not arithmetic, bounds it).
#include <atomic>
#include "kernels_rate.cuh"

// lanes per stream of the serial stage (HMP3_W is only defined in device code)
#ifndef HMP3_W_HOST_VALUE
#define HMP3_W_HOST_VALUE 32
#endif
static constexpr int HMP3_W_HOST = HMP3_W_HOST_VALUE;
// one 576-float scratch row per stream of a block (rate_scratch_row)
static constexpr size_t kRateSmem = sizeof(float) * 576 * hmp3::kRateWarpsPerBlock * (32 / HMP3_W_HOST_VALUE);

namespace hmp3 {
static inline unsigned blocks_for(long long items, int bs) { return (unsigned)((items + bs - 1) / bs); }

void launch_rate_init(const EncTables *tabs, const StreamDev *st, RateState *rs, void *cold, int n, cudaStream_t stream) {
    k_rate_init<<<blocks_for(n, 64), 64, 0, stream>>>(tabs, st, rs, (RateCold *)cold, n);
}
void launch_rate(const EncTables *tabs, const StreamDev *st, const StreamOut *so, RateState *rs, ChunkBufs cb,
                 unsigned char *main_buf, FrameRec *frames, int K0, int n, cudaStream_t stream, long long *cycles) {
    // the kernel needs 8 x 10 KB of shared memory per SM; left to itself the driver configures 135 KB, at the expense
    // of L1.  Ask for the smallest carve-out that still holds eight blocks (40 % -> the 100 KB configuration: step
    // 1.63 s -> 1.55 s; 20 % costs occupancy, >= 50 % is the driver's choice again; giving the Phase A kernels the same
    // preference starves them of shared memory and is slower; halving the step search's scratch to reach the 64 KB
    // configuration gains nothing net: the two half passes cost what the extra L1 saves).
    static std::atomic<unsigned long long> configured{0};  // one bit per device: function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((configured.fetch_or(1ull << (dev & 63)) >> (dev & 63)) & 1ull)) {
        const char *e = getenv("HMP3_RATE_CARVEOUT");
        const int pct = e ? atoi(e) : 40;
        if (pct >= 0) cudaFuncSetAttribute(k_rate, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRateSmem);
    }
    k_rate<<<blocks_for((long long)n * HMP3_W_HOST, 32 * kRateWarpsPerBlock), 32 * kRateWarpsPerBlock, kRateSmem, stream>>>(
        tabs, st, so, rs, cb, main_buf, frames, K0, n, cycles);
}
// one warp, one stream (see encoder_rebase in pipeline.cu)
__global__ void k_handle_rebase(RateState *R, FrameRec *fr, unsigned char *main_buf, int dK, StreamResult *res) {
    const int lane = threadIdx.x;
    const int f0 = R->frames_done, nf = R->frames - f0;
    unsigned mbase = R->main_tot < R->mf_tot ? R->main_tot : R->mf_tot, obase = R->out_tot;
    if (nf > 0) {  // offsets grow with the frame index: the first kept frame has the smallest
        mbase = min(mbase, min(fr[f0].main_start, fr[f0].data_start));
        obase = fr[f0].out_off;
    }
    const unsigned top = R->main_tot > R->mf_tot ? R->main_tot : R->mf_tot;
    __syncwarp();
    if (mbase > 0)
        for (unsigned i = 0; i < top - mbase; i += 32) {  // forward copy; a round's reads complete before its writes
            const unsigned k = i + lane;
            unsigned char v = 0;
            if (k < top - mbase) v = main_buf[mbase + k];
            __syncwarp();
            if (k < top - mbase) main_buf[k] = v;
            __syncwarp();
        }
    if (lane == 0) {
        for (int k = 0; k < nf; k++) {
            FrameRec f = fr[f0 + k];
            f.main_start -= mbase;
            f.data_start -= mbase;
            f.out_off -= obase;
            f.granule0 -= dK;
            f.done_after -= f0;
            fr[k] = f;
        }
        R->frames = nf;
        R->frames_done = 0;
        R->main_tot -= mbase;
        R->mf_tot -= mbase;
        R->out_tot -= obase;
        R->next_granule -= dK;
        StreamResult r;
        r.frames = 0;
        r.frames_recorded = nf;
        r.finished = R->finished;
        r.out_bytes = 0;
        r.pad_ = 0;
        *res = r;
    }
}
void launch_handle_rebase(RateState *rs, FrameRec *frames, unsigned char *main_buf, int dK, StreamResult *res,
                          cudaStream_t stream) {
    k_handle_rebase<<<1, 32, 0, stream>>>(rs, frames, main_buf, dK, res);
}
size_t sizeof_rate_state() { return sizeof(RateState); }
size_t sizeof_rate_cold() { return sizeof(RateCold); }
}  // namespace hmp3
