// K6 in its phase-scheduled form (sm_100a).  A thread block owns a SET of streams (its "slots") and runs the serial
// stage's phase machine (rate_phased.h) for them: a warp claims a slot whose stream needs the phase the block is
// currently in, runs that phase for it, publishes the stream's next phase and claims again.  When no slot is left in
// the current phase the block moves to the phase most of its streams are waiting for.  The warps of an SM therefore
// walk the same few KB of code at any one time -- the serial stage is bound by instruction supply (every SM asks the
// GPC's instruction cache for a new line every ~55 cycles with its miss queue full, profiles/r2a_rate_icache_*), and
// this is the form in which an SM's instruction cache can serve all of its warps.  A stream is only ever touched by
// the warps of one block (one SM, one L1), so its state needs no device-scope ordering -- a block-scope fence before
// the slot is handed on is enough.  Streams progress independently: there is no barrier between phases.
//
// Compiled for code size like kernels_rate.cu.  HMP3_RATE_WARPS (build) = the most warps a block may have.
#include <atomic>
#include "analysis.h"
#include "batch_types.h"
#include "rate_phased.h"

namespace hmp3 {

constexpr int kPhMaxSlots = 128;       // streams per block, at most
constexpr int kPhBusy = RP_NPHASES;    // slot claimed by a warp

__global__ void __launch_bounds__(32 * kRateWarpsPerBlock, 1)
    k_rate_ph(const EncTables *tabs, const StreamDev *st, const StreamOut *so, RateState *rs, ChunkBufs cb, FrameRec *frames,
              int K0, int nstreams, int S) {
    __shared__ int s_phase[kPhMaxSlots];
    __shared__ int s_age[RP_NPHASES];
    __shared__ int s_cur;
    const int lane = threadIdx.x & 31;
    const int base = blockIdx.x * S;
    for (int j = threadIdx.x; j < kPhMaxSlots; j += blockDim.x) {
        int ph = RP_IDLE;
        const int s = base + j;
        if (j < S && s < nstreams && tabs[st[s].cfg].cfg.allocator == 0) {
            cb.fr0[s] = cb.fr1[s] = rs[s].frames;  // nothing recorded in this chunk unless the stream runs below
            cb.fd1[s] = rs[s].frames_done;
            if (K0 < st[s].ngran) {
                rs[s].ctl.phase = RP_FRAME;
                rs[s].ctl.K = K0;
                rs[s].ctl.sub = 0;
                ph = RP_FRAME;
            }
        }
        s_phase[j] = ph;
    }
    if (threadIdx.x < RP_NPHASES) s_age[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_cur = RP_FRAME;
    __syncthreads();
    const int nr = (S + 31) >> 5;  // slots per lane
    volatile int *vphase = s_phase;
    for (;;) {
        const int cur = *(volatile int *)&s_cur;
        // ---- claim a slot in the current phase
        int found = -1;
        for (int r = 0; r < nr && found < 0; r++) {
            const int j = lane + 32 * r;
            const int p = j < S ? vphase[j] : RP_IDLE;
            unsigned m = __ballot_sync(0xffffffffu, p == cur);
            while (m) {
                const int j2 = (__ffs(m) - 1) + 32 * r;
                int ok = 0;
                if (lane == 0) ok = atomicCAS(&s_phase[j2], cur, kPhBusy) == cur;
                ok = __shfl_sync(0xffffffffu, ok, 0);
                if (ok) {
                    found = j2;
                    break;
                }
                m &= m - 1;
            }
        }
        if (found < 0) {
            // ---- nothing left in this phase: move the block to the phase with the most waiting streams (a phase
            // that keeps being passed over gains weight, so a lone stream in a rare phase is not left behind)
            int pr[kPhMaxSlots / 32];
#pragma unroll
            for (int r = 0; r < kPhMaxSlots / 32; r++) {
                const int j = lane + 32 * r;
                pr[r] = j < S ? vphase[j] : RP_IDLE;
            }
            int best = -1, bestw = 0, busy = 0;
            for (int p = 0; p <= kPhBusy; p++) {
                if (p == RP_IDLE) continue;
                int c = 0;
#pragma unroll
                for (int r = 0; r < kPhMaxSlots / 32; r++) c += __popc(__ballot_sync(0xffffffffu, pr[r] == p));
                if (p == kPhBusy) {
                    busy = c;
                    break;
                }
                if (c > 0) {
                    const int w = c + *(volatile int *)&s_age[p];
                    if (w > bestw) {
                        bestw = w;
                        best = p;
                    }
                }
            }
            if (best < 0) {
                if (!busy) break;  // every stream of the block is through the chunk
                __nanosleep(400);
                continue;
            }
            if (lane == 0 && atomicCAS(&s_cur, cur, best) == cur) {
                for (int p = 0; p < RP_IDLE; p++) s_age[p] = (p == best) ? 0 : s_age[p] + 2;
            }
            __syncwarp();
            continue;
        }
        // ---- run the phase for that stream
        const int s = base + found;
        const StreamDev sd = st[s];
        const long long q0 = (long long)s * cb.NG;
        RateCtx x;
        x.T = tabs + sd.cfg;
        x.R = rs + s;
        x.K0 = K0;
        x.NG = cb.NG;
        x.ngran = sd.ngran;
        x.ngran_real = sd.ngran_real;
        x.gi = cb.gi + q0;
        x.xr = cb.xr + q0 * 2 * 576;
        x.sm = cb.sm + q0 * 72;
        x.prep = cb.prep + q0;
        x.ms = cb.ms + q0;
        x.pack = cb.pack + q0 * 2;
        x.frames = frames + so[s].frames_off;
        const int next = rate_run_phase(&x, cur);
        __syncwarp();
        if (lane == 0) {
            if (next == RP_IDLE) {
                cb.fr1[s] = rs[s].frames;
                cb.fd1[s] = rs[s].frames_done;
            }
            __threadfence_block();
            vphase[found] = next;
        }
        __syncwarp();
    }
}

static constexpr size_t kRatePhRow = sizeof(float) * 576;

void launch_rate_ph(const EncTables *tabs, const StreamDev *st, const StreamOut *so, RateState *rs, ChunkBufs cb,
                    FrameRec *frames, int K0, int n, cudaStream_t stream) {
    static std::atomic<unsigned long long> configured{0};
    static int sm_count[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((configured.fetch_or(1ull << (dev & 63)) >> (dev & 63)) & 1ull)) {
        cudaFuncSetAttribute(k_rate_ph, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kRatePhRow * kRateWarpsPerBlock));
        const char *e = getenv("HMP3_RATE_CARVEOUT");
        const int pct = e ? atoi(e) : 40;
        if (pct >= 0) cudaFuncSetAttribute(k_rate_ph, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        int v = 148;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        sm_count[dev & 63] = v;
    }
    int sms = sm_count[dev & 63];
    if (sms <= 0) sms = 148;
    // streams per block: spread the batch over all SMs; warps per block: HMP3_RATE_PH_WARPS, at most the build's
    int S = (n + sms - 1) / sms;
    if (const char *e = getenv("HMP3_RATE_PH_SLOTS")) S = atoi(e);
    if (S < 1) S = 1;
    if (S > kPhMaxSlots) S = kPhMaxSlots;
    int W = kRateWarpsPerBlock;
    if (const char *e = getenv("HMP3_RATE_PH_WARPS")) W = atoi(e);
    if (W < 1) W = 1;
    if (W > kRateWarpsPerBlock) W = kRateWarpsPerBlock;
    if (W > S) W = S;
    const unsigned blocks = (unsigned)((n + S - 1) / S);
    k_rate_ph<<<blocks, 32 * W, kRatePhRow * W, stream>>>(tabs, st, so, rs, cb, frames, K0, n, S);
}

}  // namespace hmp3
