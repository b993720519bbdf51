// K6 in its phase-scheduled form (sm_100a).  A thread block owns a SET of streams (its "slots") and runs the serial
// stage's phase machine (rate_phased.h) for them: a warp claims a slot whose stream needs the phase the block is
// currently in (per-phase bit masks of waiting slots in shared memory), runs that phase for it, publishes the stream's next phase and claims again.  When no slot is left in
// the current phase the block moves to the phase most of its streams are waiting for.  The warps of an SM therefore
// walk the same few KB of code at any one time -- the serial stage is bound by instruction supply (every SM asks the
// GPC's instruction cache for a new line every ~55 cycles with its miss queue full, profiles/r2a_rate_icache_*), and
// this is the form in which an SM's instruction cache can serve all of its warps.  A stream is only ever touched by
// the warps of one block (one SM, one L1), so its state needs no device-scope ordering -- a block-scope fence before
// the slot is handed on is enough.  Streams progress independently: there is no barrier between phases.
//
// Compiled for code size like kernels_rate.cu, with HMP3_RATE_WARPS = 24 (the scratch rows are dealt by it).
#include <atomic>
#include "analysis.h"
#include "batch_types.h"
#include "rate_phased.h"

namespace hmp3 {

constexpr int kPhMaxSlots = 128;       // streams per block, at most
constexpr int kPhWords = kPhMaxSlots / 32;

// Warp / shared-memory primitives of the scheduler as inline PTX.  This translation unit is compiled with the front
// end at -O1 (code size), which does NOT inline the CUDA header wrappers (__syncwarp, __shfl_sync, atomicAnd ...): they
// become real calls, and a call inside a divergent region (lane 0 claiming a slot) costs the region its reconvergence
// point -- lane 0 and lanes 1..31 then ran on as two groups and EACH executed the phase (measured).  With the
// instructions in line the compiler's reconvergence barriers hold and bar.warp.sync reconverges for good.
__device__ __forceinline__ void w_sync() { asm volatile("bar.warp.sync 0xffffffff;" ::: "memory"); }
__device__ __forceinline__ int w_lane0(int v) {
    int r;
    asm volatile("shfl.sync.idx.b32 %0, %1, 0, 0x1f, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ int w_max(int v) {
    int r;
    asm volatile("redux.sync.max.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ unsigned w_active() {
    unsigned r;
    asm volatile("activemask.b32 %0;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned sm_addr(const void *p) {
    unsigned long long a;
    asm("cvta.to.shared.u64 %0, %1;" : "=l"(a) : "l"(p));
    return (unsigned)a;
}
__device__ __forceinline__ unsigned sm_and(unsigned *p, unsigned v) {
    unsigned r;
    asm volatile("atom.shared.and.b32 %0, [%1], %2;" : "=r"(r) : "r"(sm_addr(p)), "r"(v) : "memory");
    return r;
}
__device__ __forceinline__ void sm_or(unsigned *p, unsigned v) {
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(sm_addr(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int sm_add(int *p, int v) {
    int r;
    asm volatile("atom.shared.add.s32 %0, [%1], %2;" : "=r"(r) : "r"(sm_addr(p)), "r"(v) : "memory");
    return r;
}
__device__ __forceinline__ int sm_cas(int *p, int cmp, int v) {
    int r;
    asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(r) : "r"(sm_addr(p)), "r"(cmp), "r"(v) : "memory");
    return r;
}
__device__ __forceinline__ int sm_exch(int *p, int v) {
    int r;
    asm volatile("atom.shared.exch.b32 %0, [%1], %2;" : "=r"(r) : "r"(sm_addr(p)), "r"(v) : "memory");
    return r;
}
__device__ __forceinline__ int sm_ld(const int *p) {  // volatile load
    int r;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(r) : "r"(sm_addr(p)) : "memory");
    return r;
}
__device__ __forceinline__ int bit_count(unsigned v) {
    int r;
    asm("popc.b32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ int bit_index(unsigned one_bit) {  // index of the only set bit
    int r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(one_bit));
    return r;
}
__device__ __forceinline__ void nap_ns(unsigned ns) { asm volatile("nanosleep.u32 %0;" ::"r"(ns)); }
__device__ __forceinline__ void fence_block() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
__device__ __forceinline__ long long clock_now() {
    long long r;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(r));
    return r;
}

// L2 prefetch of [p, p + bytes): the granule inputs of a frame are touched for the first time by the serial stage
// (the Phase A kernels wrote them a chunk ago), so the frame prologue asks for them ahead of the phases that read them
__device__ __forceinline__ void prefetch_l2(const void *p, int bytes, int lane) {
    const char *q = (const char *)p;
    for (int o = lane * 128; o < bytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + o));
}

// MAXW = the most warps a block of this instance may have: 16 leaves every thread 128 registers, 24 leaves 80.  More
// warps only pay when the block has streams to spare for them (measured: 24 warps win from ~48 streams per block).
template <int MAXW>
__global__ void __launch_bounds__(32 * MAXW, 1)
    k_rate_ph(const EncTables *tabs, const StreamDev *st, const StreamOut *so, RateState *rs, ChunkBufs cb, FrameRec *frames,
              int K0, int nstreams, int S, int opts) {
    __shared__ unsigned s_ready[RP_NPHASES][kPhWords];  // bit j of a phase's mask: the stream in slot j waits for that phase
    __shared__ int s_age[RP_NPHASES];
    __shared__ int s_cur, s_active;
    __shared__ RateCtx s_ctx[kPhMaxSlots];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base = blockIdx.x * S;
    for (int k = threadIdx.x; k < RP_NPHASES * kPhWords; k += blockDim.x) (&s_ready[0][0])[k] = 0;
    for (int k = threadIdx.x; k < RP_NPHASES; k += blockDim.x) s_age[k] = 0;
    if (threadIdx.x == 0) {
        s_cur = RP_FRAME;
        s_active = 0;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
        const int s = base + j;
        if (s >= nstreams) continue;
        const StreamDev sd = st[s];
        if (tabs[sd.cfg].cfg.allocator != 0) continue;  // the CBitAllo1 kernel's stream
        cb.fr0[s] = cb.fr1[s] = rs[s].frames;  // nothing recorded in this chunk unless the stream runs below
        cb.fd1[s] = rs[s].frames_done;
        if (K0 >= sd.ngran) continue;
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
        s_ctx[j] = x;
        rate_ctl_enter_chunk(&x);
        sm_or(&s_ready[RP_FRAME][j >> 5], 1u << (j & 31));
        sm_add(&s_active, 1);
    }
    __syncthreads();
    const int nw = (S + 31) >> 5;
    unsigned nap = 256;  // idle back-off (ns): a warp with nothing to claim polls less and less often
    const long long t_start = clock_now();
    for (;;) {
        // (every decision of the scheduler is taken by lane 0 and broadcast: the lanes of a warp that wakes from a sleep
        // need not arrive here together, and the phase they run must be the one lane 0 claimed the slot for)
        w_sync();
        const int cur = w_lane0(sm_ld(&s_cur));
        // ---- claim a slot in the current phase.  All lanes walk the loop together (lane 0's values, broadcast) and only
        // the atomic itself is lane 0's: with the whole loop inside `if (lane == 0)` the warp came out of it in two
        // pieces whenever another warp had raced it to a slot (measured), and pieces do not merge again.
        int found = -1;
        for (int k = 0; k < nw && found < 0; k++) {
            const int w = (k + warp) % nw;
            unsigned m = (unsigned)w_lane0(sm_ld((const int *)&s_ready[cur][w]));
            while (m) {
                const unsigned bit = m & (0u - m);
                unsigned old = 0;
                if (lane == 0) old = sm_and(&s_ready[cur][w], ~bit);
                old = (unsigned)w_lane0((int)old);
                if (old & bit) {
                    found = 32 * w + bit_index(bit);
                    break;
                }
                m = old & ~bit;
            }
        }
        if (found < 0) {
            // ---- nothing left in this phase: move the block to the phase with the most waiting streams (a phase
            // that keeps being passed over gains weight, so a lone stream in a rare phase is not left behind)
            int c = 0;
            if (lane < RP_IDLE)
                for (int w = 0; w < nw; w++) c += bit_count((unsigned)sm_ld((const int *)&s_ready[lane][w]));
            // policy: the phase with the most waiting streams (aged), or (opts & 8) the next non-empty phase in
            // pipeline order after the current one, so that the block's streams move through the phases as a front
            int wgt = c > 0 ? c + sm_ld(&s_age[lane < RP_IDLE ? lane : 0]) : 0;
            if ((opts & 8) && c > 0) wgt = RP_IDLE - ((lane - cur - 1 + RP_IDLE) % RP_IDLE);
            const int top = w_max((wgt << 5) | lane);
            if ((top >> 5) == 0) {
                if (w_lane0(sm_ld(&s_active)) == 0) break;  // every stream of the block is through the chunk
                nap_ns(nap);
                if (nap < 8192) nap <<= 1;
                // watchdog: a launch lasts a fraction of a second (a few under a profiler's replay passes); a scheduler
                // that has waited for a minute is broken, and a failed launch is better than a hung device
                else if (!(opts & 16) && w_lane0((int)(clock_now() - t_start > 120000000000ll))) {
                    __trap();
                }
                continue;
            }
            const int best = top & 31;
            int won = 0;
            if (lane == 0) won = sm_cas(&s_cur, cur, best) == cur;
            won = w_lane0(won);
            if (won && lane < RP_IDLE) s_age[lane] = (lane == best) ? 0 : (c > 0 ? s_age[lane] + 2 : 0);
            continue;
        }
        // ---- run the phase for that stream (the two halves of the frame bookkeeping run on as one).  The warp must
        // enter a phase WHOLE: the phases run their scalar bookkeeping redundantly on all lanes (read-modify-write of
        // the stream's state in lock-step), and lane 0 comes out of the claim above on its own
        w_sync();
        nap = 256;
        const RateCtx *x = &s_ctx[found];
        int ran = cur, next = rate_run_phase(x, cur);
        // phases that run on in the same warp (the stream is not handed back to the block in between): the two halves
        // of the frame bookkeeping always; with opts & 32 / 64 also the pairs that the coarser phase set had as one
        for (;;) {
            bool on = (next == RP_FRAME && !(opts & 1));
            if (opts & 32) on = on || (ran == RP_QUANT && next == RP_COUNT);
            if (opts & 64)
                on = on || (ran == RP_SEEK && next == RP_TRADE) || (ran == RP_SF && next == RP_COARSE) ||
                     (ran == RP_REFIT && next == RP_GFIN);
            if (!on) break;
            w_sync();
            ran = next;
            next = rate_run_phase(x, ran);
        }
        if (ran == RP_FRAME && next == RP_GSTART && !(opts & 2)) {  // the inputs of the frame's granules
            const int o = x->R->ctl.K - K0;
            prefetch_l2(x->prep + o, 2 * (int)sizeof(PrepGranule), lane);
            prefetch_l2(x->xr + (long long)o * 2 * 576, 2 * 2 * 576 * 4, lane);
            prefetch_l2(x->sm + (long long)o * 72, 2 * 72 * (int)sizeof(SigMask), lane);
        }
        w_sync();
        if (lane == 0) {
            fence_block();
            if (next == RP_IDLE) {
                const int s = base + found;
                cb.fr1[s] = x->R->frames;
                cb.fd1[s] = x->R->frames_done;
                sm_add(&s_active, -1);
            } else sm_or(&s_ready[next][found >> 5], 1u << (found & 31));
        }
    }
}

static constexpr size_t kRatePhRow = sizeof(float) * 576;
static_assert(kRateWarpsPerBlock >= 24, "build with -DHMP3_RATE_WARPS >= 24: rate_scratch_row() deals the rows by it");

void launch_rate_ph(const EncTables *tabs, const StreamDev *st, const StreamOut *so, RateState *rs, ChunkBufs cb,
                    FrameRec *frames, int K0, int n, cudaStream_t stream) {
    static std::atomic<unsigned long long> configured{0};
    static int sm_count[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((configured.fetch_or(1ull << (dev & 63)) >> (dev & 63)) & 1ull)) {
        const char *e = getenv("HMP3_RATE_CARVEOUT");
        const int pct = e ? atoi(e) : 40;
        cudaFuncSetAttribute(k_rate_ph<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kRatePhRow * 16));
        cudaFuncSetAttribute(k_rate_ph<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kRatePhRow * 24));
        if (pct >= 0) {
            cudaFuncSetAttribute(k_rate_ph<16>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            cudaFuncSetAttribute(k_rate_ph<24>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        }
        int v = 148;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        sm_count[dev & 63] = v;
    }
    int sms = sm_count[dev & 63];
    if (sms <= 0) sms = 148;
    // streams per block: the batch spread over all SMs (HMP3_RATE_PH_SLOTS overrides); warps per block: 16, or 24 when
    // the block has at least 48 streams (HMP3_RATE_PH_WARPS overrides), never more than streams
    int S = (n + sms - 1) / sms;
    if (const char *e = getenv("HMP3_RATE_PH_SLOTS")) S = atoi(e);
    if (S < 1) S = 1;
    if (S > kPhMaxSlots) S = kPhMaxSlots;
    int W = S >= 48 ? 24 : 16;
    if (const char *e = getenv("HMP3_RATE_PH_WARPS")) W = atoi(e);
    if (W < 1) W = 1;
    if (W > 24) W = 24;
    if (W > S) W = S;
    const unsigned blocks = (unsigned)((n + S - 1) / S);
    int opts = 0;  // diagnostics: 1 = no chaining of the frame phases, 2 = no prefetch, 8 = pipeline-order policy, 16 = no watchdog, 32 / 64 = run quantise+count / seek+trade, sf+coarsen, refit+finish as one phase
    if (const char *e = getenv("HMP3_RATE_PH_OPTS")) opts = atoi(e);
    // (HMP3_RATE_PH_REGS80=1: the 80-register instance whatever the warp count -- leaves registers for Phase A blocks)
    if (W > 16 || getenv("HMP3_RATE_PH_REGS80")) k_rate_ph<24><<<blocks, 32 * W, kRatePhRow * W, stream>>>(tabs, st, so, rs, cb, frames, K0, n, S, opts);
    else k_rate_ph<16><<<blocks, 32 * W, kRatePhRow * W, stream>>>(tabs, st, so, rs, cb, frames, K0, n, S, opts);
}

}  // namespace hmp3
