// Phase B CUDA kernels (sm_100a): the per-stream serial stage (psychoacoustic stage 2, M/S decision,
// rate loop, scale-factor/Huffman packing, reservoir) and the parallel frame assembly.
#pragma once
#include <cuda_runtime.h>
#include "analysis.h"
#include "batch_types.h"
#include "rate_driver.h"

#ifdef HMP3_RATE_ALLOCATOR1
#define HMP3_RATE_KERNEL k_rate_a1
#define HMP3_RATE_WANT 1
#else
#define HMP3_RATE_KERNEL k_rate
#define HMP3_RATE_WANT 0
#endif
#ifndef HMP3_RATE_MIN_BLOCKS
#define HMP3_RATE_MIN_BLOCKS (32 / HMP3_RATE_WARPS)  // 32 warps (1024 threads) per SM => 64 registers per thread
#endif

namespace hmp3 {

// Which kernels a translation unit gets: the serial stage (size-optimised build), or the parallel passes around it
// (packing, assembly, results: ordinary -O3 build, kernels_pack.cu).  HMP3_RATE_ALLOCATOR1: only the CBitAllo1 twin.
#if defined(HMP3_RATE_PART_PACK)
#define HMP3_WANT_SERIAL 0
#define HMP3_WANT_PACK 1
#elif defined(HMP3_RATE_ALLOCATOR1)
#define HMP3_WANT_SERIAL 1
#define HMP3_WANT_PACK 0
#else
#define HMP3_WANT_SERIAL 1
#define HMP3_WANT_PACK 0
#endif
#if HMP3_WANT_SERIAL && !defined(HMP3_RATE_ALLOCATOR1)
// ---- K6: state reset, one thread per stream
__global__ void k_rate_init(const EncTables *tabs, const StreamDev *st, RateState *rs, RateCold *cold, int nstreams) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nstreams) return;
    rate_state_init(tabs + st[s].cfg, rs + s, cold + s);
}

#endif
#if HMP3_WANT_SERIAL
// ---- K6: the serial stage over one chunk of granules, one GROUP of HMP3_W lanes per stream (HMP3_W = 32: one warp
// per stream, the shipped configuration; 16 = two streams per warp, an experiment that loses once the streams
// differ): the scalar control flow of the rate loop runs uniformly on the lanes of a group, the per-line / per-band
// loops are split across them (HMP3_COOP sections of rate_*.h).
__global__ void __launch_bounds__(32 * kRateWarpsPerBlock, HMP3_RATE_MIN_BLOCKS)
    HMP3_RATE_KERNEL(const EncTables *tabs, const StreamDev *st, const StreamOut *so, RateState *rs, ChunkBufs cb,
                     unsigned char *main_buf, FrameRec *frames, int K0, int nstreams, long long *cycles) {
    const int s = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) / HMP3_W);
    // a stream belongs to this kernel or to its twin for the other allocator (kernels_rate_a1.cu)
    if (s < nstreams && tabs[st[s].cfg].cfg.allocator != HMP3_RATE_WANT) return;
#ifdef HMP3_RATE_BARRIER
    if (s >= nstreams || K0 >= st[s].ngran) {  // no stream / nothing left: still a member of the block's barriers
        if (s < nstreams && HMP3_LANE == 0) {
            cb.fr0[s] = cb.fr1[s] = rs[s].frames;
            cb.fd1[s] = rs[s].frames_done;
        }
        rate_run_chunk(nullptr, nullptr, K0, cb.NG, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, false);
        return;
    }
#else
    if (s >= nstreams) return;
#endif
    const long long t0 = clock64();
    const StreamDev sd = st[s];
    if (HMP3_LANE == 0) {  // nothing recorded in this chunk unless the loop below runs
        cb.fr0[s] = cb.fr1[s] = rs[s].frames;
        cb.fd1[s] = rs[s].frames_done;
    }
    HMP3_SYNC();
    if (K0 >= sd.ngran) return;
    const StreamOut o = so[s];
    const long long q0 = (long long)s * cb.NG;
    rate_run_chunk(tabs + sd.cfg, rs + s, K0, cb.NG, sd.ngran, sd.ngran_real, cb.gi + q0, cb.xr + q0 * 2 * 576,
                   cb.sm + q0 * 72, cb.prep + q0, cb.ms + q0, cb.pack + q0 * 2, frames + o.frames_off);
    HMP3_SYNC();
    if (HMP3_LANE == 0) {
        cb.fr1[s] = rs[s].frames;
        cb.fd1[s] = rs[s].frames_done;
        if (cycles) cycles[s] = clock64() - t0;  // load-balance diagnostics (hmp3_debug_rate_cycles)
    }
}

#endif  // HMP3_WANT_SERIAL
#if HMP3_WANT_PACK
// ---- K7a: packing pass, one BLOCK per frame recorded in this chunk, one warp per granule-channel of the frame.  The
// frame's records are staged in shared memory (cp.async), every warp ORs its granule-channel's scale-factor and
// Huffman bits into the frame's shared bit buffer at the position the part2_3_lengths before it imply, and the block
// writes the bytes out together.  (One warp per frame packed the four granule-channels one after the other: a chain of
// dependent table look-ups and warp scans four times as long, on a kernel that is bound by exactly that latency.)
constexpr int kFrameWords = 576;  // 18432 bits: four part2_3_lengths (12-bit fields, the reference can overflow them a little)
__global__ void __launch_bounds__(128)
    k_pack(const EncTables *tabs, const StreamDev *st, const StreamOut *so, ChunkBufs cb, unsigned char *main_buf,
           FrameRec *frames, int *flags, int K0, int nstreams) {
    const int s = (int)(blockIdx.x / (unsigned)cb.NG), j = (int)(blockIdx.x % (unsigned)cb.NG);  // <= NG frames per stream per chunk
    if (s >= nstreams) return;
    const int f = cb.fr0[s] + j;
    if (f >= cb.fr1[s]) return;
    const StreamOut o = so[s];
    FrameRec *fr = frames + o.frames_off + f;
    const EncTables *T = tabs + st[s].cfg;
    const PackGc *gc = cb.pack + ((long long)s * cb.NG + (fr->granule0 - K0)) * 2;
    __shared__ unsigned s_bits[kFrameWords];
    __shared__ unsigned s_gc[4][sizeof(PackGc) / 4];
    __shared__ int s_bad;
    static_assert(sizeof(PackGc) % 4 == 0, "PackGc is copied in 32-bit words");
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ngc = fr->ngr * T->cfg.nchan;
    for (int k = threadIdx.x; k < kFrameWords; k += 128) s_bits[k] = 0;
    if (threadIdx.x == 0) s_bad = 0;
    if (w < ngc) {
        const unsigned *src = (const unsigned *)(gc + w);
        for (int k = lane; k < (int)(sizeof(PackGc) / 4); k += 32)
            asm volatile("{ .reg .u64 a; cvta.to.shared.u64 a, %0; cp.async.ca.shared.global [a], [%1], 4; }" ::"l"(&s_gc[w][k]), "l"(src + k));
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const PackGc *rec = (const PackGc *)&s_gc[0][0];
    if (w < ngc) {
        int start = 0;
        for (int k = 0; k < w; k++) start += rec[k].gr.part2_3_length;
        const PackGc *p = rec + w;
        int bits = 0;
        if (p->gr.aux_not_null && start + p->gr.part2_3_length <= kFrameWords * 32 - 96) bits = pack_gc_bits(T, fr, p, w, s_bits, start);
        if (bits != p->gr.part2_3_length && lane == 0) s_bad = 1;
    }
    __syncthreads();
    unsigned char *dst = main_buf + o.main_off + fr->data_start;
    for (int i = threadIdx.x; i < fr->data_bytes; i += 128)
        dst[i] = i < kFrameWords * 4 ? (unsigned char)(s_bits[i >> 2] >> (24 - 8 * (i & 3))) : (unsigned char)0;
    if (w == 0) {
        pack_side(T, fr, rec, fr->side);
        if (s_bad && lane == 0) flags[s] = 1;
    }
}

// ---- per-stream totals after the last chunk
__global__ void k_results(const EncTables *tabs, const StreamDev *st, const StreamOut *so, const RateState *rs,
                          const FrameRec *frames, StreamResult *res, int nstreams) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nstreams) return;
    const RateState *R = rs + s;
    StreamResult r;
    r.frames = R->frames_done;
    r.frames_recorded = R->frames;
    r.finished = R->finished;
    r.out_bytes = 0;
    if (r.frames > 0) {
        const FrameRec *f = frames + so[s].frames_off + (r.frames - 1);
        r.out_bytes = (long long)f->out_off + frame_bytes(tabs + st[s].cfg, f);
    }
    res[s] = r;
}

// ---- K7: frame assembly, one warp per frame: header | side info | slice of the main-data stream
__global__ void __launch_bounds__(256) k_assemble(const EncTables *tabs, const StreamDev *st, const StreamOut *so,
                                                  const StreamResult *res, const long long *out_off,
                                                  const unsigned char *main_buf, const FrameRec *frames,
                                                  unsigned char *out, int max_frames, int nstreams, int frame_lo,
                                                  long long out_base) {
    // frames [frame_lo, res.frames) of every stream; byte `out_base` of a stream's output lands at out_off[s]
    long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int s = (int)(wid / max_frames), f = frame_lo + (int)(wid % max_frames);
    if (s >= nstreams || f >= res[s].frames) return;
    const int side = tabs[st[s].cfg].cfg.side_bytes;
    const FrameRec *fr = frames + so[s].frames_off + f;
    unsigned char *dst = out + out_off[s] + ((long long)fr->out_off - out_base);
    if (lane < 4) dst[lane] = fr->head[lane];
    if (lane < side) dst[4 + lane] = fr->side[lane];
    const unsigned char *src = main_buf + so[s].main_off + fr->main_start;
    dst += 4 + side;
    for (int k = lane; k < fr->mf_bytes; k += 32) dst[k] = src[k];
}

// ---- K7b: incremental assembly.  Frames [done_lo[s], fd1[s]) of every stream -> the stream's fixed output region
// (StreamDev::out_off); one warp per (stream, slot), slots stride over the new frames.
constexpr int kIncSlots = 64;
__global__ void __launch_bounds__(256) k_assemble_inc(const EncTables *tabs, const StreamDev *st, const StreamOut *so,
                                                      ChunkBufs cb, const int *done_lo, const unsigned char *main_buf,
                                                      const FrameRec *frames, unsigned char *out, int nstreams) {
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int s = (int)(wid / kIncSlots), j = (int)(wid % kIncSlots);
    if (s >= nstreams) return;
    const int hi = cb.fd1[s];
    const StreamDev sd = st[s];
    const int side = tabs[sd.cfg].cfg.side_bytes;
    for (int f = done_lo[s] + j; f < hi; f += kIncSlots) {
        const FrameRec *fr = frames + so[s].frames_off + f;
        if ((long long)fr->out_off + 4 + side + fr->mf_bytes > sd.out_cap) continue;  // never past the region
        unsigned char *dst = out + sd.out_off + fr->out_off;
        if (lane < 4) dst[lane] = fr->head[lane];
        if (lane < side) dst[4 + lane] = fr->side[lane];
        const unsigned char *src = main_buf + so[s].main_off + fr->main_start;
        dst += 4 + side;
        for (int k = lane; k < fr->mf_bytes; k += 32) dst[k] = src[k];
    }
}
__global__ void k_advance_inc(const EncTables *tabs, const StreamDev *st, const StreamOut *so, ChunkBufs cb, int *done_lo,
                              const FrameRec *frames, long long *bytes_done, int nstreams) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nstreams) return;
    const int hi = cb.fd1[s];
    long long nb = 0;
    if (hi > 0) {
        const FrameRec *f = frames + so[s].frames_off + (hi - 1);
        nb = (long long)f->out_off + frame_bytes(tabs + st[s].cfg, f);
    }
    bytes_done[s] = nb;
    done_lo[s] = hi;
}

#endif
}  // namespace hmp3
