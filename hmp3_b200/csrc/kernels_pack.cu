// The parallel passes around the serial stage: packing (one warp per frame), frame assembly, per-stream results.
// Its own translation unit so that its build flags are its own (see __graft_entry__.py: built for size like the
// serial stage, which measured faster than -O3 for the branchy bit-packing code).
#define HMP3_RATE_PART_PACK 1
#include "kernels_rate.cuh"

namespace hmp3 {
static inline unsigned blocks_for(long long items, int bs) { return (unsigned)((items + bs - 1) / bs); }

// exclusive scan of the per-stream output sizes -> compact output offsets (single block)
__global__ void k_out_offsets(const StreamResult *res, long long *out_off, int n) {
    __shared__ long long part[1024];
    const int t = threadIdx.x;
    const int per = (n + 1023) / 1024;
    const int lo = t * per, hi = min(n, lo + per);
    long long sum = 0;
    for (int i = lo; i < hi; i++) sum += res[i].out_bytes;
    part[t] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        long long v = (t >= d) ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    long long run = part[t] - sum;
    for (int i = lo; i < hi; i++) {
        out_off[i] = run;
        run += res[i].out_bytes;
    }
    if (t == 1023) out_off[n] = part[1023];
}


void launch_pack(const EncTables *tabs, const StreamDev *st, const StreamOut *so, ChunkBufs cb, unsigned char *main_buf,
                 FrameRec *frames, int *flags, int K0, int n, cudaStream_t stream) {
    k_pack<<<(unsigned)((long long)n * cb.NG), 128, 0, stream>>>(
        tabs, st, so, cb, main_buf, frames, flags, K0, n);
}
void launch_assemble_inc(const EncTables *tabs, const StreamDev *st, const StreamOut *so, ChunkBufs cb, int *done_lo,
                         const unsigned char *main_buf, const FrameRec *frames, unsigned char *out, long long *bytes_done,
                         int n, cudaStream_t stream) {
    k_assemble_inc<<<blocks_for((long long)n * kIncSlots * 32, 256), 256, 0, stream>>>(tabs, st, so, cb, done_lo, main_buf,
                                                                                       frames, out, n);
    k_advance_inc<<<blocks_for(n, 128), 128, 0, stream>>>(tabs, st, so, cb, done_lo, frames, bytes_done, n);
}
void launch_finish(const EncTables *tabs, const StreamDev *st, const StreamOut *so, const RateState *rs,
                   const FrameRec *frames, StreamResult *res, long long *out_off, const unsigned char *main_buf,
                   unsigned char *out, int max_frames, int n, cudaStream_t stream, cudaEvent_t before_assemble,
                   int frame_lo, long long out_base) {
    k_results<<<blocks_for(n, 64), 64, 0, stream>>>(tabs, st, so, rs, frames, res, n);
    k_out_offsets<<<1, 1024, 0, stream>>>(res, out_off, n);
    if (before_assemble) cudaEventRecord(before_assemble, stream);
    k_assemble<<<blocks_for((long long)n * max_frames * 32, 256), 256, 0, stream>>>(tabs, st, so, res, out_off, main_buf,
                                                                                     frames, out, max_frames, n, frame_lo,
                                                                                     out_base);
}
size_t sizeof_frame_rec() { return sizeof(FrameRec); }
size_t sizeof_pack_gc() { return sizeof(PackGc); }
}  // namespace hmp3
