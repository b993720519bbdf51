// Device-side views of a batch and the host-callable kernel launchers.  The kernels live in two
// translation units (kernels_analysis.cu: Phase A, kernels_rate.cu: the serial stage) so that each can
// be compiled with the flags that suit it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hmp3 {

struct EncTables;
struct GranuleInfo;
struct PsyRaw;
struct SwitchState;
struct RateState;
struct FrameRec;
struct PackGc;
struct PrepGranule;
struct PsyState;
struct SigMask;

// One stream of the batch, device view.
struct StreamDev {
    int cfg;             // index into the tables array
    int nch;
    long long pcm_off;   // offset (int16 elements) of this stream's interleaved PCM in the batch buffer
    long long nsamples;  // per channel
    int ngran;           // encode granules to run, including the flush allowance
    int ngran_real;      // granules that belong to real encode calls (2 * calls)
    long long out_off;   // byte offset of this stream's output region
    long long out_cap;
    long long pcmf_off;  // DC-filtered float PCM of this stream (float elements), -1 = none (filter off)
    long long pcmf_len;  // samples per channel held there (covers the zero tail the encoder is flushed with)
    long long rawf_off;  // float input PCM of this stream (float elements, scaled to +-32768), -1 = the input is int16
    float tail;          // value of the samples past nsamples (0 unless the caller's flush bytes decode otherwise: 8-bit WAV)
};

// Chunk work buffers (device).  G = NG + 3 polyphase granules are kept per chunk: P[K0-3 .. K0+NG-1].
struct ChunkBufs {
    float *P;        // [n][NG+3][2][576]
    int *E;          // [n][NG+3][2][9]   attack energies of P
    GranuleInfo *gi; // [n][NG]
    float *xr;       // [n][NG][2][576]
    PsyRaw *raw;     // [n][NG][2]
    int *ms_raw;     // [n][NG]
    signed char *ms; // [n][NG]  M/S decision of the granule's frame (after hysteresis)
    SigMask *sm;     // [n][NG][2][36]  psychoacoustic sig/mask per granule-channel (stage 2)
    PrepGranule *prep; // [n][NG]  state-free part of the rate-loop prologue (long blocks)
    PackGc *pack;    // [n][NG][2]  granule-channel records for the packing pass
    int *fr0, *fr1;  // [n] frames recorded by each stream before / after this chunk's serial stage
    int *fd1;        // [n] frames complete (main-data slot filled) after this chunk's serial stage
    int NG;
};

// Per-stream placement of the serial stage's buffers (device view).
struct StreamOut {
    long long main_off;    // byte offset of this stream's main-data stream in the main buffer
    long long frames_off;  // first FrameRec of this stream
    int frames_cap;
};

struct StreamResult {
    long long out_bytes;
    int frames;           // complete frames (what was written to the output)
    int finished;
    int frames_recorded;  // frames the serial stage recorded (>= frames)
    int pad_;
};

// ---- launchers (all asynchronous on `stream`)
void launch_polyphase(const EncTables *tabs, const StreamDev *st, const int16_t *pcm, const float *pcmf, ChunkBufs cb,
                      int K0, int n, cudaStream_t stream);
// DC-blocking input filter (-S1): samples [lo, hi) of every stream that has it on; dc = carried state [n][2]
void launch_dc_filter(const EncTables *tabs, const StreamDev *st, const int16_t *pcm, float *pcmf, float *dc,
                      long long lo, long long hi, int n, cudaStream_t stream);
void launch_attack(const EncTables *tabs, const StreamDev *st, ChunkBufs cb, int K0, int n, cudaStream_t stream);
void launch_switch_scan(const EncTables *tabs, const StreamDev *st, SwitchState *sw, ChunkBufs cb, int K0, int n,
                        cudaStream_t stream);
void launch_hybrid(const EncTables *tabs, const StreamDev *st, ChunkBufs cb, int K0, int n, cudaStream_t stream);
void launch_psy_stage1(const EncTables *tabs, const StreamDev *st, ChunkBufs cb, int K0, int n, cudaStream_t stream);
// scans and prepare pass that follow psy stage 1 (carry: msmem [n], psy [n][2])
void launch_ms_scan(const EncTables *tabs, const StreamDev *st, int *msmem, ChunkBufs cb, int K0, int n, cudaStream_t stream);
void launch_psy_stage2(const EncTables *tabs, const StreamDev *st, PsyState *psy, ChunkBufs cb, int K0, int n,
                       cudaStream_t stream);
void launch_prepare(const EncTables *tabs, const StreamDev *st, ChunkBufs cb, int K0, int n, cudaStream_t stream);
int fp32_peak(int device, float *ffma_tflops, float *nonfused_tflops);
// contraction form of the polyphase on the tensor cores (kernels_polymm.cu; not bit-exact, off by default):
// parts = 3 (3xTF32) or 1 (plain TF32); wmat = build_polymm_matrix's output on the device
void launch_polyphase_mm(const EncTables *tabs, const StreamDev *st, const int16_t *pcm, const float *wmat, ChunkBufs cb,
                         int K0, int n, int parts, cudaStream_t stream);
void build_polymm_matrix(const EncTables *T, float *out);
size_t polymm_matrix_floats();
void launch_prepare_init(int *msmem, PsyState *psy, int n, cudaStream_t stream);
size_t sizeof_prep_granule();
size_t sizeof_psy_state();
void launch_rate_init(const EncTables *tabs, const StreamDev *st, RateState *rs, void *cold, int n, cudaStream_t stream);
void launch_rate(const EncTables *tabs, const StreamDev *st, const StreamOut *so, RateState *rs, ChunkBufs cb,
                 unsigned char *main_buf, FrameRec *frames, int K0, int n, cudaStream_t stream,
                 long long *cycles = nullptr);
// launch_rate in its phase-scheduled form (kernels_rate_ph.cu): a block owns a set of streams and its warps run the
// phase machine of rate_phased.h for them, phase by phase
void launch_rate_ph(const EncTables *tabs, const StreamDev *st, const StreamOut *so, RateState *rs, ChunkBufs cb,
                    FrameRec *frames, int K0, int n, cudaStream_t stream);
// the twin of launch_rate for the streams whose configuration selects CBitAllo1 (each kernel skips the other's streams)
void launch_rate_a1(const EncTables *tabs, const StreamDev *st, const StreamOut *so, RateState *rs, ChunkBufs cb,
                    unsigned char *main_buf, FrameRec *frames, int K0, int n, cudaStream_t stream);
// packing pass of the frames the serial stage recorded in this chunk; `flags[s]` is set if a frame's written
// bits ever differ from the accounted ones
// Incremental output (host entry with pinned buffers): after the packing pass of a chunk, the frames that became
// complete in it are assembled into the stream's fixed output region and the byte count reached is published.
void launch_assemble_inc(const EncTables *tabs, const StreamDev *st, const StreamOut *so, ChunkBufs cb, int *done_lo,
                         const unsigned char *main_buf, const FrameRec *frames, unsigned char *out, long long *bytes_done,
                         int n, cudaStream_t stream);
void launch_pack(const EncTables *tabs, const StreamDev *st, const StreamOut *so, ChunkBufs cb, unsigned char *main_buf,
                 FrameRec *frames, int *flags, int K0, int n, cudaStream_t stream);
// per-stream totals, compact output offsets (out_off[n] = total) and frame assembly
void launch_finish(const EncTables *tabs, const StreamDev *st, const StreamOut *so, const RateState *rs,
                   const FrameRec *frames, StreamResult *res, long long *out_off, const unsigned char *main_buf,
                   unsigned char *out, int max_frames, int n, cudaStream_t stream, cudaEvent_t before_assemble,
                   int frame_lo = 0, long long out_base = 0);
// Handle streaming (one stream): shift the serial stage's positional state by dK granules, keeping the frames that
// are not complete yet (moved to index 0) and their main data; res receives the state's counters afterwards.
void launch_handle_rebase(RateState *rs, FrameRec *frames, unsigned char *main_buf, int dK, StreamResult *res,
                          cudaStream_t stream);
size_t sizeof_rate_state();
size_t sizeof_rate_cold();
size_t sizeof_frame_rec();
size_t sizeof_pack_gc();

}  // namespace hmp3
