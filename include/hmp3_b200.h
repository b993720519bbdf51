/*
 * hmp3_b200 -- C ABI of the B200-native Helix-MP3 Layer III encode path.
 *
 * Plain C, plain pointers and sizes; no CUDA or torch types cross this boundary.  Every entry point
 * names the reference interface it replaces (paths relative to the maikmerten/hmp3 tree).
 *
 * The reference has no C ABI today (the old C prototypes are commented out, hmp3/src/pub/encapp.h:172-197);
 * its surface is the C++ class CMp3Enc (hmp3/src/pub/mp3enc.h:74-141) used by the CLI
 * (hmp3/src/test/tomp3.cpp:664, 818-820, 942-943, 1025-1026).  This header exports
 *   (1) a batch entry -- N independent streams in, N MP3 byte streams out -- which is what reaches the GPU;
 *   (2) opaque-handle mirrors of the CMp3Enc init/encode/info calls for tomp3-style callers;
 *   (3) error reporting.
 * There is NO CPU fallback: every encode entry fails with HMP3_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef HMP3_B200_H_
#define HMP3_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirror of E_CONTROL (hmp3/src/pub/encapp.h:42-72): same field order, same meaning, same defaults. */
typedef struct hmp3_control {
    int mode;          /* 0 stereo, 1 joint stereo, 2 dual, 3 mono                                  */
    int bitrate;       /* CBR per-channel kbit/s; -1 lets the encoder choose                        */
    int samprate;      /* 16000, 22050, 24000, 32000, 44100 or 48000                                */
    int nsbstereo;     /* -1 = encoder default                                                      */
    int filter_select; /* -1 default, 0 none, 1 DC-blocking filter                                  */
    int freq_limit;    /* 24000 = off                                                               */
    int nsb_limit;     /* -1 = off                                                                  */
    int layer;         /* 3                                                                         */
    int cr_bit;
    int original;
    int hf_flag;       /* MPEG-1 high-frequency coding: 1 = M/S granules, 3 = all granules          */
    int vbr_flag;      /* 1 = VBR, 0 = CBR                                                          */
    int vbr_mnr;       /* 0..150                                                                    */
    int vbr_br_limit;  /* per-channel VBR bitrate cap (160)                                         */
    int vbr_delta_mnr;
    int chan_add_f0;
    int chan_add_f1;
    int sparse_scale;
    int mnr_adjust[21];
    int cpu_select;
    int quick;
    int test1;
    int test2;
    int test3;
    int short_block_threshold;
} hmp3_control;

/* Mirror of IN_OUT (hmp3/src/pub/encapp.h:159-165). */
typedef struct hmp3_in_out {
    int in_bytes;
    int out_bytes;
} hmp3_in_out;

/* Mirror of INT_PAIR (hmp3/src/pub/encapp.h:167-171). */
typedef struct hmp3_int_pair {
    int a;
    int b;
} hmp3_int_pair;

/* Mirror of MPEG_HEAD (hmp3/src/pub/encapp.h:141-157). */
typedef struct hmp3_mpeg_head {
    int sync, id, option, prot, br_index, sr_index, pad, private_bit, mode, mode_ext, cr, original, emphasis;
} hmp3_mpeg_head;

enum {
    HMP3_OK = 0,
    HMP3_ERR_NO_DEVICE = -1,   /* no usable CUDA device: there is no CPU path                       */
    HMP3_ERR_BAD_CONTROL = -2, /* init rejected the control block (reference init returns 0)        */
    HMP3_ERR_UNSUPPORTED = -3, /* configuration outside the built path (intensity stereo, dual)     */
    HMP3_ERR_OUT_SPACE = -4,   /* caller's output buffer too small                                  */
    HMP3_ERR_CUDA = -5,
    HMP3_ERR_ARG = -6,
    HMP3_ERR_INTERNAL = -7     /* the packing pass disagreed with the bit accounting (a bug)      */
};

/* Sample formats of the batch entries.  Float PCM is on the +-32768 scale CMp3Enc::L3_audio_encode takes
 * (hmp3/src/pub/mp3enc.h:88-98); other WAV sample types are converted to it the way Csrc::sr_convert does
 * (hmp3/src/srcc.cpp:804-834: int8 -> (x-128)*256, int24 -> x/256, int32 -> x/65536, float -> x*32768). */
enum { HMP3_PCM_S16 = 0, HMP3_PCM_F32 = 1 };

/* Fill `ec` with the CLI defaults (hmp3/src/test/tomp3.cpp:357-387). */
void hmp3_control_defaults(hmp3_control *ec);

/* Apply one hmp3 command-line option ("-B64", "-V100", "-HF2", "-F19000", "-M0", ...) to `ec` with the
 * CLI's semantics (hmp3/src/test/tomp3.cpp:390-566; -B => CBR, no -B => VBR).  Returns 0, or -1 for an
 * option this build does not know. */
int hmp3_control_apply_option(hmp3_control *ec, const char *opt);

/* Last error text of the calling thread. */
const char *hmp3_get_last_error(void);

/* Number of usable CUDA devices (0 => every encode call fails). */
int hmp3_device_count(void);

/* ---------------------------------------------------------------------------------------------
 * (1) Batch entry: the GPU path.  Replaces one `hmp3 in.wav out.mp3 [opts]` process per clip
 * (ff_encode, hmp3/src/test/tomp3.cpp:640-1090, minus WAV parsing and the Xing/Info frame):
 * for every stream the output is the exact frame sequence CMp3Enc::MP3_audio_encode emits for
 * that PCM with its end-of-file protocol -- 4 x 1153 zero sample frames appended, one encode call while 1153 frames
 * are buffered, then the tail flush (tomp3.cpp:908-942, 1015-1036).
 * --------------------------------------------------------------------------------------------- */
typedef struct hmp3_stream_desc {
    const hmp3_control *control; /* per-stream control (streams with equal controls share tables) */
    const void *pcm;             /* interleaved PCM in HOST memory (pinned is faster): int16, or float */
                                 /* on the +-32768 scale when pcm_format == HMP3_PCM_F32               */
    int64_t num_samples;         /* samples per channel                                           */
    uint8_t *out;                /* HOST buffer for the MP3 frames                                */
    int64_t out_capacity;        /* bytes available at `out` (see hmp3_batch_out_bound)           */
    int64_t out_bytes;           /* [out] bytes written                                           */
    int32_t out_frames;          /* [out] frames written                                          */
    int32_t status;              /* [out] HMP3_OK or an error for this stream                     */
    int32_t pcm_format;          /* HMP3_PCM_S16 (0, default) or HMP3_PCM_F32                     */
} hmp3_stream_desc;

/* Upper bound of the output size for a stream. */
int64_t hmp3_batch_out_bound(const hmp3_control *control, int64_t num_samples);

/* Encode `n` independent streams on CUDA device `device`.  Host buffers in, host buffers out
 * (H2D/D2H copies are inside).  Returns HMP3_OK if the batch ran; per-stream results in status. */
int hmp3_encode_batch(hmp3_stream_desc *streams, int n, int device);

/* Reusable batch plan: device buffers sized once for a batch shape (controls + clip lengths), then
 * any number of encodes of that shape.  hmp3_encode_batch is create + encode_host + destroy. */
typedef struct hmp3_batch hmp3_batch; /* opaque */
hmp3_batch *hmp3_batch_create(const hmp3_control *controls, const int64_t *num_samples, int n, int device);
/* ... with a sample format per stream (NULL = all int16) */
hmp3_batch *hmp3_batch_create_ex(const hmp3_control *controls, const int64_t *num_samples, const int32_t *pcm_formats,
                                 int n, int device);
void hmp3_batch_destroy(hmp3_batch *b);
/* Host buffers in, host buffers out through a plan: uploads every stream's PCM (H2D), runs all kernels,
 * copies the frames back (D2H) into out[i].  Arrays have n entries; out_bytes/out_frames/status may be NULL. */
int hmp3_batch_encode_host(hmp3_batch *b, const void *const *pcm, uint8_t *const *out, const int64_t *out_cap,
                           int64_t *out_bytes, int32_t *out_frames, int32_t *status);
/* Device-resident legs (the benchmark's kernel-only measurement and custom feeders): */
int16_t *hmp3_batch_device_pcm(hmp3_batch *b);          /* all streams' interleaved int16 PCM           */
int64_t hmp3_batch_pcm_offset(const hmp3_batch *b, int i); /* int16 element offset of stream i in it    */
uint8_t *hmp3_batch_device_out(hmp3_batch *b);          /* compact output: stream i at out_offsets[i]   */
int64_t hmp3_batch_out_capacity(const hmp3_batch *b);
int hmp3_batch_upload(hmp3_batch *b, int i, const int16_t *pcm, int64_t num_samples); /* async H2D      */
int hmp3_batch_upload_f32(hmp3_batch *b, int i, const float *pcm, int64_t num_samples);
/* Value of the samples after the end of float stream `i` (default 0).  The reference CLI flushes the encoder with
 * zero BYTES (hmp3/src/test/tomp3.cpp:925-930, 1017), which an 8-bit source decodes to (0-128)*256 = -32768: a
 * caller that wants the reference's file for 8-bit WAV input sets -32768 here.  Call before the run. */
int hmp3_batch_set_tail(hmp3_batch *b, int i, float value);
int hmp3_batch_wait_uploads(hmp3_batch *b);             /* block until queued uploads have landed       */
/* run all kernels on the plan's stream over whatever PCM is resident.  Synchronous unless `async` != 0
 * (then hmp3_batch_sync must be called before reading results). */
int hmp3_batch_run(hmp3_batch *b, int async);
int hmp3_batch_sync(hmp3_batch *b);
/* per-stream results of the last completed run (host arrays of n; any may be NULL) */
int hmp3_batch_results(hmp3_batch *b, int64_t *out_bytes, int32_t *out_frames, int64_t *out_offsets,
                       int32_t *status);
int hmp3_batch_download(hmp3_batch *b, int i, uint8_t *out, int64_t cap);
int hmp3_batch_download_all(hmp3_batch *b, uint8_t *out, int64_t cap, int64_t *total);
/* number of kernel launches issued by the last hmp3_batch_run */
int hmp3_batch_last_launches(const hmp3_batch *b);
/* encode granules each stream advances per launch of the serial-stage kernel (the chunk length) */
int hmp3_batch_chunk_granules(const hmp3_batch *b);
/* device time (CUDA events on the plan's stream, first kernel to last kernel) of the last completed run, ms */
float hmp3_batch_last_run_ms(const hmp3_batch *b);
/* per-kernel device time (CUDA events on the plan's stream) of the last synchronous run made after
 * hmp3_batch_set_timing(b, 1): fills up to `cap` entries, returns the count; names[i] are static. */
int hmp3_batch_set_timing(hmp3_batch *b, int on);
/* Diagnostics: on != 0 puts every kernel of the following runs on ONE stream (no overlap between Phase A, the serial
 * stage and the packing pass), so that the per-kernel times of hmp3_batch_phase_ms are uncontended; 0 restores the
 * three-stream pipeline.  The output is the same either way. */
int hmp3_batch_set_serialize(hmp3_batch *b, int on);
int hmp3_batch_phase_ms(const hmp3_batch *b, const char **names, float *ms, int *launches, int cap);

/* ---------------------------------------------------------------------------------------------
 * (2) CMp3Enc mirrors (hmp3/src/pub/mp3enc.h:74-141).  A handle is one stream; calls buffer PCM and
 * run the same device pipeline as the batch entry.
 * --------------------------------------------------------------------------------------------- */
typedef struct hmp3_encoder hmp3_encoder;
hmp3_encoder *hmp3_encoder_new(int device);
void hmp3_encoder_delete(hmp3_encoder *e);
/* A handle encodes a stream of any length in bounded memory: its device buffers hold a window of `seconds` of audio
 * (default 20 s, about 0.5 MB per second of window for 44.1 kHz stereo) and are recycled when the window is full.
 * Call before init. */
int hmp3_encoder_set_capacity_seconds(hmp3_encoder *e, int seconds);

/* CMp3Enc::MP3_audio_encode_init (hmp3/src/mp3enc.cpp:2655-2808): returns, like the reference, the bytes the
 * caller must have buffered for a call (what Csrc::sr_convert_init returns; a call consumes what hmp3_in_out.in_bytes
 * reports), 0 = failure.  ec->samprate is the SOURCE rate (4000 < rate <= 48000; the converter takes 8000 and up);
 * mpeg_select picks the encode rate as the reference does: 0 = the nearest MPEG rate (twice the source below 16 kHz),
 * 1 = an MPEG-1 rate, 2 = an MPEG-2 rate, any other value = that rate, which must be one of the six.
 * 8/16/24/32-bit integer and 32-bit float PCM; mono_convert down-mixes a two-channel source.  The sample-rate
 * converter is Csrc's (hmp3/src/srcc.cpp, srccf.cpp), every case: none, 1:2 up, m:n up (linear interpolation),
 * down through a polyphase FIR, down in two stages.  Like the reference's, an up-converting call may look one or two
 * sample frames past the bytes this function returns. */
int hmp3_MP3_audio_encode_init(hmp3_encoder *e, const hmp3_control *ec, int source_bits, int source_is_float,
                               int mpeg_select, int mono_convert);
/* CMp3Enc::MP3_audio_encode (hmp3/src/mp3enc.cpp:2812-2828). */
hmp3_in_out hmp3_MP3_audio_encode(hmp3_encoder *e, const unsigned char *pcm, unsigned char *bs_out);
/* CMp3Enc::L3_audio_encode_init / L3_audio_encode, float PCM scaled to +-32768
 * (hmp3/src/mp3enc.cpp:220-870, 2031-2047). */
int hmp3_L3_audio_encode_init(hmp3_encoder *e, const hmp3_control *ec);
hmp3_in_out hmp3_L3_audio_encode(hmp3_encoder *e, const float *pcm, unsigned char *bs_out);
/* CMp3Enc::L3_audio_encode_Packet / MP3_audio_encode_Packet (hmp3/src/mp3enc.cpp:2831-3440; pub/mp3enc.h:100-125):
 * the same encode call, which additionally returns the frame(s) produced by THIS call as self-contained
 * "reformatted" packets for streaming: header | side info with main_data_begin = 0 | the frame's own main data (no
 * bit reservoir).  MPEG-1 yields one packet per call, MPEG-2 two; nbytes_out[k] is the size of packet k (0 = none),
 * packet k+1 follows packet k in `packet`.  bs_out may be NULL (no standard bitstream wanted: out_bytes is then 0,
 * as in the reference), packet may be NULL (plain encode). */
hmp3_in_out hmp3_L3_audio_encode_Packet(hmp3_encoder *e, const float *pcm, unsigned char *bs_out, unsigned char *packet,
                                        int nbytes_out[2]);
hmp3_in_out hmp3_MP3_audio_encode_Packet(hmp3_encoder *e, const unsigned char *pcm, unsigned char *bs_out,
                                         unsigned char *packet, int nbytes_out[2]);
/* Info getters (hmp3/src/mp3enc.cpp:3444-3527). */
void hmp3_L3_audio_encode_info_ec(hmp3_encoder *e, hmp3_control *ec);
void hmp3_L3_audio_encode_info_head(hmp3_encoder *e, hmp3_mpeg_head *head);
void hmp3_L3_audio_encode_info_string(hmp3_encoder *e, char *s);
unsigned int hmp3_L3_audio_encode_get_frames(hmp3_encoder *e);
int hmp3_L3_audio_encode_get_bitrate(hmp3_encoder *e);
float hmp3_L3_audio_encode_get_bitrate_float(hmp3_encoder *e);
/* CMp3Enc::L3_audio_encode_get_bitrate2_float (mp3enc.cpp:3466-3480; pub/mp3enc.h:132): bitrate of the recent
 * calls, from the running average the encoder keeps (used by the CLI's progress line, test/tomp3.cpp:633). */
float hmp3_L3_audio_encode_get_bitrate2_float(hmp3_encoder *e);
/* CMp3Enc::L3_audio_encode_get_frames_bytes (mp3enc.cpp:3513-3521; pub/mp3enc.h:137): frames and bytes emitted
 * so far (the CLI's seek table, test/tomp3.cpp:981). */
hmp3_int_pair hmp3_L3_audio_encode_get_frames_bytes(hmp3_encoder *e);

/* ---------------------------------------------------------------------------------------------
 * (3) Host post-pass of the CLI: the Xing/Info frame the reference writes in front of
 * the audio frames.  Pure host code.  Replaces XingHeader / XingHeaderTOC / XingHeaderUpdateInfo
 * (hmp3/src/xhead.c:255-709) as driven by ff_encode (hmp3/src/test/tomp3.cpp:871-896, 962-984, 1055-1072).
 * --------------------------------------------------------------------------------------------- */
/* The effective control block (CMp3Enc::L3_audio_encode_info_ec, mp3enc.cpp:3490) and header of a control. */
int hmp3_effective_control(const hmp3_control *ec, hmp3_control *effective, hmp3_mpeg_head *head);
/* Size of the Xing/Info frame for this configuration (0 = the CLI writes none / it does not fit). */
int hmp3_info_frame_size(const hmp3_control *effective, int head_mode, int xing_flag, int channels);
/* Builds the final Xing/Info frame.  effective/head_mode: from hmp3_effective_control; xing_flag: the CLI's -X
 * value (default 67 = 3 | INFOTAG); source_rate/channels/nsamples: the input PCM (samples per channel);
 * audio/audio_bytes/frames: every audio frame that follows (MusicCRC, counts); frames_after_call /
 * bytes_after_call [ncalls]: the encoder's cumulative output after each encode call of the CLI's main loop
 * (seek table; see hmp3_batch_call_log).  audio == NULL builds the placeholder the CLI writes first.
 * Returns the frame size, 0 on failure. */
int hmp3_info_frame(const hmp3_control *effective, int head_mode, int xing_flag, int source_rate, int channels,
                    int64_t nsamples, const uint8_t *audio, int64_t audio_bytes, uint32_t frames,
                    const int32_t *frames_after_call, const int64_t *bytes_after_call, int ncalls, uint8_t *buf,
                    int cap);
/* Per encode call of stream i in the last completed run: frames and bytes emitted so far (what
 * CMp3Enc::L3_audio_encode_get_frames_bytes returns after that call).  Fills up to `cap` entries, returns the
 * number of calls the stream made including the tail flush (the CLI's main loop is the first
 * (num_samples + 3 * 1153 + 1152) / 1152 of them: the CLI keeps calling while 1153 sample frames are buffered). */
int hmp3_batch_call_log(hmp3_batch *b, int i, int32_t *frames_after_call, int64_t *bytes_after_call, int cap);

/* ---------------------------------------------------------------------------------------------
 * Resolved configuration, for boundary tests (what L3_audio_encode_init computes,
 * hmp3/src/mp3enc.cpp:289-870).  Pure host logic, needs no device.
 * --------------------------------------------------------------------------------------------- */
typedef struct hmp3_resolved {
    int nchan, h_id, sr_index, nband, band_limit, nsb, nsb_limit, nsb_limit_ms0, nsb_limit_ms1, ave_target_bits;
    int framebytes, main_framebytes, side_bytes, remainder, divisor, ms_flag, is_flag, frame_driver,
        granule_driver, ivbr_min, ivbr_max, vbr_pool_target, short_block_threshold, h_mode, br_index,
        totbitrate, samprate, band_limit_stereo, sf_bit_max, nsf_stereo;
    int head[4];
    int hf_flag, filter_select, bytes_in;
} hmp3_resolved;
int hmp3_resolve_control(const hmp3_control *ec, hmp3_resolved *out);

#ifdef __cplusplus
}
#endif
#endif /* HMP3_B200_H_ */
