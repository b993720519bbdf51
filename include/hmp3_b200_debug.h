/* Stage taps of the device pipeline, used only by the parity tests (tests/test_gpu_*.py). */
#ifndef HMP3_B200_DEBUG_H_
#define HMP3_B200_DEBUG_H_
#include "hmp3_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* Phase A (polyphase -> block switching -> hybrid MDCT -> psychoacoustic stage 1 / M-S measure) of one
 * stream on device `device`.  Any output pointer may be NULL.  Shapes (nch = channels of the stream):
 *   sbt   [ngran][nch][576]  frequency-inverted polyphase granules P[K] (entry ngran-1 is not filled)
 *   ginfo [ngran][4]         block_type, block_type_prev, short_flag_current, short_flag_next
 *   xr    [ngran][nch][576]  MDCT spectra handed to the rate loop
 *   raw   [ngran][nch][92]   psychoacoustic stage-1 record (PsyRaw)
 *   ms_raw[ngran]            M/S correlation measure without hysteresis
 *   att   [ngran][nch][9]    attack energies of P[K] (entry ngran-1 is not filled)
 * Replaces, for testing: sbt_L3 (sbt.c:293), attack_detectSBT_igr (detect.c:53), hybridLong/Short +
 * antialias (hwin.c:147-319), emap* (emap.c:61-121), spd_smr* (spdsmr.c:64-319). */
int hmp3_debug_analysis(const hmp3_control *ec, const int16_t *pcm, long long nsamples, int ngran, int device,
                        float *sbt, int *ginfo, float *xr, float *raw, int *ms_raw, int *att);
/* Load-balance diagnostics of the serial stage: the first call switches recording on; after a run, a call with a
 * buffer copies cycles[launch][stream] (SM clocks each stream's warp spent in that k_rate launch, up to 64 launches)
 * and returns the number of launches recorded. */
int hmp3_debug_rate_cycles(hmp3_batch *b, long long *cycles, int max_launches);
/* Tap of the serial stage (the device's rate loop, whichever kernel runs it): while set, every hmp3_batch_run copies
 * what the serial stage handed the packing pass for stream `stream` -- one record per granule-channel, record
 * 2 * K + ch for encode granule K -- into `records` (room for `cap_records`; granules beyond are dropped).  A record
 * is hmp3_debug_rate_tap_record_bytes() bytes: int16 ix[576] (quantised magnitudes in transmission order; only the
 * coded extent, 2 * big_values + 4 * count1 lines, is meaningful), uint32 sign[18], uint8 sf[64] (long: l[0..22];
 * short: s[w][i] at 23 + 13 w + i), int32 gr[27] (the granule's side information in the order of GR, pub/l3e.h:71-96,
 * followed by the packer's region sizes).  Compares with the reference's ix / scale factors / GR after
 * CBitAllo3::BitAllo and the frame driver (bitallo3.cpp:484-678, mp3enc.cpp:1492-2027).  records == NULL clears it. */
int hmp3_debug_set_rate_tap(hmp3_batch *b, int stream, void *records, long long cap_records);
int hmp3_debug_rate_tap_record_bytes(void);
/* The CMp3Enc mirror's sample-rate converter on its own (host code; Csrc cases 2-4, srcc.cpp / srccf.cpp): `ncalls`
 * calls, each producing 1152 frames at `target` Hz from float frames at `source` Hz (layout 0 mono, 1 two channels,
 * 2 two channels mixed down to one); used[c] = source frames call c consumed.  Returns the frames a call may look at
 * (what sr_convert_init's byte count is made of), 0 if the rates are refused, -1 / 0 for the 1:2 / 1:1 cases. */
int hmp3_debug_resample(int source, int target, int layout, const float *x, int ncalls, float *y, int *used);
/* Kernel timeline of the last run with timing on (hmp3_batch_set_timing): rows of (phase index as in
 * hmp3_batch_phase_ms, begin ms, end ms) since the run began; returns the number of rows. */
int hmp3_debug_timeline(const hmp3_batch *b, float *rows, int cap);
/* FP32 issue-rate microbenchmark on `device`: TFLOP/s of dependent FFMA chains and of the FMUL + FADD mix that code
 * compiled without contraction issues (the ceiling the exact-order kernels are quoted against). */
int hmp3_debug_fp32_peak(int device, float *ffma_tflops, float *nonfused_tflops);
#ifdef __cplusplus
}
#endif
#endif
