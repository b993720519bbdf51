/*
 * CMp3EncGpu -- the reference's encoder class (hmp3/src/pub/mp3enc.h:74-141) over the hmp3_b200 C ABI.
 *
 * Header-only adapter for callers that hold a CMp3Enc object (the reference CLI, hmp3/src/test/tomp3.cpp:664).
 * Include the reference's own "encapp.h" first: E_CONTROL, IN_OUT and MPEG_HEAD are the reference's types, and the
 * hmp3_* structs of hmp3_b200.h mirror them field for field (static_asserts below).  Same calls, same argument
 * meaning, same return conventions (0 = init failure, IN_OUT per encode call).  oracle/Makefile builds the
 * reference's unmodified CLI against this class (oracle/_ref/tomp3_gpu) and tests/test_gpu_shim.py checks that it
 * writes the same file as the reference's own binary.
 */
#ifndef HMP3_B200_CMP3ENC_GPU_H_
#define HMP3_B200_CMP3ENC_GPU_H_

#include "hmp3_b200.h"

#ifndef HMP3_SHIM_HAVE_INT_PAIR
typedef struct { int a; int b; } INT_PAIR; /* hmp3/src/pub/mp3enc.h:66-71 */
#endif

class CMp3EncGpu {
    hmp3_encoder *e;
    CMp3EncGpu(const CMp3EncGpu &);
    CMp3EncGpu &operator=(const CMp3EncGpu &);
    static_assert(sizeof(E_CONTROL) == sizeof(hmp3_control), "hmp3_control must mirror E_CONTROL");
    static_assert(sizeof(MPEG_HEAD) == sizeof(hmp3_mpeg_head), "hmp3_mpeg_head must mirror MPEG_HEAD");
    static_assert(sizeof(IN_OUT) == sizeof(hmp3_in_out), "hmp3_in_out must mirror IN_OUT");
    static IN_OUT io(hmp3_in_out x) { IN_OUT r; r.in_bytes = x.in_bytes; r.out_bytes = x.out_bytes; return r; }

  public:
    explicit CMp3EncGpu(int device = 0) : e(hmp3_encoder_new(device)) {}
    ~CMp3EncGpu() { hmp3_encoder_delete(e); }

    int L3_audio_encode_init(E_CONTROL *ec) { return hmp3_L3_audio_encode_init(e, (const hmp3_control *)ec); }
    IN_OUT L3_audio_encode(float *pcm, unsigned char *bs_out) { return io(hmp3_L3_audio_encode(e, pcm, bs_out)); }
    IN_OUT L3_audio_encode_Packet(float *pcm, unsigned char *bs_out, unsigned char *packet, int nbytes_out[2]) {
        return io(hmp3_L3_audio_encode_Packet(e, pcm, bs_out, packet, nbytes_out));
    }
    int MP3_audio_encode_init(E_CONTROL *ec, int source_bits, int source_is_float, int mpeg_select, int mono_convert) {
        return hmp3_MP3_audio_encode_init(e, (const hmp3_control *)ec, source_bits, source_is_float, mpeg_select,
                                          mono_convert);
    }
    IN_OUT MP3_audio_encode(unsigned char *pcm, unsigned char *bs_out) { return io(hmp3_MP3_audio_encode(e, pcm, bs_out)); }
    IN_OUT MP3_audio_encode_Packet(unsigned char *pcm, unsigned char *bs_out, unsigned char *packet, int nbytes_out[2]) {
        return io(hmp3_MP3_audio_encode_Packet(e, pcm, bs_out, packet, nbytes_out));
    }
    int L3_audio_encode_get_bitrate() { return hmp3_L3_audio_encode_get_bitrate(e); }
    float L3_audio_encode_get_bitrate_float() { return hmp3_L3_audio_encode_get_bitrate_float(e); }
    float L3_audio_encode_get_bitrate2_float() { return hmp3_L3_audio_encode_get_bitrate2_float(e); }
    unsigned int L3_audio_encode_get_frames() { return hmp3_L3_audio_encode_get_frames(e); }
    void L3_audio_encode_info_ec(E_CONTROL *ec) { hmp3_L3_audio_encode_info_ec(e, (hmp3_control *)ec); }
    void L3_audio_encode_info_head(MPEG_HEAD *head) { hmp3_L3_audio_encode_info_head(e, (hmp3_mpeg_head *)head); }
    void L3_audio_encode_info_string(char *s) { hmp3_L3_audio_encode_info_string(e, s); }
    INT_PAIR L3_audio_encode_get_frames_bytes() {
        hmp3_int_pair x = hmp3_L3_audio_encode_get_frames_bytes(e);
        INT_PAIR r; r.a = x.a; r.b = x.b; return r;
    }
    void out_stats() {}
};
#endif
