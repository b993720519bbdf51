// hmp3b200 -- command-line front end with the reference CLI's calling convention
//     hmp3b200 <input.wav> <output.mp3> [options]          (hmp3/src/test/tomp3.cpp:317-566)
// over the B200 path: options are parsed with hmp3_control_apply_option (same letters, same meaning: -B<n> CBR
// per-channel kbit/s, -V<n> VBR, -HF<n>, -F<hz>, -M<mode>, -X<flag> ...), the WAV is encoded through the C ABI and
// the output file is the Xing/Info frame followed by the audio frames -- byte-identical to what `hmp3` writes.
// Extension: `-@ <list>` encodes many files in ONE batch on the GPU (each line of <list>: input<TAB or space>output).
// Input: 8/16/24/32-bit integer or 32-bit float PCM WAV at any rate from 8 to 48 kHz; -A<n> picks the encode rate as
// the reference does.  A file at its encode rate or at half of it joins the batch (1:2 up-conversion is done here);
// any other pair of rates goes through the handle, whose MP3_audio_encode converts call by call (Csrc cases 2-4).  "-" as
// input reads the WAV from stdin and, like -IL, ignores the header's data length (tomp3.cpp:721-745); "-" as output
// writes to stdout, where -- as with the reference, which cannot re-read its own stdout -- the Xing/Info frame stays
// the placeholder written before encoding (tomp3.cpp:1055-1072).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/hmp3_b200.h"

namespace {

// encode calls of the reference CLI's main loop for n sample frames (see calls_for in pipeline.cu)
long long calls_main(long long n) { return (n + 3 * 1153 + 1152) / 1152; }

struct Wav {
    int channels = 0, rate = 0, bits = 0, type = 0;
    std::vector<int16_t> pcm;  // interleaved, 16-bit input
    std::vector<float> pcmf;   // interleaved, every other input type converted the way Csrc::sr_convert does
    int enc_channels = 0;      // channels handed to the encoder (1 after a -M3 down-mix of a stereo file)
    bool use_float = false;
    size_t audio_bytes = 0;    // bytes of the data chunk actually read (the Info tag counts whole sample frames of it)
    size_t total() const { return use_float ? pcmf.size() : pcm.size(); }
};

bool read_wav(const char *path, Wav *w, std::string *err, bool ignore_length) {
    const bool from_pipe = !strcmp(path, "-");
    if (from_pipe) ignore_length = true;
    FILE *f = from_pipe ? stdin : fopen(path, "rb");
    if (!f) {
        *err = "CANNOT_OPEN_INPUT_FILE";
        return false;
    }
    struct Closer {
        FILE *f;
        bool keep;
        ~Closer() { if (!keep) fclose(f); }
    } closer{f, from_pipe};
    unsigned char h[12];
    if (fread(h, 1, 12, f) != 12 || memcmp(h, "RIFF", 4) || memcmp(h + 8, "WAVE", 4)) {
        *err = "UNSUPPORTED PCM FILE TYPE";
        return false;
    }
    bool have_fmt = false;
    for (;;) {
        unsigned char c[8];
        if (fread(c, 1, 8, f) != 8) break;
        const uint32_t n = c[4] | (c[5] << 8) | (c[6] << 16) | ((uint32_t)c[7] << 24);
        if (!memcmp(c, "fmt ", 4)) {
            std::vector<unsigned char> b(n);
            if (fread(b.data(), 1, n, f) != n || n < 16) break;
            w->type = b[0] | (b[1] << 8);
            w->channels = b[2] | (b[3] << 8);
            w->rate = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
            w->bits = b[14] | (b[15] << 8);
            if (w->type == 0xFFFE && n >= 26) w->type = b[24] | (b[25] << 8);  // WAVE_FORMAT_EXTENSIBLE sub-format
            have_fmt = true;
            if (n & 1) fgetc(f);
        } else if (!memcmp(c, "data", 4)) {
            if (!have_fmt) break;
            const bool ok_type = (w->type == 1 && (w->bits == 8 || w->bits == 16 || w->bits == 24 || w->bits == 32)) ||
                                 (w->type == 3 && w->bits == 32);
            if (!ok_type || w->channels < 1 || w->channels > 2) {
                        *err = "UNSUPPORTED PCM FILE TYPE\n Only 8, 16, 24 and 32 bit linear PCM or 32-bit floating point supported.";
                return false;
            }
            std::vector<unsigned char> raw;
            size_t got = 0;
            if (ignore_length) {  // everything up to end of file is audio
                for (;;) {
                    raw.resize(got + (1u << 20));
                    const size_t r = fread(raw.data() + got, 1, 1u << 20, f);
                    got += r;
                    if (r < (1u << 20)) break;
                }
            } else {
                raw.resize(n == 0xFFFFFFFFu ? 0 : n);
                got = raw.empty() ? 0 : fread(raw.data(), 1, raw.size(), f);
            }
            const size_t bps = (size_t)w->bits / 8, frame = bps * (size_t)w->channels;
            w->audio_bytes = got;
            // The reference feeds BYTES: a trailing partial sample frame (possible with -IL / stdin) is followed by its
            // zero padding and encoded as one more sample per channel.  Keep it unless that would add an encode call
            // the reference does not make (its call count comes from the byte count, rounded down).
            if (got % frame) {
                const size_t whole = got / frame;
                if (calls_main((long long)whole + 1) == calls_main((long long)whole)) {
                    raw.resize((whole + 1) * frame + frame);
                    memset(raw.data() + got, 0, (whole + 1) * frame - got);
                    got = (whole + 1) * frame;
                } else got = whole * frame;
            }
            const size_t ns = got / bps;
            const unsigned char *p = raw.data();
            if (w->type == 1 && w->bits == 16) {
                w->pcm.resize(ns);
                for (size_t i = 0; i < ns; i++) w->pcm[i] = (int16_t)(p[2 * i] | (p[2 * i + 1] << 8));
            } else {  // sample conversion of Csrc::sr_convert (hmp3/src/srcc.cpp:804-834)
                w->pcmf.resize(ns);
                for (size_t i = 0; i < ns; i++) {
                    if (w->type == 3) {
                        float v;
                        memcpy(&v, p + 4 * i, 4);
                        w->pcmf[i] = (float)(v) * 32768.0f;
                    } else if (w->bits == 32) {
                        int v;
                        memcpy(&v, p + 4 * i, 4);
                        w->pcmf[i] = (float)(v / 65536.0f);
                    } else if (w->bits == 24) {
                        const int v = (int)(((unsigned)p[3 * i + 2] << 24) | ((unsigned)p[3 * i + 1] << 16) | ((unsigned)p[3 * i] << 8)) >> 8;
                        w->pcmf[i] = (float)((float)v / 256.0f);
                    } else {
                        w->pcmf[i] = (((float)p[i]) - 128.0f) * (256.0f);
                    }
                }
            }
                return true;
        } else {
            bool ok = true;  // skip the chunk (by reading: stdin cannot seek)
            for (uint32_t left = n + (n & 1); left && ok;) {
                unsigned char skip[4096];
                const size_t want = left < sizeof(skip) ? left : sizeof(skip);
                ok = fread(skip, 1, want, f) == want;
                left -= (uint32_t)want;
            }
            if (!ok) break;
        }
    }
    *err = "UNSUPPORTED PCM FILE TYPE";
    return false;
}

struct Job {
    std::string in, out;
    Wav wav;
    hmp3_control ec;
    int enc_rate = 0;  // sample rate handed to the encoder (twice the file's for 8 / 11.025 / 12 kHz input)
    bool ok = false;
    bool via_handle = false;  // needs the general sample-rate converter: encoded through the CMp3Enc mirror, call by call
};

}  // namespace

// Encode sample rate for a source rate and the -A value, as CMp3Enc::MP3_audio_encode_init picks it
// (mp3enc.cpp:2628-2651 find_nearest over {22050, 24000, 16000, 44100, 48000, 32000}, :2696-2748); 0 = init fails.
static int nearest_rate(const int *table, int n, int x) {
    int best = table[0], d0 = abs(table[0] - x);
    for (int i = 0; i < n; i++) {
        const int d = abs(table[i] - x);
        if (d < d0) {
            d0 = d;
            best = table[i];
        }
    }
    return best;
}
static int target_rate(int source, int mpeg_select) {
    static const int rates[6] = {22050, 24000, 16000, 44100, 48000, 32000};
    if (source < 4000 || source > 48000) return 0;
    int t;
    switch (mpeg_select) {
    case 0:
        if (source < 16000 && (t = nearest_rate(rates, 3, 2 * source)) == 2 * source) return t;
        return nearest_rate(rates, 6, source);
    case 1:
        if (source < 16000 && (t = nearest_rate(rates + 3, 3, 4 * source)) == 4 * source) return t;
        if (source < 32000 && (t = nearest_rate(rates + 3, 3, 2 * source)) == 2 * source) return t;
        return nearest_rate(rates + 3, 3, source);
    case 2:
        if (source < 16000 && (t = nearest_rate(rates, 3, 2 * source)) == 2 * source) return t;
        if (source > 24000 && 2 * (t = nearest_rate(rates, 3, source / 2)) == source) return t;
        return nearest_rate(rates, 3, source);
    default:
        t = nearest_rate(rates, 6, mpeg_select);
        return t == mpeg_select ? t : 0;
    }
}


// Writes one encoded stream the way ff_encode leaves the file: the Xing/Info frame (tomp3.cpp:871-896, 1055-1072), then
// the audio frames.  frames_after_call / bytes_after_call: the main loop's calls (the seek table), nsamples_enc: the
// stream length at the encode rate (for the closing report only).
static bool write_stream(const Job &j, const hmp3_control &eff, const hmp3_mpeg_head &head, int xing, const uint8_t *audio,
                         int64_t nb, int32_t nf, const int32_t *fa, const int64_t *ba, int nc, int64_t nsamples_enc) {
    uint8_t tag[2048];
    int tag_bytes = 0;
    const bool to_pipe = (j.out == "-");
    if (xing && to_pipe) {
        tag_bytes = hmp3_info_frame(&eff, head.mode, xing, j.wav.rate, j.wav.channels, nsamples_enc, nullptr, 0, 0, nullptr,
                                    nullptr, 0, tag, (int)sizeof(tag));
    } else if (xing) {
        const int64_t samples_audio = (int64_t)(j.wav.audio_bytes / ((size_t)j.wav.channels * (j.wav.bits / 8)));
        tag_bytes = hmp3_info_frame(&eff, head.mode, xing, j.wav.rate, j.wav.channels, samples_audio, audio, nb, (uint32_t)nf,
                                    fa, ba, nc, tag, (int)sizeof(tag));
    }
    FILE *f = to_pipe ? stdout : fopen(j.out.c_str(), "wb");
    if (!f) {
        fprintf(stderr, "\n CANNOT CREATE OUTPUT FILE %s\n", j.out.c_str());
        return false;
    }
    if (tag_bytes) fwrite(tag, 1, (size_t)tag_bytes, f);
    fwrite(audio, 1, (size_t)nb, f);
    if (to_pipe) fflush(f);
    else fclose(f);
    const double secs = (double)nsamples_enc / j.enc_rate;
    fprintf(stderr, "\n %s: %d frames, %lld bytes, %.2f kbps", j.out.c_str(), nf, (long long)(nb + tag_bytes),
            secs > 0 ? 8e-3 * (double)nb / secs : 0.0);
    return true;
}

// A file whose rate has to be converted by more than 1:2 (Csrc cases 2-4): the reference CLI's own loop over the
// CMp3Enc mirror (ff_encode, tomp3.cpp:908-1036).  At the end of the data 4 x bytes_in_init zero bytes are appended
// and calls go on while bytes_in_init bytes are buffered; then zero input is fed until the frame count has caught up
// with the calls made.
static bool encode_with_handle(Job &j, int mpeg_select, bool mono_convert, int device, int xing) {
    hmp3_encoder *e = hmp3_encoder_new(device);
    if (!e) {
        fprintf(stderr, "\n ENCODER INIT FAIL: %s\n", hmp3_get_last_error());
        return false;
    }
    struct Closer {
        hmp3_encoder *e;
        ~Closer() { hmp3_encoder_delete(e); }
    } closer{e};
    const int ch = j.wav.channels;
    const bool f32 = j.wav.use_float;
    const int fb = ch * (f32 ? 4 : 2);
    hmp3_control ec = j.ec;
    ec.samprate = j.wav.rate;
    // 16-bit samples go in as read.  Every other type was converted by the reader to float on the +-32768 scale, the
    // way Csrc::sr_convert converts it; handed over as float on the +-1 scale it reaches the converter with those values.
    const int bytes_in = hmp3_MP3_audio_encode_init(e, &ec, f32 ? 32 : 16, f32 ? 1 : 0, mpeg_select, mono_convert ? 1 : 0);
    if (!bytes_in) {
        fprintf(stderr, "\n ENCODER INIT FAIL: %s\n", hmp3_get_last_error());
        return false;
    }
    const size_t minfr = (size_t)bytes_in / fb, nfr = j.wav.total() / ch;
    // The CLI appends its 4 x bytes_in_init zero bytes only if they fit what it takes for the free part of its PCM
    // buffer (tomp3.cpp:268, 802, 925; pcmhpm.c:450): with wide samples and a converter that buffers a lot (two-stage
    // down-conversion of 24/32-bit stereo) they do not, and the tail of the file shorter than one call is dropped.
    const size_t src_bii = minfr * (size_t)ch * (size_t)(j.wav.bits / 8);
    const size_t psf = (size_t)ch * (size_t)((j.wav.bits * 7) / 8);
    const bool padded = 4 * src_bii < (((size_t)(2u * 128u * 4u * 2304u) / psf) & ~(size_t)1);
    const size_t tot = nfr + (padded ? 4 * minfr : 0);  // the data, then the CLI's padding
    // ... and what a call may look at past that: an up-converting call reads a frame or two more than bytes_in_init, in
    // the CLI's buffer leftovers of earlier reads when the last call starts with exactly that much buffered; zero here
    const size_t room = tot + minfr + 8;
    std::vector<uint8_t> src(room * fb);
    if (f32) {
        float *d = (float *)src.data();
        const float pad = j.wav.bits == 8 ? -1.0f : 0.0f;  // what the zero bytes decode to
        for (size_t i = 0; i < room * ch; i++) d[i] = i < nfr * ch ? j.wav.pcmf[i] * (1.0f / 32768.0f) : pad;
    } else {
        memcpy(src.data(), j.wav.pcm.data(), nfr * fb);
    }
    std::vector<uint8_t> out, bs(65536), zeros((4 * minfr + 8) * fb, 0);
    if (f32 && j.wav.bits == 8)
        for (size_t i = 0; i < zeros.size() / 4; i++) ((float *)zeros.data())[i] = -1.0f;
    std::vector<int32_t> fa;
    std::vector<int64_t> ba;
    size_t pos = 0;
    unsigned calls = 0;
    while (tot - pos >= minfr) {
        const hmp3_in_out x = hmp3_MP3_audio_encode(e, src.data() + pos * fb, bs.data());
        if (x.in_bytes <= 0) {
            fprintf(stderr, "\n ENCODE FAILED: %s\n", hmp3_get_last_error());
            return false;
        }
        pos += (size_t)x.in_bytes / fb;
        out.insert(out.end(), bs.begin(), bs.begin() + x.out_bytes);
        const hmp3_int_pair p = hmp3_L3_audio_encode_get_frames_bytes(e);
        fa.push_back(p.a);
        ba.push_back(p.b);
        calls++;
    }
    hmp3_control eff;
    hmp3_mpeg_head head;
    hmp3_L3_audio_encode_info_ec(e, &eff);
    hmp3_L3_audio_encode_info_head(e, &head);
    const unsigned expected = eff.samprate < 32000 ? 2 * calls : calls;
    for (int guard = 0; hmp3_L3_audio_encode_get_frames(e) < expected && guard < 64; guard++) {
        const hmp3_in_out x = hmp3_MP3_audio_encode(e, zeros.data(), bs.data());
        out.insert(out.end(), bs.begin(), bs.begin() + x.out_bytes);
    }
    const int32_t nf = (int32_t)hmp3_L3_audio_encode_get_frames(e);
    const int64_t ns_enc = (int64_t)((double)nfr * j.enc_rate / j.wav.rate);
    return write_stream(j, eff, head, xing, out.data(), (int64_t)out.size(), nf, fa.data(), ba.data(), (int)fa.size(), ns_enc);
}

int main(int argc, char **argv) {
    hmp3_control base;
    hmp3_control_defaults(&base);
    int xing = 3 | 64;  // the reference default: Xing header + TOC + info tag (tomp3.cpp:387)
    int device = 0;
    bool ignore_length = false;
    std::vector<std::string> names;
    const char *list = nullptr;
    int mpeg_select = 0;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (a[0] != '-' || a[1] == 0) {  // a file name, or "-" = stdin/stdout
            names.push_back(a);
            continue;
        }
        if ((a[1] == 'i' || a[1] == 'I') && (a[2] == 'l' || a[2] == 'L')) ignore_length = true;  // tomp3.cpp:521-524
        if (a[1] == '@') {
            list = a[2] ? a + 2 : (i + 1 < argc ? argv[++i] : nullptr);
            continue;
        }
        if ((a[1] == 'g' || a[1] == 'G') && (a[2] == 'p' || a[2] == 'P') && (a[3] == 'u' || a[3] == 'U')) {
            device = atoi(a + 4);  // -GPU<n>: device index (extension)
            continue;
        }
        if (a[1] == 'x' || a[1] == 'X') {
            xing = atoi(a + 2);
            if (xing == 2) xing = 3;
        }
        if (a[1] == 'a' || a[1] == 'A') {  // -A<n>: mpeg_select of MP3_audio_encode_init (tomp3.cpp:545-550)
            mpeg_select = atoi(a + 2);
            if (mpeg_select < 0) mpeg_select = 0;
        }
        if (hmp3_control_apply_option(&base, a) != 0) {
            // -h (anything but -HF) prints the usage text and ends; a letter that is no option is passed over in
            // silence, as the reference's switch does (tomp3.cpp:402-556)
            if (a[1] != 'h' && a[1] != 'H') continue;
            fprintf(stderr, "\n Usage:  hmp3b200 <input> <output> [options]   (options as hmp3; -@ <list> for a batch)\n");
            return 0;
        }
    }
    std::vector<Job> jobs;
    if (list) {
        FILE *f = fopen(list, "r");
        if (!f) {
            fprintf(stderr, "\n CANNOT_OPEN_INPUT_FILE %s\n", list);
            return 1;
        }
        char in[4096], out[4096];
        while (fscanf(f, "%4095s %4095s", in, out) == 2) {
            Job j;
            j.in = in;
            j.out = out;
            jobs.push_back(j);
        }
        fclose(f);
    } else if (names.size() >= 2) {
        Job j;
        j.in = names[0];
        j.out = names[1];
        jobs.push_back(j);
    } else {
        fprintf(stderr, "\n Usage:  hmp3b200 <input> <output> [options]\n");
        return 0;
    }
    if (hmp3_device_count() <= device) {
        fprintf(stderr, "\n NO CUDA DEVICE (this encoder has no CPU path)\n");
        return 1;
    }
    // ---- read inputs, derive each file's control the way ff_encode does (tomp3.cpp:809-815)
    std::vector<hmp3_control> ctl;
    std::vector<int64_t> ns;
    std::vector<int32_t> fmts;
    std::vector<int> idx;
    for (size_t k = 0; k < jobs.size(); k++) {
        Job &j = jobs[k];
        std::string err;
        fprintf(stderr, "\n  PCM input file: %s\nMPEG output file: %s", j.in.c_str(), j.out.c_str());
        if (!read_wav(j.in.c_str(), &j.wav, &err, ignore_length)) {
            fprintf(stderr, "\n %s\n", err.c_str());
            continue;
        }
        j.ec = base;
        const bool mono_convert = (base.mode == 3);  // -M3 on a stereo file: down-mix (tomp3.cpp:560-561, 813-815)
        if (j.ec.mode < 0) j.ec.mode = 0;
        if (j.wav.channels == 1) j.ec.mode = 3;
        j.wav.use_float = !(j.wav.type == 1 && j.wav.bits == 16);
        j.wav.enc_channels = j.wav.channels;
        // the encode rate MP3_audio_encode_init derives from the source rate and -A (mp3enc.cpp:2700-2748)
        j.enc_rate = j.wav.rate;
        const int target = target_rate(j.wav.rate, mpeg_select);
        const bool native = target == j.wav.rate;
        // Csrc case 1 (1:2) is done here, ahead of the batch; any other pair of rates (Csrc cases 2-4) runs through the
        // handle, whose MP3_audio_encode converts call by call exactly as the reference's does
        const bool up2 = target == 2 * j.wav.rate;
        if (target == 0) {
            fprintf(stderr, "\n ENCODER INIT FAIL\n");
            continue;
        }
        if (!native && !up2) {
            j.via_handle = true;
            j.enc_rate = target;
            if (j.wav.channels == 2 && j.ec.mode == 3) j.ec.mode = 1;  // tomp3.cpp:813-815; the down-mix is mono_convert's
            j.ok = true;
            continue;
        }
        const bool downmix = j.wav.channels == 2 && mono_convert;
        if (downmix || up2) {
            const int ch = j.wav.channels;
            const size_t nfr = j.wav.total() / ch;
            const float pad = j.wav.bits == 8 ? -32768.0f : 0.0f;  // what the zero bytes after the end decode to
            auto in = [&](size_t i, int c) -> float {
                if (i >= nfr) return pad;
                return j.wav.use_float ? j.wav.pcmf[ch * i + c] : (float)j.wav.pcm[ch * i + c];
            };
            std::vector<float> y;
            if (!up2) {  // Csrc::src_filter_to_mono_case0 (srccf.cpp:458-468)
                y.resize(nfr);
                for (size_t i = 0; i < nfr; i++) y[i] = (float)((in(i, 0) + in(i, 1)) * 0.5);
            } else if (ch == 1) {  // src_filter_mono_case1 (srccf.cpp:80-100): integer samples, truncated
                y.resize(2 * nfr + 3, pad);
                for (size_t i = 0; i < nfr; i++) {
                    const int a = (int)in(i, 0), b = (int)in(i + 1, 0);
                    y[2 * i] = (float)a;
                    y[2 * i + 1] = (float)((a + b) >> 1);
                }
            } else if (downmix) {  // src_filter_to_mono_case1 (srccf.cpp:472-492)
                y.resize(2 * nfr + 3, pad);
                for (size_t i = 0; i < nfr; i++) {
                    const float a = in(i, 0) + in(i, 1), b = in(i + 1, 0) + in(i + 1, 1);
                    y[2 * i] = (float)(a * 0.5);
                    y[2 * i + 1] = (float)((a + b) * 0.25);
                }
            } else {  // src_filter_dual_case1 (srccf.cpp:258-276)
                y.resize(2 * (2 * nfr + 3), pad);
                for (size_t i = 0; i < nfr; i++)
                    for (int c = 0; c < 2; c++) {
                        y[2 * (2 * i) + c] = in(i, c);
                        y[2 * (2 * i + 1) + c] = (float)((in(i, c) + in(i + 1, c)) * 0.5);
                    }
            }
            // (2 n + 3 samples: the reference's main loop makes floor((n + 3 * 577) / 576) + 1 calls of 576 input
            // samples = the calls of an encoder-rate stream of 2 n + 3 samples; the 3 extra ones are padding)
            j.wav.pcmf.swap(y);
            j.wav.pcm.clear();
            j.wav.use_float = true;
            if (downmix) {
                j.wav.enc_channels = 1;
                j.ec.mode = 3;
            }
            if (up2) {  // band limit of the up-converted signal (mp3enc.cpp:2765-2787)
                j.enc_rate = 2 * j.wav.rate;
                const int cutoff = (int)(0.90f * j.wav.rate / 2);
                int nsb = (64 * cutoff + j.enc_rate / 2) / j.enc_rate;
                if (nsb > 30) nsb = 30;
                if (j.ec.nsb_limit <= 0) j.ec.nsb_limit = 30;
                if (j.ec.nsb_limit > nsb) j.ec.nsb_limit = nsb;
            }
        }
        j.ec.samprate = j.enc_rate;
        hmp3_control eff;
        if (hmp3_effective_control(&j.ec, &eff, nullptr) != HMP3_OK) {
            fprintf(stderr, "\n ENCODER INIT FAIL\n");
            continue;
        }
        j.ok = true;
        ctl.push_back(j.ec);
        ns.push_back((int64_t)(j.wav.total() / j.wav.enc_channels));
        fmts.push_back(j.wav.use_float ? HMP3_PCM_F32 : HMP3_PCM_S16);
        idx.push_back((int)k);
    }
    int rc = 0;
    const bool mono_convert_all = (base.mode == 3);
    if (!ctl.empty()) {
        // ---- one batch on the GPU
        hmp3_batch *b = hmp3_batch_create_ex(ctl.data(), ns.data(), fmts.data(), (int)ctl.size(), device);
        if (!b) {
            fprintf(stderr, "\n ENCODER INIT FAIL: %s\n", hmp3_get_last_error());
            return 1;
        }
        const int n = (int)ctl.size();
        for (int i = 0; i < n; i++)  // the reference flushes with zero BYTES, which 8-bit samples decode to -32768
            if (jobs[idx[i]].wav.bits == 8) hmp3_batch_set_tail(b, i, -32768.0f);
        std::vector<const void *> pcm(n);
        std::vector<std::vector<uint8_t>> out(n);
        std::vector<uint8_t *> outp(n);
        std::vector<int64_t> cap(n), nb(n);
        std::vector<int32_t> nf(n), st(n);
        for (int i = 0; i < n; i++) {
            pcm[i] = fmts[i] == HMP3_PCM_F32 ? (const void *)jobs[idx[i]].wav.pcmf.data() : (const void *)jobs[idx[i]].wav.pcm.data();
            cap[i] = hmp3_batch_out_bound(&ctl[i], ns[i]);
            out[i].resize((size_t)cap[i]);
            outp[i] = out[i].data();
        }
        if (hmp3_batch_encode_host(b, pcm.data(), outp.data(), cap.data(), nb.data(), nf.data(), st.data()) != HMP3_OK) {
            fprintf(stderr, "\n ENCODE FAILED: %s\n", hmp3_get_last_error());
            return 1;
        }
        for (int i = 0; i < n; i++) {
            Job &j = jobs[idx[i]];
            if (st[i] != HMP3_OK) {
                fprintf(stderr, "\n ENCODE FAILED (%d) for %s\n", st[i], j.in.c_str());
                rc = 1;
                continue;
            }
            hmp3_control eff;
            hmp3_mpeg_head head;
            hmp3_effective_control(&ctl[i], &eff, &head);
            const int ncalls_main = (int)calls_main(ns[i]);
            std::vector<int32_t> fa(ncalls_main + 64);
            std::vector<int64_t> ba(ncalls_main + 64);
            int nc = 0;
            if (xing && j.out != "-") {
                nc = hmp3_batch_call_log(b, i, fa.data(), ba.data(), (int)fa.size());
                if (nc > ncalls_main) nc = ncalls_main;
            }
            if (!write_stream(j, eff, head, xing, out[i].data(), nb[i], nf[i], fa.data(), ba.data(), nc, ns[i])) rc = 1;
        }
        hmp3_batch_destroy(b);
    }
    // ---- files that need the general sample-rate converter: one at a time through the handle
    bool any = !ctl.empty();
    for (Job &j : jobs) {
        if (!j.ok || !j.via_handle) continue;
        any = true;
        if (!encode_with_handle(j, mpeg_select, mono_convert_all, device, xing)) rc = 1;
    }
    if (!any) return 1;
    fprintf(stderr, "\n");
    return rc;
}
