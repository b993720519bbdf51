// Host post-pass: the Xing/Info frame the reference CLI writes in front of the audio frames (seek table, frame and
// byte counts, LAME-style info tag with encoder delay, MusicCRC and tag CRC).  Pure host code, no device needed.
// Behaviour follows XingHeader / XingHeaderTOC / XingHeaderUpdateInfo / BuildTOC (hmp3/src/xhead.c:117-719) as
// driven by ff_encode (hmp3/src/test/tomp3.cpp:682-689, 871-896, 962-984, 1012-1072).
#include <stdint.h>
#include <string.h>

#include "../../include/hmp3_b200.h"
#include "enc_init.h"

namespace {

enum { FRAMES_FLAG = 1, BYTES_FLAG = 2, TOC_FLAG = 4, VBR_SCALE_FLAG = 8, RESERVEDA_FLAG = 16, RESERVEDB_FLAG = 32, INFOTAG_FLAG = 64 };

const int kRates[6] = {22050, 24000, 16000, 44100, 48000, 32000};
const int kBitrates[2][16] = {{0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160, 0},
                              {0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 0}};

void put_be32(uint8_t *p, uint32_t x) {
    p[0] = (uint8_t)(x >> 24);
    p[1] = (uint8_t)(x >> 16);
    p[2] = (uint8_t)(x >> 8);
    p[3] = (uint8_t)x;
}

// reflected CRC-16 (polynomial 0x8005), one byte at a time (xhead.c:181-221)
uint16_t crc16_step(uint16_t crc, uint8_t d) {
    crc ^= d;
    for (int j = 0; j < 8; j++) crc = (crc & 1) ? (uint16_t)((crc >> 1) ^ 0xA001) : (uint16_t)(crc >> 1);
    return crc;
}

// The seek-point table the CLI fills while encoding: one (frames, bytes) entry every `stride` calls, decimated by
// two whenever 512 entries are reached (xhead.c:684-709).
struct SeekTable {
    static const int N = 512;
    long long tab[N + 1][2];
    int count = 0, stride = 1;
    int add(long long frames, long long bytes) {
        tab[count][0] = frames;
        tab[count][1] = bytes;
        count++;
        if (count < N) return stride;
        for (int k = 1, i = 0; i < N / 2; i++, k += 2) {
            tab[i][0] = tab[k][0];
            tab[i][1] = tab[k][1];
        }
        count = N / 2;
        stride += stride;
        return stride;
    }
};

// 100 seek points: byte position (in 1/256 of the stream) of every percent of the frames (xhead.c:117-167)
void build_toc(SeekTable &t, long long tot_frames, long long tot_bytes, uint8_t *buf) {
    if (tot_frames <= 0 || tot_bytes <= 0) {
        memset(buf, 0, 100);
        return;
    }
    t.tab[t.count][0] = tot_frames;
    t.tab[t.count][1] = tot_bytes;
    t.count++;
    for (int i = 0; i < t.count; i++) t.tab[i][0] = 100 * t.tab[i][0];
    const double a = 256.0 / (double)(int)tot_bytes;
    long long target = 0, t0 = 0, t0b = 0;
    int k = 0;
    for (int i = 0; i < 100; i++) {
        while (t.tab[k][0] <= target) {
            t0 = t.tab[k][0];
            t0b = t.tab[k][1];
            k++;
        }
        const double b = (double)(int)t0b + ((double)(int)(target - t0)) * ((double)(int)(t.tab[k][1] - t0b)) /
                                                ((double)(int)(t.tab[k][0] - t0));
        int index = (int)(a * b + 0.5);
        if (index < 0) index = 0;
        if (index > 255) index = 255;
        buf[i] = (uint8_t)index;
        target += tot_frames;
    }
}

}  // namespace

extern "C" {

int hmp3_effective_control(const hmp3_control *ec, hmp3_control *out, hmp3_mpeg_head *head) {
    hmp3::EncTables *T = new hmp3::EncTables;
    int unsup = 0;
    const int r = hmp3::build_tables(ec, T, &unsup);
    if (r) {
        if (out) memcpy(out, T->cfg.info_ec, sizeof(*out));
        if (head) {
            memset(head, 0, sizeof(*head));
            head->sync = 1;
            head->id = T->cfg.h_id;
            head->option = 1;
            head->prot = 1;
            head->br_index = T->cfg.br_index;
            head->sr_index = T->cfg.sr_index;
            head->mode = T->cfg.h_mode;
            head->mode_ext = (T->cfg.head[3] >> 4) & 3;
            head->cr = (T->cfg.head[3] >> 3) & 1;
            head->original = (T->cfg.head[3] >> 2) & 1;
        }
    }
    delete T;
    return r ? HMP3_OK : (unsup ? HMP3_ERR_UNSUPPORTED : HMP3_ERR_BAD_CONTROL);
}

int hmp3_info_frame_size(const hmp3_control *eff, int head_mode, int xing_flag, int channels) {
    uint8_t tmp[2048];
    return hmp3_info_frame(eff, head_mode, xing_flag, eff->samprate, channels, 0, nullptr, 0, 0, nullptr, nullptr, 0, tmp,
                           (int)sizeof(tmp));
}

int hmp3_info_frame(const hmp3_control *eff, int head_mode, int xing_flag, int source_rate, int channels,
                    int64_t nsamples, const uint8_t *audio, int64_t audio_bytes, uint32_t frames,
                    const int32_t *frames_after_call, const int64_t *bytes_after_call, int ncalls, uint8_t *buf, int cap) {
    if (!eff || !buf || xing_flag == 0) return 0;
    // ---- flags as the CLI derives them from -X (tomp3.cpp:682-689); default -X is 3 | INFOTAG
    if (xing_flag == 2) xing_flag = 3;
    int cli_flags = FRAMES_FLAG | BYTES_FLAG | VBR_SCALE_FLAG;
    if (xing_flag & 2) cli_flags |= TOC_FLAG | INFOTAG_FLAG;
    const int vbr_scale = eff->vbr_flag ? eff->vbr_mnr : -1;
    const int nbitrate = eff->bitrate * channels;
    // ---- XingHeader (xhead.c:255-470)
    int head_flags = cli_flags & 127;
    const int h_mode = head_mode & 3;
    int sr_index;
    for (sr_index = 0; sr_index < 6; sr_index++)
        if (eff->samprate == kRates[sr_index]) break;
    if (sr_index >= 6) return 0;
    int h_id = 0;
    if (sr_index >= 3) {
        h_id = 1;
        sr_index -= 3;
    }
    const int side_bytes = h_id ? (h_mode == 3 ? 17 : 32) : (h_mode == 3 ? 9 : 17);
    int nbr_index = 0;
    for (int i = 1; i < 15; i++)
        if (kBitrates[h_id & 1][i] == nbitrate) {
            nbr_index = i;
            break;
        }
    if (vbr_scale == -1 && kBitrates[h_id][nbr_index] < 64) head_flags &= ~TOC_FLAG;
    int need = 4 + side_bytes + 8;
    if (head_flags & FRAMES_FLAG) need += 4;
    if (head_flags & BYTES_FLAG) need += 4;
    if (head_flags & TOC_FLAG) need += 100;
    if (head_flags & VBR_SCALE_FLAG) need += 4;
    if (head_flags & RESERVEDA_FLAG) need += 20;
    if (head_flags & RESERVEDB_FLAG) need += 20;
    if (head_flags & INFOTAG_FLAG) need += 36;
    const int tmp = h_id ? eff->samprate : 2 * eff->samprate;
    int br_index, frame_bytes = 0;
    if (vbr_scale != -1) {
        for (br_index = 1; br_index < 15; br_index++) {
            frame_bytes = 144000 * kBitrates[h_id][br_index] / tmp;
            if (frame_bytes >= need) break;
        }
        if (br_index >= 15) return 0;
    } else {
        if (nbr_index >= 15) return 0;
        frame_bytes = 144000 * kBitrates[h_id][nbr_index] / tmp;
        if (frame_bytes < need) return 0;
        br_index = nbr_index;
    }
    if (frame_bytes > cap) return 0;
    memset(buf, 0, frame_bytes);
    buf[0] = 0xFF;
    buf[1] = (uint8_t)(0xF3 | (h_id << 3));
    buf[2] = (uint8_t)((br_index << 4) | (sr_index << 2));
    buf[3] = (uint8_t)((h_mode << 6) | ((eff->cr_bit & 1) << 3) | ((eff->original & 1) << 2));
    uint8_t *p = buf + 4 + side_bytes;
    memcpy(p, vbr_scale != -1 ? "Xing" : "Info", 4);
    p += 4;
    put_be32(p, (uint32_t)head_flags);
    p += 4;
    if (!audio) {  // the frame XingHeader() itself writes before encoding (xhead.c:396-462): counts and seek table
                   // zero, quality field set; it stays in the file when the output cannot be updated (stdout)
        if (head_flags & FRAMES_FLAG) p += 4;
        if (head_flags & BYTES_FLAG) p += 4;
        if (head_flags & TOC_FLAG) p += 100;
        if (head_flags & VBR_SCALE_FLAG) put_be32(p, (uint32_t)vbr_scale);
        return frame_bytes;
    }

    // ---- the seek table as the CLI's main loop fills it (tomp3.cpp:976-984)
    SeekTable *toc = new SeekTable;
    if (cli_flags & TOC_FLAG) {
        int counter = 0;
        for (int u = 0; u < ncalls; u++) {
            counter--;
            if (counter <= 0) counter = toc->add((long long)frames_after_call[u] + 1, bytes_after_call[u] + frame_bytes);
        }
    }
    // ---- MusicCRC over every audio byte (tomp3.cpp:962, 1012, 1035)
    uint16_t music = 0;
    for (int64_t i = 0; i < audio_bytes; i++) music = crc16_step(music, audio[i]);
    // ---- XingHeaderUpdateInfo (xhead.c:484-680)
    const long long bs_bytes = frame_bytes + audio_bytes;
    const uint64_t samples_audio = (uint64_t)nsamples;
    unsigned in_rate = (unsigned)source_rate, out_rate = (unsigned)eff->samprate;
    if (in_rate == 0 || out_rate == 0) in_rate = out_rate = 1;
    const int pad_start = 1680;
    int pad_end = 0;
    if (samples_audio > 0) {
        const uint64_t per = (h_id == 1 ? 1152 : 576);
        const uint64_t sa = (uint64_t)((double)samples_audio * ((double)out_rate / (double)in_rate) + 0.5);
        uint64_t smp3 = (uint64_t)frames * per;
        if (smp3 - sa - pad_start >= 4096) {
            frames = (unsigned)((sa + pad_start + 1152) / per);
            smp3 = frames * per;
        }
        const long pad_total = (long)(smp3 - sa);
        pad_end = (int)(pad_total - pad_start);
    }
    if (head_flags & FRAMES_FLAG) {
        put_be32(p, frames);
        p += 4;
    }
    if (head_flags & BYTES_FLAG) {
        put_be32(p, (uint32_t)bs_bytes);
        p += 4;
    }
    if (head_flags & TOC_FLAG) {
        build_toc(*toc, (long long)(int)frames, (long long)(int)bs_bytes, p);
        p += 100;
    }
    delete toc;
    if (head_flags & VBR_SCALE_FLAG) {
        put_be32(p, (uint32_t)vbr_scale);
        p += 4;
    }
    if (head_flags & RESERVEDA_FLAG) p += 20;
    if (head_flags & RESERVEDB_FLAG) p += 20;
    if ((head_flags & INFOTAG_FLAG) && samples_audio != 0) {
        memcpy(p, "LAMEH5.24", 9);
        p += 9;
        *p++ = (uint8_t)((0x0 << 4) | (vbr_scale != -1 ? 0x0 : 0x1));
        *p++ = (uint8_t)(((unsigned)eff->freq_limit / 100) & 0xFF);
        put_be32(p, 0);
        p += 4;
        put_be32(p, 0);
        p += 4;
        *p++ = 0;
        *p++ = 0;
        p[0] = (uint8_t)((pad_start >> 4) & 0xFF);
        p[1] = (uint8_t)(((pad_start << 4) & 0xF0) | ((pad_end >> 8) & 0x0F));
        p[2] = (uint8_t)(pad_end & 0xFF);
        p += 3;
        uint8_t misc = 0x1C;
        if (in_rate == 44100) misc |= 0x40;
        else if (in_rate == 48000) misc |= 0x80;
        else if (in_rate > 48000) misc |= 0xC0;
        *p++ = misc;
        *p++ = 0;
        *p++ = 0;
        *p++ = 0;
        put_be32(p, (uint32_t)bs_bytes);
        p += 4;
        *p++ = (uint8_t)(music >> 8);
        *p++ = (uint8_t)(music & 0xFF);
        uint16_t tag = 0;
        for (const uint8_t *q = buf; q < p; q++) tag = crc16_step(tag, *q);
        *p++ = (uint8_t)(tag >> 8);
        *p++ = (uint8_t)(tag & 0xFF);
    }
    return frame_bytes;
}

}  // extern "C"
