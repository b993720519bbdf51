"""TEST INFRASTRUCTURE ONLY: a small MPEG-1 Layer III decoder (numpy), used to measure the decoded-PCM SNR of encodes
(north_star: "decoded-PCM SNR within 0.1 dB of the reference encode" for any mode that is not bit-exact).

Scope: MPEG-1 (32 / 44.1 / 48 kHz), mono / stereo / joint stereo with M/S (no intensity stereo: the built encode path
never produces it), long, start, stop and short blocks (no mixed blocks), bit reservoir, no CRC.  ISO 11172-3 2.4.3.4
(requantisation, reordering, stereo, alias reduction, IMDCT, polyphase synthesis).  The Huffman books are read from
hmp3_b200/csrc/tables_data.h (the generated data tables of this repository); the synthesis window is 32 x the analysis
window, which is measured from the host build of the analysis filterbank (tests/hostsim) by impulse probing."""
import math
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SFB_L = {44100: [0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 52, 62, 74, 90, 110, 134, 162, 196, 238, 288, 342, 418, 576],
         48000: [0, 4, 8, 12, 16, 20, 24, 30, 36, 42, 50, 60, 72, 88, 106, 128, 156, 190, 230, 276, 330, 384, 576],
         32000: [0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 54, 66, 82, 102, 126, 156, 194, 240, 296, 364, 448, 550, 576]}
SFB_S = {44100: [0, 4, 8, 12, 16, 22, 30, 40, 52, 66, 84, 106, 136, 192],
         48000: [0, 4, 8, 12, 16, 22, 28, 38, 50, 64, 80, 100, 126, 192],
         32000: [0, 4, 8, 12, 16, 22, 30, 42, 58, 78, 104, 138, 180, 192]}
SLEN1 = [0, 0, 0, 0, 3, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4]
SLEN2 = [0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3, 1, 2, 3, 2, 3]
PRETAB = [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 2, 0]
BITRATES = [0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320]
RATES = [44100, 48000, 32000]
QUAD_LEN_A = [1, 4, 4, 5, 4, 6, 5, 6, 4, 5, 5, 6, 5, 6, 6, 6]
QUAD_CODE_A = [1, 5, 4, 5, 6, 5, 4, 4, 7, 3, 6, 0, 7, 2, 3, 1]

_tables = None


def _load_tables():
    global _tables
    if _tables is not None:
        return _tables
    text = open(os.path.join(ROOT, "hmp3_b200", "csrc", "tables_data.h")).read()

    def arr(name):
        m = re.search(name + r"\[[0-9]*\]\s*=\s*\{(.*?)\};", text, re.S)
        return [int(t, 0) for t in re.findall(r"0x[0-9a-fA-F]+|-?\d+", m.group(1))]
    book = np.array(arr("kHuffBook"), dtype=np.uint64).reshape(-1, 16, 16)
    sel, linbits, dim = arr("kHuffSelBook"), arr("kHuffLinbits"), arr("kHuffBookDim")
    dec = []
    for b in range(book.shape[0]):
        d = {}
        for x in range(dim[b]):
            for y in range(dim[b]):
                e = int(book[b, x, y])
                ln, code = e >> 24, e & 0xFFFFFF
                if ln:
                    d[(ln, code)] = (x, y)
        dec.append(d)
    quad = {(QUAD_LEN_A[j], QUAD_CODE_A[j]): j for j in range(16)}
    _tables = dict(dec=dec, sel=sel, linbits=linbits, quad=quad)
    return _tables


_window = None


def analysis_window():
    """C[n], n = 0..511 (n = 0 the newest sample), measured from the host build of the encoder's polyphase."""
    global _window
    if _window is not None:
        return _window
    import simmod
    from hmp3_b200 import capi
    ec = capi.control(samprate=44100, nch=1, bitrate=64)
    amp, ng = 16384.0, 6
    resp = {}
    best = (0.0, 0, 0)
    P0 = 2048
    for q in range(64):
        pcm = np.zeros((ng * 576, 1), np.int16)
        pcm[P0 + q, 0] = int(amp)
        sbt = simmod.analysis(ec, pcm, ng, 1)["sbt"]            # [g][ch][18 * sb + t]
        r = sbt[:, 0, 0:18].reshape(-1)                         # sub-band 0 over the global slot index
        resp[q] = r
        s = int(np.argmax(np.abs(r)))
        if abs(r[s]) > best[0]:
            best = (abs(float(r[s])), q, s)
    # n = A - p + 32 s for a constant A.  The response is c[n] cos((n - 16) pi / 64): the largest one only says that n
    # is near the window's peak, so A is searched around that guess for the alignment that yields the known peak
    # value c[256] = 0.035781 with every other |c[n]| below it.
    def window_for(A):
        c = np.zeros(512)
        for n in range(512):
            cs = math.cos((n - 16) * math.pi / 64)
            if abs(cs) < 1e-6:
                continue
            for q in range(64):
                t = n - A + P0 + q
                if t % 32 == 0 and 0 <= t // 32 < resp[q].size:
                    c[n] = resp[q][t // 32] / (amp * cs)
                    break
        return c
    A0 = 256 + (P0 + best[1]) - 32 * best[2]
    c, err = None, 1e9
    for A in range(A0 - 48, A0 + 49):
        w = window_for(A)
        e = abs(float(np.max(np.abs(w))) - 0.035781) + abs(abs(float(w[256])) - 0.035781)
        if e < err:
            c, err = w, e
    if c[256] < 0:
        c = -c
    # the probe used the unfolded operator cos((2 sb + 1)(n - 16) pi / 64), n = 0..511; ISO folds the window into 64
    # phases first, which moves a factor (-1)^(n / 64) from the cosine into the tabulated window C[n]
    c = c * np.array([(-1.0) ** (n // 64) for n in range(512)])
    # the taps at n = 48 + 64 j meet cos((2 sb + 1) pi / 2) = 0 in the analysis and cannot be probed; the synthesis
    # needs them: the tabulated window is antisymmetric, C[n] = -C[512 - n]
    for n in range(48, 512, 64):
        c[n] = -c[512 - n]
    assert abs(np.max(np.abs(c)) - 0.035781) < 1e-4, np.max(np.abs(c))
    _window = c
    return c


class Bits:
    def __init__(self, data):
        self.d = data
        self.pos = 0

    def get(self, n):
        v = 0
        p = self.pos
        for _ in range(n):
            v = (v << 1) | ((self.d[p >> 3] >> (7 - (p & 7))) & 1)
            p += 1
        self.pos = p
        return v


def _imdct_tables():
    n36 = np.array([[math.cos(math.pi / 72 * (2 * i + 1 + 18) * (2 * k + 1)) for k in range(18)] for i in range(36)])
    n12 = np.array([[math.cos(math.pi / 24 * (2 * i + 1 + 6) * (2 * k + 1)) for k in range(6)] for i in range(12)])
    w = np.zeros((4, 36))
    for i in range(36):
        w[0, i] = math.sin(math.pi / 36 * (i + 0.5))
    for i in range(18):
        w[1, i] = math.sin(math.pi / 36 * (i + 0.5))
    for i in range(18, 24):
        w[1, i] = 1.0
    for i in range(24, 30):
        w[1, i] = math.sin(math.pi / 12 * (i - 18 + 0.5))
    for i in range(6, 12):
        w[3, i] = math.sin(math.pi / 12 * (i - 6 + 0.5))
    for i in range(12, 18):
        w[3, i] = 1.0
    for i in range(18, 36):
        w[3, i] = math.sin(math.pi / 36 * (i + 0.5))
    w12 = np.array([math.sin(math.pi / 12 * (i + 0.5)) for i in range(12)])
    return n36, n12, w, w12


def decode(mp3, max_frames=None):
    """mp3: bytes / uint8 array of MPEG-1 Layer III frames -> float64 PCM (nsamples, nch), +-32768 scale."""
    T = _load_tables()
    data = bytes(np.asarray(mp3, dtype=np.uint8).tobytes()) if not isinstance(mp3, (bytes, bytearray)) else bytes(mp3)
    n36, n12, wl, w12 = _imdct_tables()
    ci = [-0.6, -0.535, -0.33, -0.185, -0.095, -0.041, -0.0142, -0.0037]
    cs = np.array([1 / math.sqrt(1 + c * c) for c in ci])
    ca = np.array([c / math.sqrt(1 + c * c) for c in ci])
    D = 32.0 * analysis_window()
    N = np.array([[math.cos((16 + i) * (2 * k + 1) * math.pi / 64) for k in range(32)] for i in range(64)])
    pos, out = 0, []
    reservoir = b""
    state = None
    nfr = 0
    while pos + 4 <= len(data):
        h = int.from_bytes(data[pos:pos + 4], "big")
        if (h >> 21) != 0x7FF:
            pos += 1
            continue
        assert ((h >> 19) & 3) == 3 and ((h >> 17) & 3) == 1, "MPEG-1 Layer III only"
        prot = (h >> 16) & 1
        br, sr = BITRATES[(h >> 12) & 15], RATES[(h >> 10) & 3]
        pad, mode, mode_ext = (h >> 9) & 1, (h >> 6) & 3, (h >> 4) & 3
        nch = 1 if mode == 3 else 2
        flen = 144000 * br // sr + pad
        if pos + flen > len(data):
            break
        side_len = 17 if nch == 1 else 32
        p0 = pos + 4 + (0 if prot else 2)
        sb = Bits(data[p0:p0 + side_len])
        main_data_begin = sb.get(9)
        sb.get(5 if nch == 1 else 3)
        scfsi = [[sb.get(1) for _ in range(4)] for _ in range(nch)]
        gr = [[None] * nch for _ in range(2)]
        for g in range(2):
            for c in range(nch):
                s = dict(part2_3_length=sb.get(12), big_values=sb.get(9), global_gain=sb.get(8), sfc=sb.get(4),
                         ws=sb.get(1))
                if s["ws"]:
                    s["block_type"] = sb.get(2)
                    s["mixed"] = sb.get(1)
                    s["table_select"] = [sb.get(5), sb.get(5), 0]
                    s["subblock_gain"] = [sb.get(3), sb.get(3), sb.get(3)]
                    s["r0"], s["r1"] = (8 if s["block_type"] == 2 else 7), 36
                else:
                    s["block_type"], s["mixed"] = 0, 0
                    s["table_select"] = [sb.get(5), sb.get(5), sb.get(5)]
                    s["subblock_gain"] = [0, 0, 0]
                    s["r0"], s["r1"] = sb.get(4), sb.get(3)
                s["preflag"], s["sf_scale"], s["count1table"] = sb.get(1), sb.get(1), sb.get(1)
                assert not s["mixed"], "mixed blocks are not produced by this encoder"
                gr[g][c] = s
        frame_main = data[p0 + side_len:pos + flen]
        if main_data_begin > len(reservoir):
            reservoir = (reservoir + frame_main)[-4096:]       # not enough history yet (stream start): skip
            pos += flen
            out.append(np.zeros((1152, nch)))
            nfr += 1
            continue
        main = (reservoir[len(reservoir) - main_data_begin:] if main_data_begin else b"") + frame_main
        reservoir = (reservoir + frame_main)[-4096:]
        mb = Bits(main + b"\0" * 8)
        if state is None or state["nch"] != nch:
            state = dict(nch=nch, overlap=np.zeros((nch, 32, 18)), V=np.zeros((nch, 1024)), sf_l=np.zeros((nch, 22), int))
        sfl, sfs = SFB_L[sr], SFB_S[sr]
        pcm = np.zeros((1152, nch))
        for g in range(2):
            xr = np.zeros((nch, 576))
            for c in range(nch):
                s = gr[g][c]
                start = mb.pos
                sl1, sl2 = SLEN1[s["sfc"]], SLEN2[s["sfc"]]
                sf_l = np.zeros(22, int)
                sf_s = np.zeros((13, 3), int)
                if s["block_type"] == 2:
                    for b in range(12):
                        for w in range(3):
                            sf_s[b][w] = mb.get(sl1 if b < 6 else sl2)
                else:
                    for k, (lo, hi) in enumerate([(0, 6), (6, 11), (11, 16), (16, 21)]):
                        if g == 1 and scfsi[c][k]:
                            sf_l[lo:hi] = state["sf_l"][c][lo:hi]
                        else:
                            for b in range(lo, hi):
                                sf_l[b] = mb.get(sl1 if b < 11 else sl2)
                    state["sf_l"][c] = sf_l
                # ---- Huffman
                end = start + s["part2_3_length"]
                iv = np.zeros(576 + 4, int)
                big = min(2 * s["big_values"], 576)
                if s["ws"]:
                    r1s = 36 if s["block_type"] == 2 else sfl[8]
                    r2s = 576
                else:
                    r1s = sfl[min(s["r0"] + 1, 22)]
                    r2s = sfl[min(s["r0"] + s["r1"] + 2, 22)]
                k = 0
                while k < big:
                    t = s["table_select"][0 if k < r1s else (1 if k < r2s else 2)]
                    bookd = T["dec"][T["sel"][t]]
                    x = y = 0
                    if bookd and t not in (0, 4, 14):
                        ln, code = 0, 0
                        while True:
                            code = (code << 1) | mb.get(1)
                            ln += 1
                            hit = bookd.get((ln, code))
                            if hit is not None:
                                x, y = hit
                                break
                            assert ln < 24, "bad Huffman data"
                        lb = T["linbits"][t]
                        if t >= 16 and x == 15:
                            x += mb.get(lb)
                        if x and mb.get(1):
                            x = -x
                        if t >= 16 and y == 15:
                            y += mb.get(lb)
                        if y and mb.get(1):
                            y = -y
                    iv[k], iv[k + 1] = x, y
                    k += 2
                while mb.pos < end and k <= 572:
                    if s["count1table"]:
                        j = mb.get(4) ^ 15
                    else:
                        ln, code = 0, 0
                        while True:
                            code = (code << 1) | mb.get(1)
                            ln += 1
                            j = T["quad"].get((ln, code))
                            if j is not None:
                                break
                            assert ln < 8
                    for q in range(4):
                        v = (j >> (3 - q)) & 1
                        if v and mb.get(1):
                            v = -v
                        iv[k + q] = v
                    k += 4
                if mb.pos > end:                       # the last quad ran past the granule's data: discard it
                    k -= 4
                    iv[k:k + 4] = 0
                mb.pos = end
                # ---- requantise
                mag = np.abs(iv[:576]).astype(np.float64) ** (4.0 / 3.0) * np.sign(iv[:576])
                gain = 2.0 ** ((s["global_gain"] - 210) / 4.0)
                mult = 0.5 * (1 + s["sf_scale"])
                if s["block_type"] == 2:
                    y = np.zeros(576)
                    k = 0
                    for b in range(13):
                        wdt = sfs[b + 1] - sfs[b]
                        for w in range(3):
                            gw = 2.0 ** ((s["global_gain"] - 210 - 8 * s["subblock_gain"][w]) / 4.0)
                            f = gw * 2.0 ** (-mult * (sf_s[b][w] if b < 12 else 0))
                            for i in range(wdt):
                                y[3 * sfs[b] + 3 * i + w] = mag[k] * f         # reordered: windows interleaved by line
                                k += 1
                    xr[c] = y
                else:
                    f = np.ones(576)
                    for b in range(22):
                        f[sfl[b]:sfl[b + 1]] = 2.0 ** (-mult * (sf_l[b] + s["preflag"] * PRETAB[b]))
                    xr[c] = mag * gain * f
            if nch == 2 and mode == 1 and (mode_ext & 2):
                m, d = xr[0].copy(), xr[1].copy()
                xr[0], xr[1] = (m + d) / math.sqrt(2.0), (m - d) / math.sqrt(2.0)
            for c in range(nch):
                s = gr[g][c]
                x = xr[c].reshape(32, 18).copy()
                if s["block_type"] != 2:
                    for b in range(1, 32):
                        lo = x[b - 1, 17 - np.arange(8)].copy()
                        hi = x[b, np.arange(8)].copy()
                        x[b - 1, 17 - np.arange(8)] = lo * cs - hi * ca
                        x[b, np.arange(8)] = hi * cs + lo * ca
                ts = np.zeros((32, 18))
                for b in range(32):
                    if s["block_type"] == 2:
                        raw = np.zeros(36)
                        for w in range(3):
                            raw[6 + 6 * w:18 + 6 * w] += (n12 @ x[b, w::3]) * w12
                    else:
                        raw = (n36 @ x[b]) * wl[s["block_type"]]
                    ts[b] = raw[:18] + state["overlap"][c, b]
                    state["overlap"][c, b] = raw[18:]
                ts[1::2, 1::2] *= -1.0
                V = state["V"][c]
                for t in range(18):
                    V = np.concatenate([N @ ts[:, t], V[:960]])
                    U = np.zeros(512)
                    for i in range(8):
                        U[64 * i:64 * i + 32] = V[128 * i:128 * i + 32]
                        U[64 * i + 32:64 * i + 64] = V[128 * i + 96:128 * i + 128]
                    pcm[576 * g + 32 * t:576 * g + 32 * t + 32, c] = 32768.0 * (U * D).reshape(16, 32).sum(axis=0)
                state["V"][c] = V
        out.append(pcm)
        pos += flen
        nfr += 1
        if max_frames and nfr >= max_frames:
            break
    return np.concatenate(out) if out else np.zeros((0, 1))


def snr_db(ref_pcm, dec_pcm, search=(900, 2400)):
    """SNR of a decode against the source PCM after aligning the codec delay (integer samples, searched)."""
    a = np.asarray(ref_pcm, dtype=np.float64)
    b = np.asarray(dec_pcm, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    n = min(a.shape[0], b.shape[0] - search[1]) - 2304
    seg = a[1152:1152 + n]
    best = None
    probe = seg[:20000, 0]
    for lag in range(search[0], search[1]):
        c = float(np.dot(probe, b[1152 + lag:1152 + lag + probe.size, 0]))
        if best is None or c > best[0]:
            best = (c, lag)
    lag = best[1]
    d = b[1152 + lag:1152 + lag + n] - seg
    return 10.0 * math.log10(float((seg ** 2).sum()) / max(float((d ** 2).sum()), 1e-30)), lag
