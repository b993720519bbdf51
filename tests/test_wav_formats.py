"""8/24/32-bit integer and 32-bit float WAV input.  The reference CLI feeds every sample type through
Csrc::sr_convert (hmp3/src/srcc.cpp:804-834) to the float entry of the encoder.  CPU: the conversion restated in
tests/wavutil.py + the float path of the kernel bodies (host simulator) reproduce the audio frames of files written by
the reference CLI.  GPU: `hmp3b200` writes the same file as `hmp3` for each sample type, the float batch entry and the
handle's non-16-bit entry equal the oracle."""
import os
import subprocess

import numpy as np
import pytest

import refmod
import simmod
import wavutil
from hmp3_b200 import capi
from hmp3_b200.synth import synth_pcm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "hmp3")
CLI = os.path.join(ROOT, "hmp3_b200", "_lib", "hmp3b200")
needs_ref = pytest.mark.skipif(not (os.path.exists(REF_BIN) and refmod.available()), reason="oracle/_ref not built")

CASES = [(44100, 2, ["-B64"], dict(bitrate=64)), (22050, 1, ["-B32"], dict(bitrate=32)),
         (48000, 2, ["-V100", "-HF2", "-F19000", "-S1"], dict(vbr_mnr=100, hf=2, freq_limit=19000, filter_select=1))]


def ref_cli_file(tmp_path, samples, kind, sr, nch, opts, name="a", binary=None):
    wav, mp3 = str(tmp_path / (name + ".wav")), str(tmp_path / (name + "_ref.mp3"))
    wavutil.write_wav(wav, samples, kind, sr, nch)
    subprocess.run([binary or REF_BIN, wav, mp3] + opts, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    return wav, np.fromfile(mp3, dtype=np.uint8)


@needs_ref
@pytest.mark.parametrize("kind", wavutil.KINDS)
@pytest.mark.parametrize("sr,nch,opts,kw", CASES)
def test_host_float_path_reproduces_reference_cli_audio(tmp_path, kind, sr, nch, opts, kw):
    samples = wavutil.make_samples(synth_pcm(31, 2.6, sr, nch), kind, seed=5)
    _, whole = ref_cli_file(tmp_path, samples, kind, sr, nch, opts)
    src = samples if kind == "s16" else wavutil.to_encoder_float(samples, kind)
    got, _, _ = simmod.encode_clip(capi.control(samprate=sr, nch=nch, **kw), src, tail=wavutil.tail_value(kind))
    head = whole.size - got.size
    assert head > 0 and np.array_equal(whole[head:], got)
    # ... and the oracle library, fed the converted floats, agrees wherever the flush is silence
    if kind != "u8":
        ref, _ = refmod.ref_encode_clip(refmod.make_ec(samprate=sr, nch=nch, **kw), src)
        assert np.array_equal(ref, got)


@pytest.mark.gpu
@needs_ref
def test_cli_identity_for_every_wav_sample_type(tmp_path):
    assert os.path.exists(CLI)
    for kind in wavutil.KINDS:
        for k, (sr, nch, opts, kw) in enumerate(CASES):
            samples = wavutil.make_samples(synth_pcm(41 + k, 3.3, sr, nch), kind, seed=6)
            name = "%s_%d" % (kind, k)
            wav, want = ref_cli_file(tmp_path, samples, kind, sr, nch, opts, name)
            out = str(tmp_path / (name + "_gpu.mp3"))
            subprocess.run([CLI, wav, out] + opts, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
            got = np.fromfile(out, dtype=np.uint8)
            assert got.size == want.size and np.array_equal(got, want), name


@pytest.mark.gpu
@needs_ref
def test_float_and_int16_streams_in_one_batch():
    """Float streams (fractional sample values, with and without the DC filter) next to int16 streams, ragged."""
    rng = np.random.default_rng(7)
    ctl, pcms, refs = [], [], []
    specs = [(44100, 2, dict(bitrate=64), 3.1, True), (44100, 2, dict(bitrate=64), 2.0, False),
             (48000, 2, dict(vbr_mnr=100, hf=2, freq_limit=19000, filter_select=1), 2.4, True),
             (22050, 1, dict(bitrate=32), 4.2, True), (32000, 2, dict(vbr_mnr=50), 1.3, True),
             (44100, 2, dict(vbr_mnr=50, filter_select=1), 1.9, False), (44100, 1, dict(bitrate=48), 0.0, True)]
    for k, (sr, nch, kw, secs, is_float) in enumerate(specs):
        pcm = synth_pcm(500 + k, max(secs, 0.01), sr, nch)
        if secs == 0.0:
            pcm = pcm[:0]
        if is_float:
            pcm = (pcm.astype(np.float32) + rng.uniform(-0.5, 0.5, size=pcm.shape).astype(np.float32)).astype(np.float32)
        ctl.append(capi.control(samprate=sr, nch=nch, **kw))
        pcms.append(pcm)
        refs.append(refmod.ref_encode_clip(refmod.make_ec(samprate=sr, nch=nch, **kw), pcm)[0])
    outs = capi.encode_batch(ctl, pcms)
    for k, (o, r) in enumerate(zip(outs, refs)):
        assert o.size == r.size and np.array_equal(o, r), "stream %d differs" % k
    # device-resident leg with float uploads
    b = capi.Batch(ctl, [p.shape[0] for p in pcms],
                   formats=[capi.PCM_F32 if p.dtype == np.float32 else capi.PCM_S16 for p in pcms])
    for i, p in enumerate(pcms):
        b.upload(i, p)
    b.run()
    flat, off, nb, nf, st = b.download_all()
    assert (st == 0).all()
    for i, r in enumerate(refs):
        assert np.array_equal(flat[off[i]:off[i] + nb[i]], r), i
    b.close()


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("kind", ["u8", "s24", "s32", "f32"])
def test_handle_mp3_entry_takes_every_sample_type(kind):
    """hmp3_MP3_audio_encode with source bits/is_float as CMp3Enc::MP3_audio_encode_init takes them
    (mp3enc.cpp:2655-2810): raw bytes of the source type per call, converted like Csrc::sr_convert."""
    sr, nch, kw = 44100, 2, dict(bitrate=64)
    samples = wavutil.make_samples(synth_pcm(77, 1.0, sr, nch)[:24 * 1152], kind, seed=8)
    f = wavutil.to_encoder_float(samples, kind)
    bits = {"u8": 8, "s24": 24, "s32": 32, "f32": 32}[kind]
    bps = bits // 8
    enc = capi.Encoder(capacity_seconds=10)
    assert enc.init_mp3(capi.control(samprate=sr, nch=nch, **kw), source_bits=bits,
                        source_is_float=1 if kind == "f32" else 0) == 1153 * nch * bps
    raw = np.frombuffer(wavutil.raw_bytes(samples, kind), dtype=np.uint8)
    zero = np.zeros(1152 * nch * bps, np.uint8)
    out = []
    for c in range(24 + 4 + 6):
        blk = raw[c * 1152 * nch * bps:(c + 1) * 1152 * nch * bps] if c < 24 else zero
        used, b = enc.encode_mp3(blk, raw=True)
        assert used == 1152 * nch * bps
        out.append(b)
    got = np.concatenate(out)
    enc.close()
    # reference: the same floats through the oracle; the zero-byte flush is the tail (u8: -32768)
    if kind == "u8":
        want, _, _ = simmod.encode_clip(capi.control(samprate=sr, nch=nch, **kw), f, tail=-32768.0) \
            if simmod.available() else (None, 0, 0)
        if want is None:
            pytest.skip("host simulator not built")
    else:
        want, _ = refmod.ref_encode_clip(refmod.make_ec(samprate=sr, nch=nch, **kw), f)
    m = min(got.size, want.size)
    assert m > 0.9 * want.size and np.array_equal(got[:m], want[:m])


@needs_ref
@pytest.mark.parametrize("sr,nch,opts,kw", CASES)
def test_stdout_keeps_the_placeholder_info_frame(tmp_path, sr, nch, opts, kw):
    """`hmp3 in.wav -` cannot re-read stdout, so its first frame is what XingHeader() wrote before encoding
    (xhead.c:255-462); hmp3_info_frame(audio=NULL) builds that frame."""
    import ctypes as C
    pcm = synth_pcm(13, 1.5, sr, nch)
    wav = str(tmp_path / "a.wav")
    wavutil.write_wav(wav, pcm, "s16", sr, nch)
    so = np.frombuffer(subprocess.run([REF_BIN, wav, "-"] + opts, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                      check=True).stdout, dtype=np.uint8)
    eff, head = capi.effective_control(capi.control(samprate=sr, nch=nch, **kw))
    buf = np.zeros(2048, np.uint8)
    f = capi.lib().hmp3_info_frame
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p,
                  C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    n = f(eff.ctypes.data_as(C.c_void_p), head.mode, 67, sr, nch, pcm.shape[0], None, 0, 0, None, None, 0,
          buf.ctypes.data_as(C.c_void_p), buf.size)
    assert n > 0 and np.array_equal(so[:n], buf[:n])


@pytest.mark.gpu
@needs_ref
def test_cli_pipes_and_ignore_length(tmp_path):
    """`-` as input (stdin, data length ignored), `-` as output (stdout, placeholder Info frame) and -IL, each
    against the reference CLI run the same way."""
    sr, nch, opts = 44100, 2, ["-B64"]
    samples = wavutil.make_samples(synth_pcm(91, 2.2, sr, nch), "s24", seed=9)
    wav = str(tmp_path / "p.wav")
    wavutil.write_wav(wav, samples, "s24", sr, nch)
    with open(wav, "ab") as f:                        # trailing bytes after the data chunk: audio only under -IL / stdin
        f.write(b"LIST" + (20).to_bytes(4, "little") + bytes(range(20)))

    def run(binary, src, dst, extra=()):
        with open(wav, "rb") as fin:
            r = subprocess.run([binary, src, dst] + opts + list(extra), stdin=fin if src == "-" else None,
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True)
        return np.frombuffer(r.stdout, dtype=np.uint8) if dst == "-" else np.fromfile(dst, dtype=np.uint8)

    for src, dst, extra in [("-", "out", ()), (wav, "-", ()), ("-", "-", ()), (wav, "out", ("-IL",)), (wav, "out", ())]:
        a = run(REF_BIN, src, str(tmp_path / "ref.mp3") if dst == "out" else "-", extra)
        b = run(CLI, src, str(tmp_path / "gpu.mp3") if dst == "out" else "-", extra)
        diff = np.nonzero(a[:min(a.size, b.size)] != b[:min(a.size, b.size)])[0]
        assert a.size > 1000 and a.size == b.size and diff.size == 0, (src, dst, extra, a.size, b.size, diff[:8])


@needs_ref
@pytest.mark.parametrize("kind", ["s16", "s24", "u8"])
@pytest.mark.parametrize("sr,opts,kw", [(44100, ["-B64", "-M3"], dict(bitrate=64)), (22050, ["-M3"], dict())])
def test_mono_downmix_reproduces_reference_cli_audio(tmp_path, kind, sr, opts, kw):
    """-M3 on a stereo file: the reference down-mixes (L + R) * 0.5 and encodes mono (tomp3.cpp:560-561, 813-820;
    srccf.cpp:458-468)."""
    samples = wavutil.make_samples(synth_pcm(51, 2.4, sr, 2), kind, seed=4)
    _, whole = ref_cli_file(tmp_path, samples, kind, sr, 2, opts)
    mono = wavutil.downmix(wavutil.to_encoder_float(samples, kind))
    got, _, _ = simmod.encode_clip(capi.control(samprate=sr, nch=1, **kw), mono, tail=wavutil.tail_value(kind))
    head = whole.size - got.size
    assert head > 0 and np.array_equal(whole[head:], got)


@pytest.mark.gpu
@needs_ref
def test_cli_and_handle_mono_downmix(tmp_path):
    sr = 44100
    for kind, opts in [("s16", ["-B64", "-M3"]), ("s24", ["-M3", "-V80"]), ("f32", ["-B48", "-M3"])]:
        samples = wavutil.make_samples(synth_pcm(61, 3.0, sr, 2), kind, seed=2)
        wav, want = ref_cli_file(tmp_path, samples, kind, sr, 2, opts, "dm_" + kind)
        out = str(tmp_path / ("dm_%s_gpu.mp3" % kind))
        subprocess.run([CLI, wav, out] + opts, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
        got = np.fromfile(out, dtype=np.uint8)
        assert got.size == want.size and np.array_equal(got, want), kind
    # the handle: MP3_audio_encode_init(..., mono_convert = 1) with a two-channel 16-bit source
    samples = wavutil.make_samples(synth_pcm(62, 1.0, sr, 2)[:24 * 1152], "s16")
    enc = capi.Encoder(capacity_seconds=10)
    assert enc.init_mp3(capi.control(samprate=sr, nch=2, bitrate=64), mono_convert=1) == 1153 * 2 * 2
    zero = np.zeros((1152, 2), np.int16)
    out = []
    for c in range(24 + 4 + 6):
        used, b = enc.encode_mp3(samples[c * 1152:(c + 1) * 1152] if c < 24 else zero)
        assert used == 1152 * 2 * 2
        out.append(b)
    got = np.concatenate(out)
    enc.close()
    want, _ = refmod.ref_encode_clip(refmod.make_ec(samprate=sr, nch=1, bitrate=64), wavutil.downmix(samples.astype(np.float32)))
    m = min(got.size, want.size)
    assert m > 0.9 * want.size and np.array_equal(got[:m], want[:m])


EDGE_LENGTHS = [5, 1149, 1152 * 40 + 1148, 1152 * 40 + 1149, 1152 * 40 + 1150, 1152 * 40 + 1151, 1152 * 41, 1152 * 41 + 1]


@needs_ref
@pytest.mark.parametrize("sr,nch,opts,kw", CASES[:2])
def test_end_of_file_call_count_matches_reference_cli(tmp_path, sr, nch, opts, kw):
    """The CLI keeps calling the encoder while bytes_in_init = 1153 sample frames are buffered (what
    Csrc::sr_convert_init returns, srcc.cpp:185-187, 769-773) after appending 4 x 1153 frames of zero bytes: stream
    lengths with n % 1152 in {1149, 1150, 1151} get one more frame than "n plus four frames" would give."""
    base = synth_pcm(71, 1.2, sr, nch)
    for n in EDGE_LENGTHS:
        _, whole = ref_cli_file(tmp_path, base[:n], "s16", sr, nch, opts)
        got, _, _ = simmod.encode_clip(capi.control(samprate=sr, nch=nch, **kw), base[:n])
        ref, _ = refmod.ref_encode_clip(refmod.make_ec(samprate=sr, nch=nch, **kw), base[:n])
        head = whole.size - got.size
        assert head > 0 and np.array_equal(whole[head:], got), n
        assert np.array_equal(ref, got), n


@pytest.mark.gpu
@needs_ref
def test_cli_identity_at_end_of_file_edge_lengths(tmp_path):
    sr, nch, opts = 44100, 2, ["-B64"]
    base = synth_pcm(72, 1.2, sr, nch)
    lst = str(tmp_path / "edge.txt")
    wants = []
    with open(lst, "w") as f:
        for n in EDGE_LENGTHS:
            wav, want = ref_cli_file(tmp_path, base[:n], "s16", sr, nch, opts, "e%d" % n)
            wants.append(want)
            f.write("%s %s\n" % (wav, tmp_path / ("e%d_gpu.mp3" % n)))
    subprocess.run([CLI, "-@", lst] + opts, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    for n, want in zip(EDGE_LENGTHS, wants):
        got = np.fromfile(str(tmp_path / ("e%d_gpu.mp3" % n)), dtype=np.uint8)
        assert got.size == want.size and np.array_equal(got, want), n


FUZZ_COMBOS = [(44100, 2, ["-B64"], dict(bitrate=64)), (44100, 2, [], dict()),
               (48000, 2, ["-V100", "-HF2", "-F19000"], dict(vbr_mnr=100, hf=2, freq_limit=19000)),
               (22050, 1, ["-B32"], dict(bitrate=32)), (32000, 2, [], dict()), (16000, 2, ["-B24"], dict(bitrate=24)),
               (24000, 1, ["-V30"], dict(vbr_mnr=30)), (44100, 1, ["-B96"], dict(bitrate=96)),
               (48000, 2, ["-B160"], dict(bitrate=160)), (32000, 2, ["-B160"], dict(bitrate=160)),
               (44100, 2, ["-V150"], dict(vbr_mnr=150)), (44100, 2, ["-B64", "-M0"], dict(bitrate=64, mode=0)),
               (44100, 2, ["-B64", "-S1"], dict(bitrate=64, filter_select=1))]


@needs_ref
def test_fuzz_against_the_reference_cli(tmp_path):
    """Random stream lengths (tiny, ragged, around multiples of 1152), sample types and option sets: the audio frames
    of the file the reference CLI writes equal the host build of the kernel bodies."""
    rng = np.random.default_rng(2024)
    for it in range(26):
        sr, nch, opts, kw = FUZZ_COMBOS[it % len(FUZZ_COMBOS)]
        n = int([rng.integers(1, 3000), rng.integers(3000, 60000), 1152 * rng.integers(1, 40) + rng.integers(-3, 4)][it % 3])
        kind = wavutil.KINDS[it % len(wavutil.KINDS)]
        samples = wavutil.make_samples(synth_pcm(300 + it, n / sr + 0.1, sr, nch)[:n], kind, seed=it)
        _, whole = ref_cli_file(tmp_path, samples, kind, sr, nch, opts, "fz")
        src = samples if kind == "s16" else wavutil.to_encoder_float(samples, kind)
        got, _, _ = simmod.encode_clip(capi.control(samprate=sr, nch=nch, **kw), src, tail=wavutil.tail_value(kind))
        head = whole.size - got.size
        assert head > 0 and np.array_equal(whole[head:], got), (sr, nch, opts, n, kind)


@pytest.mark.gpu
@needs_ref
def test_gpu_cli_fuzz_against_the_reference_cli(tmp_path):
    """Whole files (Info frame included) for random lengths and sample types, one GPU batch per option set."""
    rng = np.random.default_rng(77)
    for c, (sr, nch, opts, kw) in enumerate([FUZZ_COMBOS[0], FUZZ_COMBOS[2], FUZZ_COMBOS[3], FUZZ_COMBOS[4], FUZZ_COMBOS[12]]):
        lst = str(tmp_path / ("fz%d.txt" % c))
        wants = []
        with open(lst, "w") as f:
            for it in range(6):
                n = int([rng.integers(1, 3000), rng.integers(3000, 120000), 1152 * rng.integers(1, 60) + rng.integers(-3, 4)][it % 3])
                kind = wavutil.KINDS[(it + c) % len(wavutil.KINDS)]
                samples = wavutil.make_samples(synth_pcm(500 + 10 * c + it, n / sr + 0.1, sr, nch)[:n], kind, seed=it)
                name = "g%d_%d" % (c, it)
                wav, want = ref_cli_file(tmp_path, samples, kind, sr, nch, opts, name)
                wants.append((name, want))
                f.write("%s %s\n" % (wav, tmp_path / (name + "_gpu.mp3")))
        subprocess.run([CLI, "-@", lst] + opts, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
        for name, want in wants:
            got = np.fromfile(str(tmp_path / (name + "_gpu.mp3")), dtype=np.uint8)
            assert got.size == want.size and np.array_equal(got, want), (name, opts)


UP2_CASES = [(1, ["-B24"], dict(bitrate=24), False), (2, [], dict(), False), (2, ["-M3", "-B32"], dict(bitrate=32), True)]


@needs_ref
@pytest.mark.parametrize("sr", [8000, 11025, 12000])
def test_up_conversion_1_to_2_reproduces_reference_cli_audio(tmp_path, sr):
    """8 / 11.025 / 12 kHz input: the reference doubles the rate (Csrc case 1: linear interpolation, integer for mono
    sources) and limits the coded sub-bands (mp3enc.cpp:2700-2714, 2765-2787; srccf.cpp:80-100, 258-276, 472-492)."""
    for nch, opts, kw, to_mono in UP2_CASES:
        for kind, n in [("s16", 20000), ("u8", 1152 * 9 + 575), ("s24", 1152 * 9 + 576), ("f32", 577)]:
            samples = wavutil.make_samples(synth_pcm(81, 3.0, sr, nch)[:n], kind, seed=1)
            _, whole = ref_cli_file(tmp_path, samples, kind, sr, nch, opts)
            f = wavutil.to_encoder_float(samples, kind).reshape(-1, nch)
            y = wavutil.upsample2(f, wavutil.tail_value(kind), to_mono=to_mono)
            ec = capi.control(samprate=2 * sr, nch=y.shape[1], nsb_limit=wavutil.up2_nsb_limit(sr), **kw)
            got, _, _ = simmod.encode_clip(ec, y, tail=wavutil.tail_value(kind))
            head = whole.size - got.size
            assert head > 0 and np.array_equal(whole[head:], got), (sr, nch, opts, kind, n)


@pytest.mark.gpu
@needs_ref
def test_cli_identity_for_up_converted_input(tmp_path):
    for sr in (8000, 11025, 12000):
        for c, (nch, opts, kw, to_mono) in enumerate(UP2_CASES):
            for kind, n in [("s16", 30000), ("u8", 1152 * 9 + 575), ("s24", 1152 * 9 + 576)]:
                samples = wavutil.make_samples(synth_pcm(82, 4.0, sr, nch)[:n], kind, seed=2)
                name = "up_%d_%d_%s" % (sr, c, kind)
                wav, want = ref_cli_file(tmp_path, samples, kind, sr, nch, opts, name)
                out = str(tmp_path / (name + "_gpu.mp3"))
                subprocess.run([CLI, wav, out] + opts, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
                got = np.fromfile(out, dtype=np.uint8)
                assert got.size == want.size and np.array_equal(got, want), name
    # -A<n> (mpeg_select of MP3_audio_encode_init): values that resolve to the source rate or to its 1:2 up-conversion
    # give the reference's file (other targets: test_cli_identity_with_general_rate_conversion); no MPEG rate is refused
    for sr, nch, opts in [(22050, 2, ["-A1"]), (24000, 2, ["-A48000"]), (44100, 2, ["-A1"]),
                          (44100, 2, ["-A44100", "-B64"]), (22050, 2, ["-A2"]), (11025, 1, ["-A2"]), (32000, 2, ["-A0"])]:
        samples = wavutil.make_samples(synth_pcm(84, 2.0, sr, nch)[:30000], "s16", seed=3)
        name = "asel_%d_%d_%s" % (sr, nch, "".join(opts).replace("-", ""))
        wav, want = ref_cli_file(tmp_path, samples, "s16", sr, nch, opts, name)
        out = str(tmp_path / (name + "_gpu.mp3"))
        subprocess.run([CLI, wav, out] + opts, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
        got = np.fromfile(out, dtype=np.uint8)
        assert got.size == want.size and np.array_equal(got, want), name
    # no MPEG rate; below the range; rates the reference's converter has no filter for (its init fails as well)
    for sr, opts in [(44100, ["-A12345"]), (3000, []), (44100, ["-A8000"]), (22255, []), (33075, [])]:
        samples = wavutil.make_samples(synth_pcm(85, 0.5, sr, 2), "s16")
        wav = str(tmp_path / ("asel_bad_%d%s.wav" % (sr, "".join(opts))))
        wavutil.write_wav(wav, samples, "s16", sr, 2)
        outp = wav[:-4] + ".mp3"
        r = subprocess.run([CLI, wav, outp] + opts, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        assert r.returncode != 0 and not os.path.exists(outp), (sr, opts)


RESAMPLE_CLI_CASES = [(37800, 2, ["-B64"], "s16"), (47250, 2, [], "s16"), (44100, 2, ["-A2"], "s16"), (11025, 1, ["-A1", "-B48"], "s16"),
                      (44100, 2, ["-A32000", "-B96"], "s24"), (48000, 2, ["-A22050", "-M3", "-B40"], "f32"),
                      (32000, 1, ["-A44100"], "u8"), (16000, 2, ["-A1", "-B48"], "s16"), (16000, 1, ["-A32000", "-B48"], "s16"),
                      (44100, 1, ["-A24000", "-B32"], "s32"), (44000, 2, [], "s16"), (30000, 1, ["-B48"], "s24")]


@pytest.mark.gpu
@needs_ref
def test_cli_identity_with_general_rate_conversion(tmp_path):
    """Source rates that are no MPEG rate, and -A targets that need more than 1:2 (Csrc cases 2-4: up by m:n, down
    through a polyphase FIR, down in two stages), for every sample type; one batch list mixing such files with native-
    rate ones.  The reference is oracle/_ref/hmp3_zi (the unmodified sources, automatic variables zero-initialised): with
    the converter running inside the encode call the plain build reads the converter's leftovers through its one
    uninitialised local (spdsmr.c:193, 283; tests/test_resample.py, tests/test_gpu_boundary.py)."""
    zi = os.path.join(os.path.dirname(REF_BIN), "hmp3_zi")
    if not os.path.exists(zi):
        pytest.skip("oracle/_ref/hmp3_zi not built")
    for k, (sr, nch, opts, kind) in enumerate(RESAMPLE_CLI_CASES):
        samples = wavutil.make_samples(synth_pcm(90 + k, 2.5, sr, nch)[:int(2.2 * sr) + 7 * k], kind, seed=4)
        name = "rs%d" % k
        wav, want = ref_cli_file(tmp_path, samples, kind, sr, nch, opts, name, binary=zi)
        out = str(tmp_path / (name + "_gpu.mp3"))
        subprocess.run([CLI, wav, out] + opts, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
        got = np.fromfile(out, dtype=np.uint8)
        assert want.size > 2000 and got.size == want.size and np.array_equal(got, want), (sr, nch, opts, kind)
    # a list: two files that go through the converter between two that do not
    lst, wants = str(tmp_path / "mixed.txt"), []
    with open(lst, "w") as f:
        for k, sr in enumerate([44100, 37800, 32000, 47250]):
            samples = wavutil.make_samples(synth_pcm(120 + k, 1.5, sr, 2), "s16")
            name = "mx%d" % k
            wav, want = ref_cli_file(tmp_path, samples, "s16", sr, 2, ["-B64"], name, binary=zi)
            wants.append((name, want))
            f.write("%s %s\n" % (wav, tmp_path / (name + "_gpu.mp3")))
    subprocess.run([CLI, "-@", lst, "-B64"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    for name, want in wants:
        got = np.fromfile(str(tmp_path / (name + "_gpu.mp3")), dtype=np.uint8)
        assert got.size == want.size and np.array_equal(got, want), name


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("nch,to_mono,kw", [(1, False, dict(bitrate=24)), (2, False, dict()), (2, True, dict(bitrate=32))])
def test_handle_up_conversion(nch, to_mono, kw):
    """MP3_audio_encode_init with an 11.025 kHz source: 577 sample frames buffered per call, 576 consumed, encoded at
    22.05 kHz; the same bytes as the (CPU-validated) restatement of the reference's up-conversion gives."""
    sr, ncalls = 11025, 30
    samples = wavutil.make_samples(synth_pcm(91, 2.0, sr, nch)[:576 * ncalls], "s16")
    enc = capi.Encoder(capacity_seconds=10)
    ec = capi.control(samprate=sr, nch=nch, **kw)
    assert enc.init_mp3(ec, mono_convert=1 if to_mono else 0) == 577 * nch * 2
    padded = np.zeros((576 * (ncalls + 12) + 1, nch), np.int16)
    padded[:samples.shape[0]] = samples
    out = []
    for c in range(ncalls + 10):
        used, b = enc.encode_mp3(padded[576 * c:576 * c + 577])
        assert used == 576 * nch * 2
        out.append(b)
    got = np.concatenate(out)
    enc.close()
    y = wavutil.upsample2(samples.astype(np.float32).reshape(-1, nch), 0.0, to_mono=to_mono)
    want, _, _ = simmod.encode_clip(capi.control(samprate=2 * sr, nch=y.shape[1], nsb_limit=wavutil.up2_nsb_limit(sr), **kw), y)
    m = min(got.size, want.size)
    assert m > 0.85 * want.size and np.array_equal(got[:m], want[:m])


def control_from_options(opts, sr, nch):
    """The control block the CLI builds: hmp3_control_defaults + hmp3_control_apply_option per argument, then the
    channel / rate fix-ups of ff_encode (tomp3.cpp:357-566, 809-815)."""
    import ctypes as C
    L = capi.lib()
    L.hmp3_control_apply_option.argtypes = [C.c_void_p, C.c_char_p]
    ec = np.zeros(len(capi.EC_FIELDS), np.int32)
    L.hmp3_control_defaults(capi.vp(ec))
    for o in opts:
        assert L.hmp3_control_apply_option(capi.vp(ec), o.encode()) == 0, o
    f = capi.EC_FIELDS
    if ec[f.index("mode")] < 0:
        ec[f.index("mode")] = 0
    if nch == 1:
        ec[f.index("mode")] = 3
    elif ec[f.index("mode")] == 3:
        ec[f.index("mode")] = 1
    ec[f.index("samprate")] = sr
    return ec


OPTION_SETS = [["-B64", "-C1", "-O0"], ["-V75", "-L96"], ["-F16000", "-B56"], ["-Q1"], ["-T1", "-B80"], ["-U2"], ["-V20"],
               ["-V120", "-HF1"], ["-B112", "-M1"], ["-B64", "-M0"], ["-SBT20", "-B64"], ["-TX5"], ["-V60", "-F14000"],
               ["-B128"], ["-B48", "-S1"], ["-B40"]]


@needs_ref
def test_option_strings_through_the_parser_against_the_reference_cli(tmp_path):
    """Command-line arguments -> hmp3_control_apply_option -> encode (host build) == the reference CLI given the same
    arguments; a control the reference refuses (e.g. -B40 at 44.1 kHz) is refused too."""
    for it, opts in enumerate(OPTION_SETS):
        for sr, nch in [(44100, 2), (32000, 1), (24000, 2)]:
            pcm = synth_pcm(600 + it, 1.5, sr, nch)
            wav, mp3 = str(tmp_path / "o.wav"), str(tmp_path / "o.mp3")
            if os.path.exists(mp3):
                os.remove(mp3)
            wavutil.write_wav(wav, pcm, "s16", sr, nch)
            subprocess.run([REF_BIN, wav, mp3] + opts, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            ref_ok = os.path.exists(mp3) and os.path.getsize(mp3) > 0
            try:
                got, _, _ = simmod.encode_clip(control_from_options(opts, sr, nch), pcm)
            except Exception:
                got = None
            assert (got is not None) == ref_ok, (opts, sr, nch)
            if got is not None:
                whole = np.fromfile(mp3, dtype=np.uint8)
                head = whole.size - got.size
                assert head >= 0 and np.array_equal(whole[head:], got), (opts, sr, nch)


def test_cli_passes_over_unknown_option_letters_like_the_reference():
    """tomp3.cpp's option switch has no default: a letter that is no option is ignored, -h prints the usage text."""
    if not os.path.exists(CLI):
        pytest.skip("CLI not built")
    r = subprocess.run([CLI, "in.wav", "out.mp3", "-Y9", "-k"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert "Usage" not in r.stdout            # (goes on to open the input, or to report that there is no device)
    r = subprocess.run([CLI, "in.wav", "out.mp3", "-h"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert "Usage" in r.stdout and r.returncode == 0
