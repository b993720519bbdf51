"""CPU suite: the handle's sample-rate converter (hmp3_b200/csrc/resample.h, Csrc cases 2-4) against the reference's
Csrc (oracle/_ref, srcc.cpp / srccf.cpp) on the same PCM: same buffering requirement, same source frames consumed by
every call, same output floats bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import refmod
from hmp3_b200 import capi
from hmp3_b200.synth import synth_pcm
from wavutil import write_wav

needs_ref = pytest.mark.skipif(not refmod.available(), reason="oracle/_ref not built")


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def ref_convert(source, channels, target, target_channels, pcm_i16, ncalls):
    L = refmod.lib()
    L.ref_src_convert.restype = C.c_int
    L.ref_src_convert.argtypes = [C.c_int] * 6 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    tch = min(channels, target_channels)
    out = np.zeros((ncalls, 1152, tch), np.float32)
    used = np.zeros(ncalls, np.int32)
    cut = C.c_int(0)
    buf = np.ascontiguousarray(pcm_i16, np.int16)
    r = L.ref_src_convert(source, channels, 16, 0, target, target_channels, vp(buf), ncalls, vp(out), vp(used), C.byref(cut))
    return r, out, used, cut.value


def our_convert(source, target, layout, pcm_i16, ncalls):
    L = capi.lib()
    L.hmp3_debug_resample.restype = C.c_int
    L.hmp3_debug_resample.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    x = np.ascontiguousarray(pcm_i16, np.int16).astype(np.float32)
    ow = 2 if layout == 1 else 1
    y = np.zeros((ncalls, 1152, ow), np.float32)
    used = np.zeros(ncalls, np.int32)
    r = L.hmp3_debug_resample(source, target, layout, vp(x), ncalls, vp(y), vp(used))
    return r, y, used


# (source, target): up by m:n (case 2), down with few phases (case 3), two-stage (case 4)
PAIRS = [(32000, 44100), (32000, 48000), (24000, 44100), (8000, 32000), (11025, 32000), (12000, 44100), (16000, 22050), (22050, 24000), (37800, 32000), (44100, 22050),
         (48000, 32000), (48000, 44100), (44100, 32000), (32000, 24000), (44100, 16000), (47250, 44100), (32000, 22050),
         (22255, 22050), (33075, 32000), (48000, 22050)]


@needs_ref
@pytest.mark.parametrize("source,target", PAIRS)
@pytest.mark.parametrize("layout", [0, 1, 2])
def test_converter_matches_reference(source, target, layout):
    channels = 1 if layout == 0 else 2
    target_channels = 2 if layout == 1 else 1
    ncalls = 12
    need = int(ncalls * 1152 * source / target) + 8000       # frames: what the calls consume + what a call may read ahead
    pcm = synth_pcm(500 + layout, need / source + 0.1, source, channels)[:need]
    pad = np.zeros((4096, channels), np.int16)               # the reference converts 1152+ frames per call whatever it uses
    pcm = np.concatenate([pcm, pad])
    r_ref, y_ref, u_ref, _ = ref_convert(source, channels, target, target_channels, pcm, ncalls)
    r_our, y_our, u_our = our_convert(source, target, layout, pcm, ncalls)
    if r_ref <= 0:                                             # a pair the reference cannot set up: refused here too
        assert r_our <= 0, (source, target, r_our)
        return
    assert r_our * channels * 2 == r_ref                       # bytes to buffer = frames x channels x 2 (16-bit)
    assert np.array_equal(u_our * channels * 2, u_ref)         # source consumed by every call
    assert np.array_equal(y_our.view(np.uint32), y_ref.reshape(y_our.shape).view(np.uint32))


@needs_ref
def test_refusals_agree():
    """Rate pairs the reference's converter cannot set up (coefficient table, no two-stage factoring) are refused."""
    for source, target in [(44100, 8000), (47999, 44100), (48000, 11025), (44101, 44100)]:
        pcm = np.zeros((20000, 1), np.int16)
        r_ref, _, _, _ = ref_convert(source, 1, target, 1, pcm, 1) if target >= 5000 else (0, 0, 0, 0)
        L = capi.lib()
        L.hmp3_debug_resample.restype = C.c_int
        L.hmp3_debug_resample.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        x = np.zeros(40000, np.float32)
        y = np.zeros(1152 * 2, np.float32)
        u = np.zeros(1, np.int32)
        r_our = L.hmp3_debug_resample(source, target, 0, vp(x), 1, vp(y), vp(u))
        assert (r_ref > 0) == (r_our > 0), (source, target, r_ref, r_our)


REFDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref")
CLI_CASES = [(["-B64"], 44100, 2), ([], 44100, 2), (["-B32"], 22050, 1), (["-B48", "-X0"], 32000, 2), ([], 16000, 2),
             (["-B64"], 37800, 2), ([], 47250, 2), (["-B48", "-A1"], 8000, 1), (["-A2"], 44100, 2),
             (["-B64", "-A32000"], 48000, 2), (["-A44100"], 32000, 2), (["-B48", "-A1"], 22050, 1)]


@pytest.mark.parametrize("opts,sr,nch", CLI_CASES)
def test_zero_initialised_reference_build_is_the_reference(tmp_path, opts, sr, nch):
    """oracle/_ref/hmp3_zi (unmodified sources, -ftrivial-auto-var-init=zero) is what the rate-conversion parity tests
    compare against; it writes the file the plain build writes at native rates and for every conversion whose output
    the plain build defines."""
    plain, zi = os.path.join(REFDIR, "hmp3"), os.path.join(REFDIR, "hmp3_zi")
    if not (os.path.exists(plain) and os.path.exists(zi)):
        pytest.skip("oracle/_ref not built")
    wav = str(tmp_path / "in.wav")
    write_wav(wav, synth_pcm(6200 + sr // 1000, 2.0, sr, nch), "s16", sr, nch)
    a, b = str(tmp_path / "a.mp3"), str(tmp_path / "b.mp3")
    subprocess.run([plain, wav, a] + opts, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    subprocess.run([zi, wav, b] + opts, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    ga = open(a, "rb").read()
    assert len(ga) > 2000 and ga == open(b, "rb").read()


@needs_ref
@pytest.mark.parametrize("source,target,layout", [(32000, 44100, 0), (37800, 44100, 1), (48000, 22050, 2), (44100, 32000, 1),
                                                  (11025, 32000, 0), (47250, 48000, 1)])
def test_converter_matches_reference_over_a_long_stream(source, target, layout):
    """400 calls (about ten seconds at the encode rate): every phase counter wraps many times."""
    channels = 1 if layout == 0 else 2
    target_channels = 2 if layout == 1 else 1
    ncalls = 400
    need = int(ncalls * 1152 * source / target) + 8000
    base = synth_pcm(700 + layout, 4.0, source, channels)
    pcm = np.concatenate([base] * (need // base.shape[0] + 1))[:need]
    pcm = np.concatenate([pcm, np.zeros((4096, channels), np.int16)])
    r_ref, y_ref, u_ref, _ = ref_convert(source, channels, target, target_channels, pcm, ncalls)
    r_our, y_our, u_our = our_convert(source, target, layout, pcm, ncalls)
    assert r_ref > 0 and r_our * channels * 2 == r_ref
    assert np.array_equal(u_our * channels * 2, u_ref)
    assert np.array_equal(y_our.view(np.uint32), y_ref.reshape(y_our.shape).view(np.uint32))


def test_zero_initialised_reference_build_over_the_option_sets(tmp_path):
    """... and for the option strings of the CLI tests at three native rates (48 encodes with each build)."""
    from test_wav_formats import OPTION_SETS
    plain, zi = os.path.join(REFDIR, "hmp3"), os.path.join(REFDIR, "hmp3_zi")
    if not (os.path.exists(plain) and os.path.exists(zi)):
        pytest.skip("oracle/_ref not built")
    wav, a, b = str(tmp_path / "in.wav"), str(tmp_path / "a.mp3"), str(tmp_path / "b.mp3")
    for it, opts in enumerate(OPTION_SETS):
        for sr, nch in [(44100, 2), (32000, 1), (24000, 2)]:
            write_wav(wav, synth_pcm(900 + it, 1.5, sr, nch), "s16", sr, nch)
            for f in (a, b):
                if os.path.exists(f):
                    os.remove(f)
            subprocess.run([plain, wav, a] + opts, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
            subprocess.run([zi, wav, b] + opts, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
            if not os.path.exists(a):                    # a combination the reference's init refuses
                assert not os.path.exists(b), (opts, sr, nch)
                continue
            assert open(a, "rb").read() == open(b, "rb").read(), (opts, sr, nch)


@needs_ref
@pytest.mark.parametrize("sr,nch,opts,kw,target,to_mono", [
    (37800, 2, ["-B64"], dict(bitrate=64), 32000, False), (37800, 2, ["-A44100"], dict(), 44100, False), (48000, 2, ["-A32000", "-B64"], dict(bitrate=64), 32000, False),
    (44100, 1, ["-A24000", "-B32"], dict(bitrate=32), 24000, False), (32000, 1, ["-A44100", "-B64"], dict(bitrate=64), 44100, False),
    (48000, 2, ["-A22050", "-M3", "-B40"], dict(bitrate=40), 22050, True), (44100, 2, ["-A2"], dict(), 22050, False),
    (47250, 2, [], dict(), 48000, False)])
def test_converter_and_host_encoder_reproduce_reference_cli_audio(tmp_path, sr, nch, opts, kw, target, to_mono):
    """The whole MP3_audio_encode path on the CPU: the handle's converter, then the host build of the kernel bodies at
    the encode rate (with the sub-band limit of an up-converted signal, mp3enc.cpp:2765-2787), against the audio frames
    the reference CLI (hmp3_zi) writes for the same WAV -- the composition the GPU tests check through the C ABI."""
    import simmod
    zi = os.path.join(REFDIR, "hmp3_zi")
    if not (os.path.exists(zi) and simmod.available()):
        pytest.skip("oracle/_ref/hmp3_zi or the host build missing")
    N = int(2.3 * sr)
    pcm = synth_pcm(1300 + sr // 100, 2.5, sr, nch)[:N]
    wav, mp3 = str(tmp_path / "in.wav"), str(tmp_path / "ref.mp3")
    write_wav(wav, pcm, "s16", sr, nch)
    subprocess.run([zi, wav, mp3] + opts, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, check=True)
    whole = np.fromfile(mp3, dtype=np.uint8)
    # the CLI's main loop (tomp3.cpp:908-942): 4 x bytes_in_init zero bytes after the data, a call while bytes_in_init
    # bytes are buffered
    layout = 0 if nch == 1 else (2 if to_mono else 1)
    padded = np.concatenate([pcm.reshape(N, nch), np.zeros((40000, nch), np.int16)])
    ncalls = int(N * target / sr / 1152) + 8
    minfr, y, used = our_convert(sr, target, layout, padded, ncalls)
    assert minfr > 0
    pos = np.concatenate([[0], np.cumsum(used)])
    calls = int(np.sum(N + 4 * minfr - pos[:ncalls] >= minfr))
    assert 0 < calls < ncalls
    y = y[:calls].reshape(calls * 1152, -1)
    # the same calls from the host encoder's own loop: a stream of n samples makes (n + 3 * 1153 + 1152) // 1152 calls
    # and everything behind n is the zero tail, so n must lie behind the last sample that is not zero
    n = 1152 * (calls - 1) - 3 * 1153
    nz = np.nonzero(np.any(y != 0, axis=1))[0]
    extent = int(nz[-1]) + 1 if nz.size else 0
    n = max(n, extent)
    assert (n + 3 * 1153 + 1152) // 1152 == calls, "clip length lands on a call boundary: pick another"
    ctl = dict(kw)
    if sr < target:
        ctl["nsb_limit"] = min(30, (64 * int(0.90 * sr / 2) + target // 2) // target)
    ec = capi.control(samprate=target, nch=y.shape[1], **ctl)
    got, _, _ = simmod.encode_clip(ec, np.ascontiguousarray(y[:n]), tail=0.0)
    head = whole.size - got.size
    assert head > 0 and np.array_equal(whole[head:], got), (sr, target, opts)
