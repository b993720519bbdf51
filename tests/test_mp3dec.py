"""The test-only Layer III decoder (tests/mp3dec.py) is itself pinned: it must reconstruct what the reference encoder
coded -- tonal material to > 40 dB, with long blocks, with nearly all granules forced short, mono and M/S stereo --
before its SNR figures are used to judge a non-bit-exact encode mode."""
import numpy as np
import pytest

import mp3dec
import refmod

pytestmark = pytest.mark.skipif(not refmod.available(), reason="oracle/_ref not built")


def _tones(n, nch):
    t = np.arange(n) / 44100.0
    x = 8000 * np.sin(2 * np.pi * 1000 * t) * (1 + 0.5 * np.sin(2 * np.pi * 3 * t)) + 3000 * np.sin(2 * np.pi * 5013 * t) \
        + 2000 * np.sin(2 * np.pi * 237 * t)
    ch = [x, 0.5 * x + 2000 * np.sin(2 * np.pi * 3001 * t)][:nch]
    return np.stack(ch, axis=1).astype(np.int16)


@pytest.mark.parametrize("nch,sbt,floor", [(1, 700, 45.0), (2, 700, 45.0), (2, 1, 35.0)])
def test_decoder_reconstructs_reference_encodes(nch, sbt, floor):
    pcm = _tones(44100, nch)
    mp3, tr = refmod.ref_encode_clip(refmod.make_ec(samprate=44100, nch=nch, bitrate=96, short_block_threshold=sbt), pcm,
                                     max_trace_calls=64)
    if sbt == 1:
        assert (tr["g"]["block_type"] == 2).sum() > 60        # the short-block path really is exercised
    snr, lag = mp3dec.snr_db(pcm, mp3dec.decode(mp3))
    assert snr > floor, snr


def test_analysis_window_is_the_iso_prototype():
    c = mp3dec.analysis_window()
    assert abs(c[256] - 0.035780907) < 1e-7 and abs(c[1] + 0.000000477) < 1e-8 and c[0] == 0.0
    for n in range(1, 256):                      # C[n] = -C[512 - n], except where both sit on a 64-sample boundary
        want = c[512 - n] if n % 64 == 0 else -c[512 - n]
        assert abs(c[n] - want) < 2e-6, n
