"""The Xing/Info frame (host post-pass) against files written by the reference CLI itself: the frame
is rebuilt from the reference's own audio frames and per-call output log and must equal the file's first frame."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import refmod
from configs import CONFIGS
from hmp3_b200 import capi
from hmp3_b200.synth import synth_pcm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (write_wav)

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "hmp3")
pytestmark = pytest.mark.skipif(not (os.path.exists(REF_BIN) and refmod.available()), reason="oracle/_ref not built")

CLI_OPTS = {"c1_cbr128_44k": ["-B64"], "c2_vbr50_44k": [], "c3_v100_hf2_48k": ["-V100", "-HF2", "-F19000"],
            "c4a_cbr32_22k_mono": ["-B32"], "c4b_vbr50_32k": []}


def ref_cli_encode(pcm, sr, nch, opts):
    d = tempfile.mkdtemp(prefix="hmp3_tag_")
    wav, mp3 = os.path.join(d, "in.wav"), os.path.join(d, "out.mp3")
    data = np.ascontiguousarray(pcm, dtype="<i2").tobytes()
    import struct
    with open(wav, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " +
                struct.pack("<IHHIIHH", 16, 1, nch, sr, sr * nch * 2, nch * 2, 16) + b"data" + struct.pack("<I", len(data)))
        f.write(data)
    subprocess.run([REF_BIN, wav, mp3] + opts, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    out = np.fromfile(mp3, dtype=np.uint8)
    for p in (wav, mp3):
        os.remove(p)
    os.rmdir(d)
    return out


@pytest.mark.parametrize("name,seed,sr,nch,kw", CONFIGS)
@pytest.mark.parametrize("secs", [2.5, 31.0])
def test_info_frame_matches_reference_cli(name, seed, sr, nch, kw, secs):
    pcm = synth_pcm(seed + 3, secs, sr, nch)
    whole = ref_cli_encode(pcm, sr, nch, CLI_OPTS[name])
    audio, tr = refmod.ref_encode_clip(refmod.make_ec(samprate=sr, nch=nch, **kw), pcm, max_trace_calls=100000)
    head_bytes = whole.size - audio.size
    assert head_bytes > 0 and np.array_equal(whole[head_bytes:], audio)       # the CLI's audio frames = the harness'
    ncalls = (pcm.shape[0] + 3 * 1153 + 1152) // 1152                         # the CLI's main loop
    bytes_after = np.cumsum(tr["out_bytes"][:ncalls].astype(np.int64))
    # frames are emitted whole: count the frames that end at or before each cumulative byte count
    ends, p = [], 0
    br = ([0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320] if sr >= 32000 else
          [0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160])
    while p < audio.size:
        h2 = int(audio[p + 2])
        n = (144000 if sr >= 32000 else 72000) * br[h2 >> 4] // sr + ((h2 >> 1) & 1)
        p += n
        ends.append(p)
    ends = np.array(ends)
    frames_after = np.searchsorted(ends, bytes_after, side="right")
    got = capi.info_frame(capi.control(samprate=sr, nch=nch, **kw), nch, pcm.shape[0], audio, len(ends), frames_after,
                          bytes_after)
    assert got.size == head_bytes
    assert np.array_equal(got, whole[:head_bytes])


def _frame_ends(audio, sr):
    ends, p = [], 0
    br = ([0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320] if sr >= 32000 else
          [0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160])
    while p < audio.size:
        h2 = int(audio[p + 2])
        p += (144000 if sr >= 32000 else 72000) * br[h2 >> 4] // sr + ((h2 >> 1) & 1)
        ends.append(p)
    return np.array(ends)


@pytest.mark.parametrize("xflag", [1, 2, 3, 65, 66, 67])
def test_info_frame_for_every_x_option(xflag):
    """-X<n>: 1 = frame/byte counts only, 2 / 3 = with seek table, +64 = with the info tag (tomp3.cpp:682-689)."""
    for sr, nch, kw, opts in [(44100, 2, dict(bitrate=64), ["-B64"]), (44100, 2, dict(), []),
                              (22050, 1, dict(bitrate=32), ["-B32"])]:
        pcm = synth_pcm(5, 6.0, sr, nch)
        whole = ref_cli_encode(pcm, sr, nch, opts + ["-X%d" % xflag])
        audio, tr = refmod.ref_encode_clip(refmod.make_ec(samprate=sr, nch=nch, **kw), pcm, max_trace_calls=100000)
        head_bytes = whole.size - audio.size
        ncalls = (pcm.shape[0] + 3 * 1153 + 1152) // 1152
        bytes_after = np.cumsum(tr["out_bytes"][:ncalls].astype(np.int64))
        ends = _frame_ends(audio, sr)
        frames_after = np.searchsorted(ends, bytes_after, side="right")
        got = capi.info_frame(capi.control(samprate=sr, nch=nch, **kw), nch, pcm.shape[0], audio, len(ends), frames_after,
                              bytes_after, xing_flag=xflag)
        assert got.size == head_bytes and np.array_equal(got, whole[:head_bytes]), (sr, nch, opts, xflag)
