"""WAV writers for every sample type the reference CLI accepts (hmp3/src/pcmhpm.c, test/tomp3.cpp:722-760) and the
sample conversion Csrc::sr_convert applies to each (hmp3/src/srcc.cpp:804-834) restated in numpy for the tests."""
import struct

import numpy as np

KINDS = ("u8", "s16", "s24", "s32", "f32")


def make_samples(pcm_i16, kind, seed=0):
    """From int16 PCM make native-typed samples of `kind` that use the type's extra precision."""
    rng = np.random.default_rng(seed)
    p = np.asarray(pcm_i16, dtype=np.int64)
    if kind == "u8":
        return ((p >> 8) + 128).astype(np.uint8)
    if kind == "s16":
        return p.astype(np.int16)
    if kind == "s24":
        return (p * 256 + rng.integers(0, 256, size=p.shape)).astype(np.int32)         # 24-bit range in an int32
    if kind == "s32":
        return (p * 65536 + rng.integers(0, 65536, size=p.shape)).astype(np.int32)
    if kind == "f32":
        return (p.astype(np.float32) / np.float32(32768.0) +
                rng.uniform(-1e-5, 1e-5, size=p.shape).astype(np.float32)).astype(np.float32)
    raise ValueError(kind)


def to_encoder_float(samples, kind):
    """What the encoder core is fed for each source type (float PCM on the +-32768 scale)."""
    if kind == "u8":
        return (samples.astype(np.float32) - np.float32(128.0)) * np.float32(256.0)
    if kind == "s16":
        return samples.astype(np.float32)
    if kind == "s24":
        return samples.astype(np.float32) / np.float32(256.0)
    if kind == "s32":
        return samples.astype(np.float32) / np.float32(65536.0)        # int -> float first, as the C expression does
    if kind == "f32":
        return samples.astype(np.float32) * np.float32(32768.0)
    raise ValueError(kind)


def tail_value(kind):
    """What the CLI's zero flush BYTES decode to: silence, except for 8-bit (unsigned) sources."""
    return -32768.0 if kind == "u8" else 0.0


def raw_bytes(samples, kind):
    if kind == "s24":
        b = np.ascontiguousarray(samples.astype("<i4")).view(np.uint8).reshape(-1, 4)[:, :3]
        return np.ascontiguousarray(b).tobytes()
    dt = {"u8": "u1", "s16": "<i2", "s32": "<i4", "f32": "<f4"}[kind]
    return np.ascontiguousarray(samples.astype(dt)).tobytes()


def write_wav(path, samples, kind, sr, nch):
    bits = {"u8": 8, "s16": 16, "s24": 24, "s32": 32, "f32": 32}[kind]
    tag = 3 if kind == "f32" else 1
    data = raw_bytes(samples, kind)
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " +
                struct.pack("<IHHIIHH", 16, tag, nch, sr, sr * nch * bits // 8, nch * bits // 8, bits) +
                b"data" + struct.pack("<I", len(data)))
        f.write(data)


def downmix(f):
    """Csrc::src_filter_to_mono_case0 (hmp3/src/srccf.cpp:458-468): (float)((L + R) * 0.5), the sum taken in float."""
    return ((f[:, 0] + f[:, 1]).astype(np.float32).astype(np.float64) * 0.5).astype(np.float32).reshape(-1, 1)


def upsample2(f, pad, to_mono=False):
    """Csrc 1:2 up-conversion (hmp3/src/srccf.cpp:80-100 mono, 258-276 stereo, 472-492 stereo -> mono) of float samples
    f (n, ch); `pad` is the value of the samples after the end.  Returns (2 n + 3, ch') samples: the reference's main
    loop makes as many calls as an encoder-rate stream of 2 n + 3 samples would (see calls_for in pipeline.cu)."""
    n, ch = f.shape
    x = np.concatenate([f, np.full((1, ch), pad, np.float32)]).astype(np.float32)
    if ch == 1:
        a = np.trunc(x[:, 0]).astype(np.int64)                   # int a = x[i]
        y = np.full((2 * n + 3, 1), pad, np.float32)
        y[0:2 * n:2, 0] = a[:n].astype(np.float32)
        y[1:2 * n:2, 0] = ((a[:n] + a[1:]) >> 1).astype(np.float32)
        return y
    if to_mono:
        s = (x[:, 0] + x[:, 1]).astype(np.float32)
        y = np.full((2 * n + 3, 1), pad, np.float32)
        y[0:2 * n:2, 0] = (s[:n].astype(np.float64) * 0.5).astype(np.float32)
        y[1:2 * n:2, 0] = ((s[:n] + s[1:]).astype(np.float32).astype(np.float64) * 0.25).astype(np.float32)
        return y
    y = np.full((2 * n + 3, 2), pad, np.float32)
    y[0:2 * n:2] = x[:n]
    y[1:2 * n:2] = ((x[:n] + x[1:]).astype(np.float32).astype(np.float64) * 0.5).astype(np.float32)
    return y


def up2_nsb_limit(source_rate):
    """Sub-band limit CMp3Enc::MP3_audio_encode_init sets for an up-converted source (mp3enc.cpp:2765-2787)."""
    cutoff = int(np.float32(0.90) * np.float32(source_rate) / np.float32(2))
    return min(30, (64 * cutoff + source_rate) // (2 * source_rate))
