"""ctypes binding of the product library hmp3_b200/_lib/libhmp3_b200.so (C ABI: include/hmp3_b200.h).

There is no CPU fallback: loading fails loudly if the CUDA library has not been built, and every
encode entry fails if no CUDA device is present."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HMP3_B200_LIB", os.path.join(HERE, "_lib", "libhmp3_b200.so"))

EC_FIELDS = ["mode", "bitrate", "samprate", "nsbstereo", "filter_select", "freq_limit", "nsb_limit", "layer",
             "cr_bit", "original", "hf_flag", "vbr_flag", "vbr_mnr", "vbr_br_limit", "vbr_delta_mnr",
             "chan_add_f0", "chan_add_f1", "sparse_scale"] + ["mnr_adjust%d" % i for i in range(21)] + \
            ["cpu_select", "quick", "test1", "test2", "test3", "short_block_threshold"]

_lib = None


class Hmp3Error(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Hmp3Error("CUDA library %s is missing: run __graft_entry__.build() (no CPU fallback exists)"
                            % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.hmp3_get_last_error.restype = C.c_char_p
    return _lib


def last_error():
    return lib().hmp3_get_last_error().decode()


def vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def control(samprate=44100, nch=2, bitrate=-1, vbr_mnr=50, hf=0, freq_limit=24000, mode=None, **kw):
    """hmp3_control image with the CLI defaults and option semantics (-B => CBR, otherwise VBR)."""
    ec = dict.fromkeys(EC_FIELDS, 0)
    ec.update(mode=1, bitrate=bitrate, samprate=samprate, nsbstereo=-1, filter_select=-1, nsb_limit=-1,
              freq_limit=freq_limit, cr_bit=1, original=1, layer=3, hf_flag=(1 | hf) if hf else 0,
              vbr_flag=1 if bitrate < 0 else 0, vbr_mnr=vbr_mnr, vbr_br_limit=160, chan_add_f0=24000,
              chan_add_f1=24000, sparse_scale=-1, vbr_delta_mnr=0, cpu_select=0, quick=-1, test1=-1, test2=0,
              test3=0, short_block_threshold=700)
    if mode is not None:
        ec["mode"] = mode
    if nch == 1:
        ec["mode"] = 3
    elif ec["mode"] == 3:
        ec["mode"] = 1
    ec.update(kw)
    return np.array([ec[f] for f in EC_FIELDS], dtype=np.int32)


def debug_analysis(ec, pcm_i16, ngran, nch, device=0):
    pcm = np.ascontiguousarray(pcm_i16, dtype=np.int16)
    out = dict(sbt=np.zeros((ngran, nch, 576), np.float32), ginfo=np.zeros((ngran, 4), np.int32),
               xr=np.zeros((ngran, nch, 576), np.float32), raw=np.zeros((ngran, nch, 92), np.float32),
               ms_raw=np.zeros(ngran, np.int32), att=np.zeros((ngran, nch, 9), np.int32))
    f = lib().hmp3_debug_analysis
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int] + [C.c_void_p] * 6
    r = f(vp(ec), vp(pcm), pcm.shape[0], ngran, device, vp(out["sbt"]), vp(out["ginfo"]), vp(out["xr"]),
          vp(out["raw"]), vp(out["ms_raw"]), vp(out["att"]))
    if r != 0:
        raise Hmp3Error("hmp3_debug_analysis failed (%d): %s" % (r, last_error()))
    return out


def fp32_peak(device=0):
    """FP32 issue-rate microbenchmark (hmp3_debug_fp32_peak): TFLOP/s of FFMA chains and of the non-fused mix."""
    a, b = C.c_float(0), C.c_float(0)
    r = lib().hmp3_debug_fp32_peak(device, C.byref(a), C.byref(b))
    if r != 0:
        return {"ffma_tflops": None, "nonfused_tflops": None}
    return {"ffma_tflops": a.value, "nonfused_tflops": b.value}


PCM_S16, PCM_F32 = 0, 1        # hmp3_stream_desc.pcm_format / hmp3_batch_create_ex formats


class Batch:
    """Reusable batch plan (hmp3_batch_*): n streams of fixed control + length on one device."""

    def __init__(self, controls, num_samples, device=0, formats=None):
        """formats: per stream PCM_S16 (default) or PCM_F32 (float PCM on the +-32768 scale)."""
        L = lib()
        self.n = len(controls)
        self.ctl = np.ascontiguousarray(np.stack(controls).astype(np.int32))
        self.ns = np.ascontiguousarray(np.asarray(num_samples, dtype=np.int64))
        self.fmt = np.ascontiguousarray(np.zeros(self.n, np.int32) if formats is None else np.asarray(formats, np.int32))
        L.hmp3_batch_create_ex.restype = C.c_void_p
        L.hmp3_batch_create_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        self.h = L.hmp3_batch_create_ex(vp(self.ctl), vp(self.ns), vp(self.fmt), self.n, device)
        if not self.h:
            raise Hmp3Error("hmp3_batch_create failed: " + last_error())
        L.hmp3_batch_out_bound.restype = C.c_int64
        L.hmp3_batch_out_bound.argtypes = [C.c_void_p, C.c_int64]
        self.bound = np.array([L.hmp3_batch_out_bound(vp(self.ctl[i]), int(self.ns[i])) for i in range(self.n)],
                              dtype=np.int64)
        for name in ("hmp3_batch_destroy", "hmp3_batch_sync", "hmp3_batch_wait_uploads"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.hmp3_batch_run.argtypes = [C.c_void_p, C.c_int]
        L.hmp3_batch_set_timing.argtypes = [C.c_void_p, C.c_int]
        L.hmp3_batch_last_launches.argtypes = [C.c_void_p]
        L.hmp3_batch_last_run_ms.argtypes = [C.c_void_p]
        L.hmp3_batch_last_run_ms.restype = C.c_float
        L.hmp3_batch_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
        L.hmp3_batch_upload_f32.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
        L.hmp3_batch_results.argtypes = [C.c_void_p] * 5
        L.hmp3_batch_download_all.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.hmp3_batch_encode_host.argtypes = [C.c_void_p] * 7
        L.hmp3_batch_phase_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.hmp3_batch_device_pcm.restype = C.c_void_p
        L.hmp3_batch_device_pcm.argtypes = [C.c_void_p]

    def close(self):
        if self.h:
            lib().hmp3_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:       # interpreter shutdown: module globals may already be gone
            pass

    def _ck(self, r, what):
        if r != 0:
            raise Hmp3Error("%s failed (%d): %s" % (what, r, last_error()))

    def upload(self, i, pcm):
        if self.fmt[i] == PCM_F32:
            a = np.ascontiguousarray(pcm, dtype=np.float32)
            self._ck(lib().hmp3_batch_upload_f32(self.h, i, vp(a), a.shape[0]), "upload_f32")
        else:
            a = np.ascontiguousarray(pcm, dtype=np.int16)
            self._ck(lib().hmp3_batch_upload(self.h, i, vp(a), a.shape[0]), "upload")

    def upload_ptr(self, i, ptr, nsamples):
        self._ck(lib().hmp3_batch_upload(self.h, i, C.c_void_p(ptr), nsamples), "upload")

    def run(self, async_=False):
        self._ck(lib().hmp3_batch_run(self.h, 1 if async_ else 0), "run")

    def sync(self):
        self._ck(lib().hmp3_batch_sync(self.h), "sync")

    def sync_stream(self):
        self._ck(lib().hmp3_batch_wait_uploads(self.h), "wait_uploads")

    def set_rate_tap(self, stream, ngran):
        """Diagnostics: have every run copy what the serial stage hands the packing pass for `stream` (hmp3_debug_set_
        rate_tap).  Returns the structured array [ngran][2] that the runs fill: ix, sign, sf, gr per granule-channel."""
        L = lib()
        L.hmp3_debug_set_rate_tap.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong]
        dt = np.dtype([("ix", "<i2", 576), ("sign", "<u4", 18), ("sf", "u1", 64), ("gr", "<i4", 27)])
        assert dt.itemsize == L.hmp3_debug_rate_tap_record_bytes()
        self._tap = np.zeros((ngran, 2), dt)
        self._ck(L.hmp3_debug_set_rate_tap(self.h, stream, vp(self._tap), 2 * ngran), "set_rate_tap")
        return self._tap

    def set_timing(self, on):
        lib().hmp3_batch_set_timing(self.h, 1 if on else 0)

    def set_serialize(self, on):
        f = lib().hmp3_batch_set_serialize
        f.argtypes = [C.c_void_p, C.c_int]
        f(self.h, 1 if on else 0)

    def launches(self):
        return lib().hmp3_batch_last_launches(self.h)

    def chunk_granules(self):
        f = lib().hmp3_batch_chunk_granules
        f.argtypes = [C.c_void_p]
        return int(f(self.h))

    def last_run_ms(self):
        return float(lib().hmp3_batch_last_run_ms(self.h))

    def results(self):
        nb, nf = np.zeros(self.n, np.int64), np.zeros(self.n, np.int32)
        off, st = np.zeros(self.n, np.int64), np.zeros(self.n, np.int32)
        self._ck(lib().hmp3_batch_results(self.h, vp(nb), vp(nf), vp(off), vp(st)), "results")
        return nb, nf, off, st

    def download_all(self, out=None):
        nb, nf, off, st = self.results()
        total = int(nb.sum())
        if out is None:
            out = np.zeros(max(total, 1), np.uint8)
        tot = C.c_int64(0)
        self._ck(lib().hmp3_batch_download_all(self.h, vp(out), out.size, C.byref(tot)), "download_all")
        return out, off, nb, nf, st

    def phase_ms(self):
        names = (C.c_char_p * 16)()
        ms = (C.c_float * 16)()
        ln = (C.c_int * 16)()
        k = lib().hmp3_batch_phase_ms(self.h, names, ms, ln, 16)
        return {names[i].decode(): (ms[i], ln[i]) for i in range(k)}

    def encode_host_ptrs(self, pcm_ptrs, out_ptrs, out_caps):
        """pcm_ptrs/out_ptrs: uint64 arrays of host addresses; returns (out_bytes, out_frames, status)."""
        nb, nf, st = np.zeros(self.n, np.int64), np.zeros(self.n, np.int32), np.zeros(self.n, np.int32)
        self._ck(lib().hmp3_batch_encode_host(self.h, vp(pcm_ptrs), vp(out_ptrs), vp(out_caps), vp(nb), vp(nf),
                                              vp(st)), "encode_host")
        return nb, nf, st

    def encode_host(self, pcms):
        pcms = [np.ascontiguousarray(p, dtype=np.float32 if self.fmt[i] == PCM_F32 else np.int16)
                for i, p in enumerate(pcms)]
        outs = [np.zeros(int(b), np.uint8) for b in self.bound]
        pp = np.array([p.ctypes.data for p in pcms], dtype=np.uint64)
        op = np.array([o.ctypes.data for o in outs], dtype=np.uint64)
        nb, nf, st = self.encode_host_ptrs(pp, op, self.bound)
        for i in range(self.n):
            if st[i] != 0:
                raise Hmp3Error("stream %d failed with status %d" % (i, st[i]))
        return [outs[i][:nb[i]].copy() for i in range(self.n)], nf


def encode_batch(controls, pcms, device=0):
    """Encode a list of PCM arrays (nsamples, nch), int16 or float32 (+-32768 scale) -> list of MP3 byte arrays
    (no Xing/Info frame)."""
    fmts = [PCM_F32 if np.asarray(p).dtype == np.float32 else PCM_S16 for p in pcms]
    b = Batch(controls, [p.shape[0] for p in pcms], device, formats=fmts)
    try:
        outs, _ = b.encode_host(pcms)
    finally:
        b.close()
    return outs


class Encoder:
    """CMp3Enc-style handle (hmp3_encoder_*): init, then one call per 1152 samples per channel."""

    def __init__(self, device=0, capacity_seconds=None):
        L = lib()
        L.hmp3_encoder_new.restype = C.c_void_p
        L.hmp3_encoder_new.argtypes = [C.c_int]
        L.hmp3_encoder_delete.argtypes = [C.c_void_p]
        L.hmp3_encoder_set_capacity_seconds.argtypes = [C.c_void_p, C.c_int]
        L.hmp3_MP3_audio_encode_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.hmp3_L3_audio_encode_init.argtypes = [C.c_void_p, C.c_void_p]
        L.hmp3_MP3_audio_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hmp3_MP3_audio_encode.restype = C.c_uint64      # struct {int in_bytes; int out_bytes;} by value
        L.hmp3_L3_audio_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hmp3_L3_audio_encode.restype = C.c_uint64
        L.hmp3_L3_audio_encode_get_frames.argtypes = [C.c_void_p]
        L.hmp3_L3_audio_encode_get_frames.restype = C.c_uint
        L.hmp3_L3_audio_encode_get_bitrate_float.argtypes = [C.c_void_p]
        L.hmp3_L3_audio_encode_get_bitrate_float.restype = C.c_float
        L.hmp3_L3_audio_encode_info_string.argtypes = [C.c_void_p, C.c_char_p]
        L.hmp3_L3_audio_encode_info_ec.argtypes = [C.c_void_p, C.c_void_p]
        self.h = L.hmp3_encoder_new(device)
        if not self.h:
            raise Hmp3Error("hmp3_encoder_new failed: " + last_error())
        if capacity_seconds:
            L.hmp3_encoder_set_capacity_seconds(self.h, int(capacity_seconds))
        self.out = np.zeros(1 << 17, np.uint8)

    def close(self):
        if self.h:
            lib().hmp3_encoder_delete(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:       # interpreter shutdown
            pass

    def init_mp3(self, ec, source_bits=16, source_is_float=0, mpeg_select=0, mono_convert=0):
        return lib().hmp3_MP3_audio_encode_init(self.h, vp(ec), source_bits, source_is_float, mpeg_select, mono_convert)

    def init_l3(self, ec):
        return lib().hmp3_L3_audio_encode_init(self.h, vp(ec))

    @staticmethod
    def _io(v):
        return int(v & 0xFFFFFFFF), int(v >> 32)

    def encode_mp3(self, pcm, raw=False):
        """pcm: int16 samples, or (raw=True) the caller's bytes in the source format given to init_mp3."""
        a = np.ascontiguousarray(pcm) if raw else np.ascontiguousarray(pcm, dtype=np.int16)
        i, o = self._io(lib().hmp3_MP3_audio_encode(self.h, vp(a), vp(self.out)))
        return i, self.out[:o].copy()

    def encode_l3(self, pcm_f32):
        a = np.ascontiguousarray(pcm_f32, dtype=np.float32)
        i, o = self._io(lib().hmp3_L3_audio_encode(self.h, vp(a), vp(self.out)))
        return i, self.out[:o].copy()

    def encode_l3_packet(self, pcm_f32, want_bs=True):
        """hmp3_L3_audio_encode_Packet: returns (in_bytes, bitstream bytes, [packet bytes, ...])."""
        a = np.ascontiguousarray(pcm_f32, dtype=np.float32)
        pk = np.zeros(1 << 14, np.uint8)
        nb = (C.c_int * 2)(0, 0)
        f = lib().hmp3_L3_audio_encode_Packet
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        f.restype = C.c_uint64
        i, o = self._io(f(self.h, vp(a), vp(self.out) if want_bs else None, vp(pk), nb))
        return i, self.out[:o].copy(), [pk[:nb[0]].copy(), pk[nb[0]:nb[0] + nb[1]].copy()]

    def frames(self):
        return int(lib().hmp3_L3_audio_encode_get_frames(self.h))

    def frames_bytes(self):
        f = lib().hmp3_L3_audio_encode_get_frames_bytes
        f.argtypes = [C.c_void_p]
        f.restype = C.c_uint64
        return self._io(f(self.h))

    def bitrate2(self):
        f = lib().hmp3_L3_audio_encode_get_bitrate2_float
        f.argtypes = [C.c_void_p]
        f.restype = C.c_float
        return float(f(self.h))

    def bitrate(self):
        return float(lib().hmp3_L3_audio_encode_get_bitrate_float(self.h))

    def info_string(self):
        buf = C.create_string_buffer(256)
        lib().hmp3_L3_audio_encode_info_string(self.h, buf)
        return buf.value.decode()

    def info_ec(self):
        a = np.zeros(len(EC_FIELDS), np.int32)
        lib().hmp3_L3_audio_encode_info_ec(self.h, vp(a))
        return dict(zip(EC_FIELDS, a.tolist()))


class MpegHead(C.Structure):
    _fields_ = [(n, C.c_int) for n in ["sync", "id", "option", "prot", "br_index", "sr_index", "pad", "private_bit", "mode",
                                       "mode_ext", "cr", "original", "emphasis"]]


def effective_control(ec):
    """(effective control image, MpegHead) as CMp3Enc::L3_audio_encode_info_ec / _info_head report them."""
    out = np.zeros(len(EC_FIELDS), np.int32)
    h = MpegHead()
    r = lib().hmp3_effective_control(vp(ec), vp(out), C.byref(h))
    if r != 0:
        raise Hmp3Error("control rejected (%d)" % r)
    return out, h


def info_frame(ec, channels, nsamples, audio, frames, frames_after_call, bytes_after_call, xing_flag=67, source_rate=None):
    """The Xing/Info frame the reference CLI puts in front of `audio` (hmp3_info_frame)."""
    eff, head = effective_control(ec)
    buf = np.zeros(2048, np.uint8)
    fa = np.ascontiguousarray(frames_after_call, dtype=np.int32)
    ba = np.ascontiguousarray(bytes_after_call, dtype=np.int64)
    a = np.ascontiguousarray(audio, dtype=np.uint8)
    f = lib().hmp3_info_frame
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p,
                  C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    n = f(vp(eff), head.mode, xing_flag, int(source_rate or eff[EC_FIELDS.index("samprate")]), channels, nsamples, vp(a),
          a.size, frames, vp(fa), vp(ba), fa.size, vp(buf), buf.size)
    return buf[:n].copy()
