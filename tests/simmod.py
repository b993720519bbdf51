"""ctypes view of the test-only host simulator (tests/hostsim/hostsim.cpp)."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM_SO = os.path.join(ROOT, "hmp3_b200", "_lib", "libhmp3_sim.so")
_lib = None


def available():
    return os.path.exists(SIM_SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(SIM_SO)
    return _lib


def vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


INFO_NAMES = ["nchan", "h_id", "sr_index", "nband", "band_limit", "nsb", "nsb_limit", "nsb_limitMS0",
              "nsb_limitMS1", "AveTargetBits", "framebytes", "main_framebytes", "side_bytes", "remainder",
              "divisor", "ms_flag", "is_flag", "iL3", "iencode", "ivbr_min", "ivbr_max", "vbr_pool_target",
              "short_block_threshold", "h_mode", "br_index", "totbitrate", "samprate", "band_limit_stereo",
              "sf_bit_max", "nsf_stereo", "head0", "head1", "head2", "head3", "hf_flag", "filter_select"]


def resolve(ec):
    out = np.zeros(40, np.int32)
    r = lib().sim_resolve(vp(ec), vp(out))
    d = dict(zip(INFO_NAMES, out[:36].tolist()))
    d["unsupported"] = int(out[38])
    d["bytes_in"] = r
    return d


def table(ec, name, dtype, count):
    a = np.zeros(count, dtype)
    r = lib().sim_table(vp(ec), name.encode(), vp(a), a.nbytes)
    assert r >= 0, (name, r)
    return a


def analysis(ec, pcm_i16, ngran, nch):
    pcm = np.ascontiguousarray(pcm_i16)
    out = dict(sbt=np.zeros((ngran, nch, 576), np.float32), ginfo=np.zeros((ngran, 4), np.int32),
               xr=np.zeros((ngran, nch, 576), np.float32), sigmask=np.zeros((ngran, nch, 36, 2), np.float32),
               ms_raw=np.zeros(ngran, np.int32), att=np.zeros((ngran, nch, 9), np.int32),
               raw=np.zeros((ngran, nch, 92), np.float32))
    f = lib().sim_analysis
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int] + [C.c_void_p] * 7
    r = f(vp(ec), vp(pcm), pcm.shape[0], ngran, vp(out["sbt"]), vp(out["ginfo"]), vp(out["xr"]),
          vp(out["sigmask"]), vp(out["ms_raw"]), vp(out["att"]), vp(out["raw"]))
    assert r == 0
    return out


def encode_clip(ec, pcm_i16, max_trace_granules=0, tail=0.0):
    """Whole-clip encode through the host build of the kernel bodies. Returns (mp3 bytes, nframes, trace)."""
    is_float = np.asarray(pcm_i16).dtype == np.float32      # float PCM on the +-32768 scale
    pcm = np.ascontiguousarray(pcm_i16, dtype=np.float32 if is_float else np.int16)
    n = pcm.shape[0]
    cap = 4096 + int(n / 1152 + 80) * 2100
    out = np.zeros(cap, np.uint8)
    tr = np.zeros((max_trace_granules, 1400), np.int32) if max_trace_granules else None
    nf = C.c_int(0)
    f = lib().sim_encode_clip_any
    f.restype = C.c_long
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_int, C.c_void_p]
    r = f(vp(ec), vp(pcm), 1 if is_float else 0, float(tail), n, vp(out), cap, vp(tr), max_trace_granules, C.byref(nf))
    if r < 0:
        raise RuntimeError("sim_encode_clip failed %d" % r)
    return out[:r].copy(), nf.value, tr
