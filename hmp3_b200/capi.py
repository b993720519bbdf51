"""ctypes binding of the product library hmp3_b200/_lib/libhmp3_b200.so (C ABI: include/hmp3_b200.h).

There is no CPU fallback: loading fails loudly if the CUDA library has not been built, and every
encode entry fails if no CUDA device is present."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_lib", "libhmp3_b200.so")

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
