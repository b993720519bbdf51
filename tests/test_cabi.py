"""The C-ABI library without a GPU: it loads, exports every symbol include/*.h declares, its pure-host
entries work, and every device entry fails loudly (there is no CPU fallback)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from configs import CONFIGS
from hmp3_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def declared_symbols():
    names = set()
    for h in ("hmp3_b200.h", "hmp3_b200_debug.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
        names |= set(re.findall(r"\b(hmp3_[A-Za-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    syms = declared_symbols()
    assert len(syms) >= 35
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_control_defaults_and_cli_options():
    L = capi.lib()
    ec = np.zeros(len(capi.EC_FIELDS), np.int32)
    L.hmp3_control_defaults(capi.vp(ec))
    assert np.array_equal(ec, capi.control())                      # the CLI defaults: VBR-50
    L.hmp3_control_apply_option.argtypes = [C.c_void_p, C.c_char_p]
    for opt in (b"-B64",):
        assert L.hmp3_control_apply_option(capi.vp(ec), opt) == 0
    assert np.array_equal(ec, capi.control(bitrate=64))            # -B => CBR
    ec2 = capi.control()
    for opt in (b"-V100", b"-HF2", b"-F19000"):
        assert L.hmp3_control_apply_option(capi.vp(ec2), opt) == 0
    assert np.array_equal(ec2, capi.control(vbr_mnr=100, hf=2, freq_limit=19000))
    assert L.hmp3_control_apply_option(capi.vp(ec2), b"-?") == -1


@pytest.mark.parametrize("name,seed,sr,nch,kw", CONFIGS)
def test_resolve_control_is_pure_host(name, seed, sr, nch, kw):
    class R(C.Structure):
        _fields_ = [(n, C.c_int) for n in
                    ["nchan", "h_id", "sr_index", "nband", "band_limit", "nsb", "nsb_limit", "nsb_limit_ms0",
                     "nsb_limit_ms1", "ave_target_bits", "framebytes", "main_framebytes", "side_bytes", "remainder",
                     "divisor", "ms_flag", "is_flag", "frame_driver", "granule_driver", "ivbr_min", "ivbr_max",
                     "vbr_pool_target", "short_block_threshold", "h_mode", "br_index", "totbitrate", "samprate",
                     "band_limit_stereo", "sf_bit_max", "nsf_stereo"]] + [("head", C.c_int * 4)] + \
                   [(n, C.c_int) for n in ["hf_flag", "filter_select", "bytes_in"]]
    r = R()
    assert capi.lib().hmp3_resolve_control(capi.vp(capi.control(samprate=sr, nch=nch, **kw)), C.byref(r)) == 0
    want = GOLD["configs"][name]["resolved"]
    assert (r.nchan, r.h_id, r.nsb, r.nsb_limit, r.band_limit, r.ave_target_bits, r.framebytes, r.main_framebytes,
            r.side_bytes, r.bytes_in) == (want["nchan"], want["h_id"], want["nsb"], want["nsb_limit"],
                                          want["band_limit"], want["AveTargetBits"], want["framebytes"],
                                          want["main_framebytes"], want["side_bytes"], want["bytes_in"])
    assert list(r.head) == [want["head0"], want["head1"], want["head2"], want["head3"]]


def test_out_bound_covers_the_output():
    L = capi.lib()
    L.hmp3_batch_out_bound.restype = C.c_int64
    L.hmp3_batch_out_bound.argtypes = [C.c_void_p, C.c_int64]
    for name, seed, sr, nch, kw in CONFIGS:
        n = int(GOLD["seconds"] * sr)
        assert L.hmp3_batch_out_bound(capi.vp(capi.control(samprate=sr, nch=nch, **kw)), n) >= \
            GOLD["configs"][name]["mp3_bytes"]


def test_no_device_means_loud_failure():
    """Only meaningful on a box without a GPU: no entry point may silently compute on the CPU."""
    L = capi.lib()
    if L.hmp3_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.Hmp3Error) as e:
        capi.encode_batch([capi.control(bitrate=64)], [np.zeros((4000, 2), np.int16)])
    assert "no usable CUDA device" in str(e.value)
    with pytest.raises(capi.Hmp3Error):
        capi.Encoder()
    with pytest.raises(capi.Hmp3Error):
        capi.debug_analysis(capi.control(bitrate=64), np.zeros((4000, 2), np.int16), 4, 2)
