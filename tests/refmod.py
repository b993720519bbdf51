"""ctypes view of oracle/_ref/libhmp3ref.so (the unmodified reference + tap harness).
TEST INFRASTRUCTURE ONLY: imported by tests/, tools/ golden generators and bench.py's cpu_baseline
leg; never by the product package."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libhmp3ref.so")

EC_FIELDS = ["mode", "bitrate", "samprate", "nsbstereo", "filter_select", "freq_limit", "nsb_limit", "layer",
             "cr_bit", "original", "hf_flag", "vbr_flag", "vbr_mnr", "vbr_br_limit", "vbr_delta_mnr",
             "chan_add_f0", "chan_add_f1", "sparse_scale"] + ["mnr_adjust%d" % i for i in range(21)] + \
            ["cpu_select", "quick", "test1", "test2", "test3", "short_block_threshold"]
assert len(EC_FIELDS) == 45


def make_ec(samprate=44100, nch=2, bitrate=-1, vbr_mnr=50, hf=0, freq_limit=24000, mode=None, **kw):
    """E_CONTROL image with the CLI defaults (test/tomp3.cpp:357-387, 563-566, 809-815)."""
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


GR_FIELDS = ["part2_3_length", "big_values", "global_gain", "scalefac_compress", "window_switching_flag",
             "block_type", "mixed_block_flag", "table_select0", "table_select1", "table_select2",
             "subblock_gain0", "subblock_gain1", "subblock_gain2", "region0_count", "region1_count", "preflag",
             "scalefac_scale", "count1table_select", "aux_nquads", "aux_bits", "aux_not_null", "aux_nreg0",
             "aux_nreg1", "aux_nreg2", "block_type_prev", "short_flag_current", "short_flag_next"]

GRANULE_DT = np.dtype([
    ("valid", "<i4"), ("nchan", "<i4"), ("ms_flag", "<i4"), ("min_bits", "<i4"), ("target_bits", "<i4"),
    ("max_bits", "<i4"), ("bit_pool", "<i4"), ("block_type", "<i4"), ("block_type_prev", "<i4"),
    ("short_flag_current", "<i4"), ("short_flag_next", "<i4"), ("ms_corr", "<i4"),
    ("xr", "<f4", (2, 576)), ("sigmask", "<f4", (2, 36, 2)), ("gr", "<i4", (2, 27)), ("sf_l", "<i4", (2, 23)),
    ("sf_s", "<i4", (2, 3, 13)), ("ix", "<i4", (2, 576)), ("signx", "u1", (2, 576)), ("scfsi", "<i4", (2,)),
    ("mode_ext", "<i4")])
CALL_DT = np.dtype([
    ("out_bytes", "<i4"), ("byte_pool", "<i4", (2,)), ("byte_min", "<i4", (2,)), ("byte_max", "<i4", (2,)),
    ("attack_buf", "<i4", (2, 32)), ("sbt", "<f4", (2, 2, 576)), ("ecsave", "<f4", (2, 64)),
    ("g", GRANULE_DT, (2,))])

_lib = None


def available():
    return os.path.exists(REF_SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(REF_SO)
        _lib.ref_encode_clip.restype = C.c_long
        _lib.ref_encode_clip.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p,
                                         C.c_long, C.c_void_p]
        g, c = C.c_int(), C.c_int()
        _lib.ref_trace_sizes(C.byref(g), C.byref(c))
        assert g.value == GRANULE_DT.itemsize, (g.value, GRANULE_DT.itemsize)
        assert c.value == CALL_DT.itemsize, (c.value, CALL_DT.itemsize)
    return _lib


def ref_info(ec):
    L = lib()
    r = L.ref_init(ec.ctypes.data_as(C.c_void_p))
    if r == 0:
        return None
    out = np.zeros(64, np.int32)
    n = L.ref_info(out.ctypes.data_as(C.c_void_p), 64)
    names = ["nchan", "h_id", "sr_index", "nband", "band_limit", "nsb", "nsb_limit", "nsb_limitMS0",
             "nsb_limitMS1", "AveTargetBits", "framebytes", "main_framebytes", "side_bytes", "remainder",
             "divisor", "ms_flag", "is_flag", "iL3", "iencode", "ivbr_min", "ivbr_max", "vbr_pool_target",
             "short_block_threshold", "h_mode", "br_index", "totbitrate", "samprate", "band_limit_stereo",
             "sf_bit_max", "nsf_stereo", "head0", "head1", "head2", "head3", "hf_flag", "filter_select"]
    d = dict(zip(names, out[:n].tolist()))
    d["bytes_in"] = r
    return d


def ref_encode_clip(ec, pcm_i16, max_trace_calls=0):
    """Encode int16 PCM (nsamples, nch) the way the CLI does (int16 -> float cast, the CLI's zero padding at EOF,
    tail flush) and return (mp3_bytes, traces or None).  No Xing/Info frame."""
    L = lib()
    pcm = np.ascontiguousarray(pcm_i16.astype(np.float32))
    n = pcm.shape[0]
    cap = 4096 + int(n / 1152 + 80) * 2100
    out = np.zeros(cap, np.uint8)
    tr = np.zeros(max_trace_calls, CALL_DT) if max_trace_calls else None
    nc = C.c_long(0)
    r = L.ref_encode_clip(ec.ctypes.data_as(C.c_void_p), pcm.ctypes.data_as(C.c_void_p), n,
                          out.ctypes.data_as(C.c_void_p), cap,
                          tr.ctypes.data_as(C.c_void_p) if tr is not None else None, max_trace_calls,
                          C.byref(nc))
    if r < 0:
        raise RuntimeError("reference encode failed: %d" % r)
    if tr is not None:
        tr = tr[:min(max_trace_calls, nc.value)]
    return out[:r].copy(), tr


def same_bytes_or_sf_defect(ref, got):
    """True if two streams are identical, or differ only the way the reference's CBitAllo1 defect makes them differ.
    That allocator can hand the packer a NEGATIVE scale factor (a band whose lines all quantise to zero gets sf = 0 and
    then loses its pre-emphasis amount, bitallo1.cpp:1805-1811, 566-570); l3pack.c writes fields without masking
    (:122-134), so -3 sets every bit still in the packer's 32-bit buffer: up to four bytes of the PREVIOUS
    granule-channel's Huffman data become 0xFF-ish garbage.  The GPU path writes the field masked.  Everything else
    must be byte-identical: same sizes, and at the few differing bytes the reference only has extra one-bits."""
    ref = np.asarray(ref, np.uint8)
    got = np.asarray(got, np.uint8)
    if ref.size != got.size:
        return False
    d = np.nonzero(ref != got)[0]
    if d.size == 0:
        return True
    if d.size > max(8, ref.size // 2000):
        return False
    return bool(np.all((ref[d] & got[d]) == got[d]))
