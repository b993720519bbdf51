#!/usr/bin/env python3
"""Generate tests/golden/ from the UNMODIFIED reference (oracle/_ref/libhmp3ref.so; needs /root/reference to
have been built by `make -C oracle`).  The reference ships no test vectors of its own,
so these fixtures pin the oracle build itself: a later oracle build, the host build of the kernel bodies and
the CUDA path must all reproduce them.

  golden.json      per BASELINE config (10 s synthetic clip): resolved init values, MP3 size / md5 / frame
                   count, md5s of the stage traces (block types, MDCT lines, sig/mask, ix, side info)
  c1_head.npz      the first 24 granules of config C1 in full (xr, sig_mask, GR fields, scale factors, ix)
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refmod  # noqa: E402
from configs import CONFIGS  # noqa: E402
from hmp3_b200.synth import synth_pcm  # noqa: E402

SECONDS = 10.0


def md5(a):
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    out = {"seconds": SECONDS, "configs": {}}
    for name, seed, sr, nch, kw in CONFIGS:
        ec = refmod.make_ec(samprate=sr, nch=nch, **kw)
        pcm = synth_pcm(seed, SECONDS, sr, nch)
        mp3, tr = refmod.ref_encode_clip(ec, pcm, max_trace_calls=4000)
        g = tr["g"].reshape(-1)
        v = g["valid"] > 0
        info = refmod.ref_info(ec)
        out["configs"][name] = {
            "pcm_md5": md5(pcm), "resolved": info, "mp3_bytes": int(mp3.size), "mp3_md5": md5(mp3),
            "calls": int(len(tr)), "valid_granules": int(v.sum()),
            "block_type_md5": md5(g["block_type"][v]), "block_type_hist": np.bincount(g["block_type"][v], minlength=4).tolist(),
            "ms_granules": int((g["ms_flag"][v] > 0).sum()),
            "xr_md5": md5(g["xr"][v][:, :nch]), "sigmask_md5": md5(g["sigmask"][v][:, :nch]),
            "ix_md5": md5(g["ix"][v][:, :nch]), "gr_md5": md5(g["gr"][v][:, :nch]),
            "out_bytes_per_call_md5": md5(tr["out_bytes"]),
        }
        if name.startswith("c1"):
            h = g[:24]
            np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c1_head.npz"), xr=h["xr"], sigmask=h["sigmask"],
                                gr=h["gr"], sf_l=h["sf_l"], sf_s=h["sf_s"], ix=h["ix"], valid=h["valid"],
                                block_type=h["block_type"], ms_flag=h["ms_flag"])
    # rejected / out-of-scope controls
    out["rejected"] = {"cbr16_44k": refmod.ref_info(refmod.make_ec(samprate=44100, nch=2, bitrate=16)) is None}
    with open(os.path.join(ROOT, "tests", "golden", "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote tests/golden/golden.json and c1_head.npz")


if __name__ == "__main__":
    main()
