"""The BASELINE.json configurations as (name, seed, samprate, nch, control kwargs)."""
CONFIGS = [
    ("c1_cbr128_44k", 1234, 44100, 2, dict(bitrate=64)),
    ("c2_vbr50_44k", 1234, 44100, 2, dict()),
    ("c3_v100_hf2_48k", 1235, 48000, 2, dict(vbr_mnr=100, hf=2, freq_limit=19000)),
    ("c4a_cbr32_22k_mono", 1236, 22050, 1, dict(bitrate=32)),
    ("c4b_vbr50_32k", 1237, 32000, 2, dict()),
]
