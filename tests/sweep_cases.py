"""A broad sweep of controls: every sample rate, mono and stereo, CBR from 24 to 160 kbit/s per channel, VBR scales,
-HF, -F, -S1 (DC filter), plain stereo (mode 0), short-block threshold, -Q0, -T, -TX0, -L, and the CBitAllo1 configurations:
dual channel (-M2), intensity stereo by -N, and the low rates at which joint stereo turns intensity coding on by itself."""
OPTION_SETS = [dict(), dict(bitrate=24), dict(bitrate=32), dict(bitrate=48), dict(bitrate=64), dict(bitrate=96), dict(bitrate=160),
               dict(vbr_mnr=0), dict(vbr_mnr=120), dict(vbr_mnr=150, hf=2), dict(bitrate=112, hf=1), dict(mode=0, bitrate=64),
               dict(mode=0), dict(freq_limit=8000), dict(bitrate=64, filter_select=1), dict(short_block_threshold=300),
               dict(bitrate=80, quick=0), dict(vbr_delta_mnr=20), dict(bitrate=64, test1=0), dict(vbr_br_limit=96),
               dict(bitrate=64, mode=2), dict(bitrate=32, mode=2), dict(bitrate=64, nsbstereo=8), dict(bitrate=48, nsbstereo=4),
               dict(bitrate=16), dict(bitrate=20, nsbstereo=6), dict(bitrate=40, nsbstereo=12)]


def sweep_cases():
    k = 0
    for sr in (16000, 22050, 24000, 32000, 44100, 48000):
        for nch in (1, 2):
            for kw in OPTION_SETS:
                yield k, sr, nch, kw
                k += 1
