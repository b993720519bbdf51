import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np
from hmp3_b200 import capi
from hmp3_b200.synth import synth_pcm
n = int(sys.argv[1]); secs = float(sys.argv[2]); same = int(sys.argv[3])
base = [synth_pcm(1000 + i, secs, 44100, 2) for i in range(8)]
b = capi.Batch([capi.control(bitrate=64)] * n, [base[0].shape[0]] * n)
for i in range(n):
    b.upload(i, base[0] if same else np.roll(base[i % 8], 997 * (i // 8), axis=0))
b.set_timing(True)
for it in range(2):
    b.run()
print("n=%d same=%d rate_loop %.1f ms" % (n, same, b.phase_ms()["rate_loop"][0]))
