"""Kernel timeline of one batch run (CUDA events on the launching streams). usage: timeline.py [nstreams] [seconds]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hmp3_b200 import capi
from hmp3_b200.synth import synth_pcm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
NB = 16
base = [synth_pcm(1000 + i, secs + 10.0, 44100, 2) for i in range(NB)]
ns = int(secs * 44100)
b = capi.Batch([capi.control(bitrate=64)] * n, [ns] * n)
for i in range(n):
    s0 = (563 * (i // NB)) % (10 * 44100)
    b.upload(i, base[i % NB][s0:s0 + ns])
b.run()
b.set_timing(True)
b.run()
names = list(b.phase_ms().keys())
L = capi.lib()
L.hmp3_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
rows = np.zeros((4096, 3), np.float32)
m = L.hmp3_debug_timeline(b.h, rows.ctypes.data_as(C.c_void_p), 4096)
rows = rows[:m]
print("run ms %.1f" % b.last_run_ms())
order = np.argsort(rows[:, 1], kind="stable")
for r in rows[order]:
    print("%-16s %9.2f -> %9.2f  (%8.2f ms)" % (names[int(r[0])], r[1], r[2], r[2] - r[1]))
rate = rows[rows[:, 0] == names.index("rate_loop")]
gaps = rate[1:, 1] - rate[:-1, 2]
print("k_rate busy %.1f ms; before first %.1f; gaps between launches: %s; after last %.1f" % (
    (rate[:, 2] - rate[:, 1]).sum(), rate[0, 1], " ".join("%.1f" % g for g in gaps), b.last_run_ms() - rate[-1, 2]))
