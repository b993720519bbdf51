"""Ad-hoc device timing: N streams of S seconds (44.1k stereo CBR128), per-kernel CUDA-event times."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hmp3_b200 import capi
from hmp3_b200.synth import synth_pcm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
base = [synth_pcm(1000 + i, secs, 44100, 2) for i in range(8)]
ctl = [capi.control(bitrate=64)] * n
b = capi.Batch(ctl, [base[0].shape[0]] * n)
for i in range(n):
    b.upload(i, np.roll(base[i % 8], 997 * (i // 8), axis=0))
b.set_timing(True)
for it in range(2):
    t0 = time.time(); b.run(); t1 = time.time()
    nb, nf, off, st = b.results()
    print("run %d: %.3f s wall, %d streams x %.1f s -> %.0f x realtime; launches %d; bytes %d; status ok %s" % (
        it, t1 - t0, n, secs, n * secs / (t1 - t0), b.launches(), nb.sum(), (st == 0).all()))
    for k, (ms, ln) in b.phase_ms().items():
        print("   %-12s %10.3f ms  %5d launches" % (k, ms, ln))
