"""Load balance of the serial stage: per k_rate launch, the clocks every stream's warp spent; the launch lasts as
long as its slowest stream.  usage: rate_balance.py [nstreams] [seconds]"""
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
L = capi.lib()
L.hmp3_debug_rate_cycles.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
b.run()
L.hmp3_debug_rate_cycles(b.h, None, 0)
b.run()
cyc = np.zeros((64, n), np.int64)
m = L.hmp3_debug_rate_cycles(b.h, cyc.ctypes.data_as(C.c_void_p), 64)
cyc = cyc[:m].astype(np.float64)
print("launches", m, "streams", n, " run ms", b.last_run_ms())
print("launch   mean Mclk    max Mclk   max/mean")
for k in range(m):
    print("%4d   %10.2f  %10.2f   %.3f" % (k, cyc[k].mean() / 1e6, cyc[k].max() / 1e6, cyc[k].max() / cyc[k].mean()))
tot = cyc.sum(0)
print("sum over launches of max: %.1f Mclk;  max over streams of sum: %.1f;  mean of sums: %.1f" % (
    cyc.max(1).sum() / 1e6, tot.max() / 1e6, tot.mean() / 1e6))
print("per base clip mean of sums (Mclk):", " ".join("%.0f" % (tot[i::NB].mean() / 1e6) for i in range(NB)))
print("slot utilisation if launches end at their slowest stream: %.3f; if only the slowest stream bounds the run: %.3f" % (
    tot.mean() / cyc.max(1).sum(), tot.mean() / tot.max()))
