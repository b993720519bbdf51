"""Experiment: the same 4736 streams as ONE plan vs TWO plans of half the streams running concurrently on one GPU
(their serial-stage gaps do not coincide when the chunk lengths differ). usage: two_plans.py [nstreams] [seconds] [ngA] [ngB]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hmp3_b200 import capi
from hmp3_b200.synth import synth_pcm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
ngs = [int(sys.argv[3]) if len(sys.argv) > 3 else 256, int(sys.argv[4]) if len(sys.argv) > 4 else 192]
NB = 16
base = [synth_pcm(1000 + i, secs + 10.0, 44100, 2) for i in range(NB)]
ns = int(secs * 44100)

def clip(i):
    s0 = (563 * (i // NB)) % (10 * 44100)
    return base[i % NB][s0:s0 + ns]

def make(idx, ng):
    os.environ["HMP3_CHUNK_GRANULES"] = str(ng)
    b = capi.Batch([capi.control(bitrate=64)] * len(idx), [ns] * len(idx))
    for j, i in enumerate(idx):
        b.upload(j, clip(i))
    return b

one = make(list(range(n)), 256)
for it in range(2):
    t0 = time.time(); one.run(); t1 = time.time()
print("one plan : %.1f ms -> %.0f x realtime, bytes %d" % ((t1 - t0) * 1e3, n * secs / (t1 - t0), one.results()[0].sum()))
tot1 = one.results()[0].sum()
one.close()
halves = [make(list(range(0, n, 2)), ngs[0]), make(list(range(1, n, 2)), ngs[1])]
for it in range(2):
    t0 = time.time()
    for b in halves: b.run(async_=True)
    for b in halves: b.sync()
    t1 = time.time()
tot2 = sum(b.results()[0].sum() for b in halves)
print("two plans (%d/%d granule chunks): %.1f ms -> %.0f x realtime, bytes %d (%s)" % (
    ngs[0], ngs[1], (t1 - t0) * 1e3, n * secs / (t1 - t0), tot2, "same" if tot1 == tot2 else "DIFFERENT"))
