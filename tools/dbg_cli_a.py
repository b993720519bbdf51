import os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, wavutil
from hmp3_b200.synth import synth_pcm
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CLI = os.path.join(ROOT, "hmp3_b200", "_lib", "hmp3b200"); REF = os.path.join(ROOT, "oracle", "_ref", "hmp3")
for sr, nch, opts in [(16000, 1, ["-A1", "-B48"]), (16000, 2, ["-A1", "-B48"]), (22050, 1, ["-A1", "-B48"]), (24000, 1, ["-A1"]), (16000, 1, ["-A32000", "-B64"]), (12000, 1, ["-B24"])]:
    samples = wavutil.make_samples(synth_pcm(84, 2.0, sr, nch)[:30000], "s16", seed=3)
    wav = "/tmp/a_%d.wav" % sr
    wavutil.write_wav(wav, samples, "s16", sr, nch)
    r = subprocess.run([CLI, wav, "/tmp/a_gpu.mp3"] + opts, capture_output=True, text=True)
    r2 = subprocess.run([REF, wav, "/tmp/a_ref.mp3"] + opts, capture_output=True, text=True)
    a = np.fromfile("/tmp/a_gpu.mp3", dtype=np.uint8) if os.path.exists("/tmp/a_gpu.mp3") else np.zeros(0, np.uint8)
    b = np.fromfile("/tmp/a_ref.mp3", dtype=np.uint8) if os.path.exists("/tmp/a_ref.mp3") else np.zeros(0, np.uint8)
    n = min(a.size, b.size)
    d = np.nonzero(a[:n] != b[:n])[0]
    print(sr, nch, opts, "sizes", a.size, b.size, "diff bytes", d.size, "first", d[:6].tolist(), "last", d[-3:].tolist(), "|", r.stderr.strip().split("\n")[-1][:80])
    for f in ("/tmp/a_gpu.mp3", "/tmp/a_ref.mp3"):
        if os.path.exists(f): os.remove(f)
