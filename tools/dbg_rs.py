import os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from wavutil import write_wav
from hmp3_b200.synth import synth_pcm
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref")
for opts, sr, nch in [(["-A44100"], 32000, 1), (["-B48", "-A1"], 16000, 1), (["-B64", "-A44100"], 32000, 1), (["-A44100"], 32000, 2), (["-B48", "-A1"], 16000, 2)]:
    wav = "/tmp/rs.wav"
    write_wav(wav, synth_pcm(6200 + sr // 1000, 3.0, sr, nch), "s16", sr, nch)
    r1 = subprocess.run([os.path.join(R, "tomp3_gpu"), wav, "/tmp/a.mp3"] + opts, capture_output=True, text=True)
    r2 = subprocess.run([os.path.join(R, "hmp3"), wav, "/tmp/b.mp3"] + opts, capture_output=True, text=True)
    a = np.fromfile("/tmp/a.mp3", np.uint8); b = np.fromfile("/tmp/b.mp3", np.uint8)
    n = min(a.size, b.size); d = np.nonzero(a[:n] != b[:n])[0]
    print(opts, sr, nch, "sizes", a.size, b.size, "ndiff", d.size, "first", d[:8].tolist(), "last", d[-3:].tolist())
    print("   gpu:", [l for l in r1.stderr.splitlines() if "Kbps" in l or "rame" in l][-2:])
    print("   ref:", [l for l in r2.stderr.splitlines() if "Kbps" in l or "rame" in l][-2:])
