"""g1 evidence: the exact FP32 SIMT polyphase (k_polyphase) against the tensor-core contraction (k_polyphase_mm,
3xTF32 and plain TF32).  Prints one JSON object: kernel time per step (every kernel on one stream), max relative error
of the sub-band samples, frame byte-match rate of whole encodes and the decoded-PCM SNR (test-only decoder).
  python tools/polymm_eval.py [streams] [seconds]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

def child(mode, n, secs):
    from hmp3_b200 import capi
    from hmp3_b200.synth import synth_pcm
    base = [synth_pcm(10000 + i, secs + 1.0, 44100, 2) for i in range(8)]
    N = int(secs * 44100)
    rng = np.random.default_rng(5)
    ctl = [capi.control(bitrate=64)] * n
    b = capi.Batch(ctl, [N] * n)
    for i in range(n):
        k, s0 = i % 8, int(rng.integers(0, 44100))
        b.upload(i, base[k][s0:s0 + N])
    b.set_serialize(True); b.set_timing(True)
    for _ in range(3):
        b.run()
    ph = b.phase_ms()
    flat, off, nb, nf, st = b.download_all()
    assert (st == 0).all()
    first = [flat[off[i]:off[i] + nb[i]].tolist() for i in range(min(n, 4))]
    ec = capi.control(bitrate=64)
    sbt = capi.debug_analysis(ec, base[0][:N], N // 576, 2)["sbt"]
    np.save("/tmp/polymm_sbt_%s.npy" % (mode or "exact"), sbt)
    print(json.dumps({"polyphase_ms": ph["polyphase"][0], "step_ms": b.last_run_ms(), "first": first}))

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    child(sys.argv[2] if sys.argv[2] != "exact" else None, int(sys.argv[3]), float(sys.argv[4]))
    sys.exit(0)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
res = {}
for mode in ("exact", "3xtf32", "tf32"):
    env = dict(os.environ)
    env.pop("HMP3_POLY_MODE", None)
    if mode != "exact":
        env["HMP3_POLY_MODE"] = mode
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", mode, str(n), str(secs)], env=env,
                       stdout=subprocess.PIPE, text=True, timeout=600)
    res[mode] = json.loads(r.stdout.strip().splitlines()[-1])
import mp3dec
from hmp3_b200.synth import synth_pcm
ex = np.load("/tmp/polymm_sbt_exact.npy")
out = {"streams": n, "seconds": secs, "modes": {}}
for mode in res:
    sbt = np.load("/tmp/polymm_sbt_%s.npy" % mode)
    fm, snr = [], []
    for i, bytes_ in enumerate(res[mode]["first"]):
        a, r = np.array(bytes_, np.uint8), np.array(res["exact"]["first"][i], np.uint8)
        k = min(a.size, r.size) // 417
        fm.append(float(np.mean([np.array_equal(a[j * 417:(j + 1) * 417], r[j * 417:(j + 1) * 417]) for j in range(k)])))
    out["modes"][mode] = {"polyphase_ms_per_step": res[mode]["polyphase_ms"], "step_ms": res[mode]["step_ms"],
                          "max_rel_err_subband": float(np.abs(sbt - ex).max() / np.abs(ex).max()),
                          "frame_match_vs_exact": float(np.mean(fm))}
# decoded-PCM SNR of stream 0 (its PCM is rebuilt the same way as in the child)
base0 = synth_pcm(10000, secs + 1.0, 44100, 2)
s0 = int(np.random.default_rng(5).integers(0, 44100))
pcm0 = base0[s0:s0 + int(secs * 44100)]
for mode in res:
    dec = mp3dec.decode(np.array(res[mode]["first"][0], np.uint8), max_frames=200)
    out["modes"][mode]["decoded_snr_db"] = mp3dec.snr_db(pcm0[:dec.shape[0]], dec)[0]
print(json.dumps(out, indent=1))
