"""GPU parity, end to end: the CUDA pipeline (Phase A + serial stage + frame assembly) through the C ABI
against the unmodified reference on the same PCM.  Bar: byte-identical MP3 streams."""
import numpy as np
import pytest

import refmod
from configs import CONFIGS
from hmp3_b200 import capi
from hmp3_b200.synth import synth_pcm

pytestmark = pytest.mark.gpu


def ref_encode(sr, nch, kw, pcm):
    out, _ = refmod.ref_encode_clip(refmod.make_ec(samprate=sr, nch=nch, **kw), pcm)
    return out


@pytest.mark.parametrize("name,seed,sr,nch,kw", CONFIGS)
def test_clip_bytes_match_reference(name, seed, sr, nch, kw):
    pcm = synth_pcm(seed, 12.0, sr, nch)
    got = capi.encode_batch([capi.control(samprate=sr, nch=nch, **kw)], [pcm])[0]
    ref = ref_encode(sr, nch, kw, pcm)
    assert got.size == ref.size
    assert np.array_equal(got, ref)


def test_baseline_configs_at_their_stated_60_seconds():
    """BASELINE.json configs 1-4 at the length they are stated for (60 s), all five in one batch."""
    ctl = [capi.control(samprate=sr, nch=nch, **kw) for (_, _, sr, nch, kw) in CONFIGS]
    pcms = [synth_pcm(seed, 60.0, sr, nch) for (_, seed, sr, nch, _) in CONFIGS]
    outs = capi.encode_batch(ctl, pcms)
    for (name, seed, sr, nch, kw), pcm, got in zip(CONFIGS, pcms, outs):
        ref = ref_encode(sr, nch, kw, pcm)
        assert got.size == ref.size and np.array_equal(got, ref), name


def test_mixed_batch_matches_reference():
    """Streams with different controls, rates, channel counts and ragged lengths in one batch."""
    ctl, pcms, refs = [], [], []
    for k, (name, seed, sr, nch, kw) in enumerate(CONFIGS * 2):
        secs = [3.0, 1.7, 5.03, 0.4, 2.5, 0.02, 4.0, 1.0, 0.0, 2.2][k]
        pcm = synth_pcm(seed + 100 + k, max(secs, 0.001), sr, nch)
        if secs == 0.0:
            pcm = pcm[:0]
        ctl.append(capi.control(samprate=sr, nch=nch, **kw))
        pcms.append(pcm)
        refs.append(ref_encode(sr, nch, kw, pcm))
    outs = capi.encode_batch(ctl, pcms)
    for k, (o, r) in enumerate(zip(outs, refs)):
        assert o.size == r.size and np.array_equal(o, r), "stream %d differs" % k


@pytest.mark.parametrize("slots,warps,opts", [(8, 3, 0), (16, 16, 8), (5, 2, 0)])
def test_phase_scheduler_with_many_streams_per_block(monkeypatch, slots, warps, opts):
    """The phase-scheduled serial stage with several streams per block and several warps per block (the shape it has at
    scale: a warp claims a stream that needs the block's current phase, another warp runs that stream's next phase),
    on a mixed ragged batch; nested-driver kernel and reference as the checks."""
    ctl, pcms = [], []
    for k in range(44):
        name, seed, sr, nch, kw = CONFIGS[k % len(CONFIGS)]
        secs = 0.6 + 0.37 * (k % 7)
        pcms.append(synth_pcm(seed + 300 + k, secs, sr, nch))
        ctl.append(capi.control(samprate=sr, nch=nch, **kw))
    monkeypatch.setenv("HMP3_RATE_MODE", "phased")   # (a batch this small would get the nested kernel by itself)
    monkeypatch.setenv("HMP3_RATE_PH_SLOTS", str(slots))
    monkeypatch.setenv("HMP3_RATE_PH_WARPS", str(warps))
    monkeypatch.setenv("HMP3_RATE_PH_OPTS", str(opts))
    outs = capi.encode_batch(ctl, pcms)
    outs2 = capi.encode_batch(ctl, pcms)
    monkeypatch.setenv("HMP3_RATE_MODE", "nested")
    outs3 = capi.encode_batch(ctl, pcms)
    for k in range(len(pcms)):
        assert np.array_equal(outs[k], outs2[k]), "stream %d: run-to-run difference" % k
        assert np.array_equal(outs[k], outs3[k]), "stream %d: differs from the nested-driver kernel" % k
    for k in range(0, len(pcms), 3):
        name, seed, sr, nch, kw = CONFIGS[k % len(CONFIGS)]
        ref = ref_encode(sr, nch, kw, pcms[k])
        assert outs[k].size == ref.size and np.array_equal(outs[k], ref), "stream %d differs from the reference" % k


@pytest.mark.parametrize("mode", ["phased", "nested"])
@pytest.mark.parametrize("name,seed,sr,nch,kw", [CONFIGS[0], CONFIGS[2], CONFIGS[3]])
def test_serial_stage_tap_on_the_device(monkeypatch, name, seed, sr, nch, kw, mode):
    """Stage parity of the device's rate loop itself (not only of the bytes it leads to): the quantised lines, scale
    factors and side information the serial stage hands the packing pass, granule by granule, against the host build of
    the same routines -- whose taps are pinned to the reference's (tests/test_cpu_parity.py) -- and, for C1, against
    the reference's own dump of the first granules (tests/golden/c1_head.npz)."""
    import os
    import simmod
    monkeypatch.setenv("HMP3_RATE_MODE", mode)   # both forms of the serial stage: phase-scheduled and one warp per stream
    pcm = synth_pcm(seed, 10.0, sr, nch)[:int(4.0 * sr)]
    ctl = capi.control(samprate=sr, nch=nch, **kw)
    G = 2 * (pcm.shape[0] // 1152)
    b = capi.Batch([ctl], [pcm.shape[0]])
    tap = b.set_rate_tap(0, G)
    outs, _ = b.encode_host([pcm])
    got = tap.copy()
    b.close()
    ref_bytes, _, tr = simmod.encode_clip(ctl, pcm, max_trace_granules=G)
    assert np.array_equal(outs[0], ref_bytes)
    shorts = 0
    for K in range(G):
        for c in range(nch):
            gr = got[K, c]["gr"]
            assert np.array_equal(gr[:24], tr[K, 27 * c:27 * c + 24]), "side information of granule %d ch %d" % (K, c)
            if gr[20] == 0:      # aux_not_null: nothing coded, scale factors and lines are not transmitted
                continue
            if gr[5] == 2:       # short block: s[w][i]
                shorts += 1
                want = tr[K, 100 + 39 * c:100 + 39 * c + 39].reshape(3, 13)[:, :12]
                have = got[K, c]["sf"][23:23 + 39].reshape(3, 13)[:, :12]
            else:
                want = tr[K, 54 + 23 * c:54 + 23 * c + 21]
                have = got[K, c]["sf"][:21]
            assert np.array_equal(have.astype(np.int32), want), "scale factors of granule %d ch %d" % (K, c)
            if K & 1:            # the host build's trace keeps the lines of the call's last granule
                extent = 2 * int(gr[21] + gr[22] + gr[23]) + 4 * int(gr[18])
                want_ix = tr[K, 200 + 576 * c:200 + 576 * c + extent]
                assert np.array_equal(got[K, c]["ix"][:extent].astype(np.int32), want_ix), "lines of granule %d ch %d" % (K, c)
    if name == CONFIGS[2][0]:
        assert shorts > 0    # the 48 kHz clip switches blocks
    if name == CONFIGS[0][0]:
        h = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_head.npz"))
        for k in np.nonzero(h["valid"] > 0)[0]:
            if k >= G:
                continue
            for c in range(nch):
                assert np.array_equal(got[k, c]["gr"][:24], h["gr"][k, c, :24]), "reference side info, granule %d ch %d" % (k, c)
                assert np.array_equal(got[k, c]["sf"][:21].astype(np.int32), h["sf_l"][k, c, :21])
                gr = got[k, c]["gr"]
                extent = 2 * int(gr[21] + gr[22] + gr[23]) + 4 * int(gr[18]) if gr[20] else 0
                assert np.array_equal(got[k, c]["ix"][:extent].astype(np.int32), h["ix"][k, c, :extent]), \
                    "reference lines, granule %d ch %d" % (k, c)


def test_plan_reuse_and_device_resident_run():
    """A plan encodes repeatedly with identical results; the device-resident leg equals the host leg."""
    name, seed, sr, nch, kw = CONFIGS[0]
    pcms = [synth_pcm(seed + i, 2.0, sr, nch) for i in range(40)]
    ctl = [capi.control(samprate=sr, nch=nch, **kw)] * len(pcms)
    b = capi.Batch(ctl, [p.shape[0] for p in pcms])
    outs1, _ = b.encode_host(pcms)
    b.run()
    flat, off, nb, nf, st = b.download_all()
    assert (st == 0).all()
    for i in range(len(pcms)):
        assert np.array_equal(flat[off[i]:off[i] + nb[i]], outs1[i])
    for i in (0, 17, 39):
        assert np.array_equal(outs1[i], ref_encode(sr, nch, kw, pcms[i]))
    assert b.launches() > 0
    b.close()


def test_bad_control_is_rejected():
    ec = capi.control(samprate=44100, nch=2, bitrate=16)   # below the MPEG-1 minimum: reference init returns 0
    with pytest.raises(capi.Hmp3Error):
        capi.encode_batch([ec], [synth_pcm(1, 0.5, 44100, 2)])


@pytest.mark.parametrize("name,seed,sr,nch,kw", [CONFIGS[0], CONFIGS[3], CONFIGS[4]])
def test_encoder_handle_matches_reference_call_by_call(name, seed, sr, nch, kw):
    """The CMp3Enc mirror: same bytes per call as CMp3Enc::L3_audio_encode, same flush protocol, same getters."""
    pcm = synth_pcm(seed + 7, 4.0, sr, nch)
    ref_bytes, tr = refmod.ref_encode_clip(refmod.make_ec(samprate=sr, nch=nch, **kw), pcm, max_trace_calls=4000)
    enc = capi.Encoder(capacity_seconds=30)
    # init returns what Csrc::sr_convert_init does: the bytes that must be buffered for a call = 1153 sample frames;
    # a call consumes 1152
    assert enc.init_mp3(capi.control(samprate=sr, nch=nch, **kw)) == 1153 * 2 * nch
    per_call = 2304 * nch
    ncalls_real = (pcm.shape[0] + 3 * 1153 + 1152) // 1152
    padded = np.zeros(((ncalls_real + 40) * 1152, nch), np.int16)
    padded[:pcm.shape[0]] = pcm
    out, c = [], 0
    frames_expected = ncalls_real * (1 if sr >= 32000 else 2)
    # the CLI's protocol: real calls, then zero PCM until every expected frame has been written
    while c < ncalls_real or enc.frames() < frames_expected:
        used, b = enc.encode_mp3(padded[c * 1152:(c + 1) * 1152])
        assert used == per_call
        if c < len(tr):
            assert b.size == tr["out_bytes"][c], "call %d" % c
        out.append(b)
        c += 1
    got = np.concatenate(out)
    assert got.size == ref_bytes.size and np.array_equal(got, ref_bytes)
    assert enc.frames() >= frames_expected
    assert abs(enc.bitrate() - 8e-3 * got.size * sr / ((1152 if sr >= 32000 else 576) * enc.frames())) < 1e-3
    assert "Layer III" in enc.info_string()
    enc.close()


def test_encoder_handle_float_entry_and_rejections():
    enc = capi.Encoder(capacity_seconds=10)
    assert enc.init_mp3(capi.control(samprate=44100, nch=2, bitrate=64), source_bits=12) == 0   # no such source type
    assert enc.init_mp3(capi.control(samprate=44100, nch=2, bitrate=64), source_bits=16, source_is_float=1) == 0
    assert enc.init_mp3(capi.control(samprate=44100, nch=2, bitrate=16)) == 0                    # reference rejects
    assert enc.init_l3(capi.control(samprate=44100, nch=2, bitrate=64)) == 2 * 4 * 1152
    pcm = synth_pcm(3, 1.0, 44100, 2).astype(np.float32)
    pcm += np.random.default_rng(2).uniform(-0.5, 0.5, size=pcm.shape).astype(np.float32)      # not 16-bit values
    ref_bytes, _ = refmod.ref_encode_clip(refmod.make_ec(samprate=44100, nch=2, bitrate=64), pcm[:20 * 1152])
    out = []
    for c in range(20 + 4 + 6):
        blk = pcm[c * 1152:(c + 1) * 1152].astype(np.float32) if c < 20 else np.zeros((1152, 2), np.float32)
        used, b = enc.encode_l3(blk)
        assert used == 2 * 4 * 1152
        out.append(b)
    got = np.concatenate(out)
    assert np.array_equal(got[:ref_bytes.size], ref_bytes)
    assert enc.info_ec()["bitrate"] == 64
    enc.close()


def test_pinned_host_buffers_take_the_staged_path():
    """Pinned PCM is copied chunk by chunk ahead of each chunk's Phase A and pinned outputs are written directly:
    same bytes as the pageable path, for aligned and odd-aligned sources, stereo and mono."""
    import torch
    specs = [(44100, 2, dict(bitrate=64), 7.3, 0), (44100, 2, dict(bitrate=64), 3.1, 1), (22050, 1, dict(bitrate=32), 6.2, 1),
             (22050, 1, dict(bitrate=32), 6.2, 2), (48000, 2, dict(vbr_mnr=100, hf=2, freq_limit=19000), 5.0, 3)]
    ctl, pcms, pins, ptrs = [], [], [], []
    for k, (sr, nch, kw, secs, skip) in enumerate(specs):
        full = synth_pcm(900 + k, secs, sr, nch)
        t = torch.from_numpy(full).pin_memory()
        view = t[skip:]                                   # odd sample offsets: 2- or 4-byte aligned only
        pins.append(t)
        pcms.append(view.numpy())
        ptrs.append(view.data_ptr())
        ctl.append(capi.control(samprate=sr, nch=nch, **kw))
    want = capi.encode_batch(ctl, [np.array(p) for p in pcms])      # pageable copies
    b = capi.Batch(ctl, [p.shape[0] for p in pcms])
    outs = [torch.zeros(int(c), dtype=torch.uint8).pin_memory() for c in b.bound]
    nb, nf, st = b.encode_host_ptrs(np.array(ptrs, dtype=np.uint64), np.array([o.data_ptr() for o in outs], dtype=np.uint64),
                                    b.bound)
    assert (st == 0).all()
    for i in range(len(specs)):
        assert nb[i] == want[i].size and np.array_equal(outs[i].numpy()[:nb[i]], want[i]), i
    b.close()


def test_cli_writes_the_same_file_as_the_reference_cli(tmp_path):
    """hmp3b200 in.wav out.mp3 [opts] against oracle/_ref/hmp3 with the same arguments: identical files, Xing/Info
    frame included; plus the batch extension (-@ list)."""
    import os
    import struct
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "hmp3_b200", "_lib", "hmp3b200")
    ref = os.path.join(root, "oracle", "_ref", "hmp3")
    assert os.path.exists(cli) and os.path.exists(ref)
    cases = [("c1", 1234, 44100, 2, ["-B64"], 7.0), ("c2", 1234, 44100, 2, [], 33.0), ("c3", 1235, 48000, 2, ["-V100", "-HF2", "-F19000"], 5.0),
             ("c4a", 1236, 22050, 1, ["-B32"], 6.0), ("c4b", 1237, 32000, 2, [], 4.0), ("x0", 1, 44100, 2, ["-B96", "-X0"], 3.0)]
    wavs = []
    for name, seed, sr, nch, opts, secs in cases:
        pcm = synth_pcm(seed + 11, secs, sr, nch)
        data = np.ascontiguousarray(pcm, dtype="<i2").tobytes()
        wav = str(tmp_path / (name + ".wav"))
        with open(wav, "wb") as f:
            f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " +
                    struct.pack("<IHHIIHH", 16, 1, nch, sr, sr * nch * 2, nch * 2, 16) + b"data" + struct.pack("<I", len(data)))
            f.write(data)
        wavs.append(wav)
        a, b = str(tmp_path / (name + "_ref.mp3")), str(tmp_path / (name + "_gpu.mp3"))
        subprocess.run([ref, wav, a] + opts, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
        subprocess.run([cli, wav, b] + opts, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
        x, y = np.fromfile(a, dtype=np.uint8), np.fromfile(b, dtype=np.uint8)
        assert x.size == y.size and np.array_equal(x, y), name
    # batch extension: the two 44.1 kHz files with the same options in one GPU batch
    lst = str(tmp_path / "list.txt")
    with open(lst, "w") as f:
        f.write("%s %s\n%s %s\n" % (wavs[0], tmp_path / "b0.mp3", wavs[1], tmp_path / "b1.mp3"))
    subprocess.run([cli, "-@", lst, "-B64"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    assert np.array_equal(np.fromfile(str(tmp_path / "b0.mp3"), dtype=np.uint8), np.fromfile(str(tmp_path / "c1_ref.mp3"), dtype=np.uint8))
