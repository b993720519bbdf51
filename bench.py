#!/usr/bin/env python3
"""Throughput benchmark of the B200-native Helix-MP3 hot path (BASELINE.json metric).

metric   encoded audio seconds per second (x realtime), 44.1 kHz stereo CBR128
step     one pass of the whole hot path (polyphase -> ... -> packed frames) over one batch of synthetic
         clips: --clips-per-gpu clips of 30 s per GPU (the C5 workload, sharded: no collective, weak scaling), or
         --total-clips T (BASELINE config 5 as written: 10 000 clips in all, split over the GPUs: strong scaling)
input    every stream is its own clip: a seeded mix of two of N_BASE synthetic base clips (seeded gains and window
         positions), so no two streams are equal; the e2e leg reads every stream from its own region of one pinned
         host buffer (25 GB per GPU at the default size)
value    whole-job throughput with the PCM already resident in HBM (default: 9472 clips per GPU = 64 streams per SM)
e2e      the same through the C-ABI host entry (hmp3_batch_encode_host): pinned host PCM -> H2D ->
         kernels -> D2H of the MP3 frames, every step
roofline the dominant kernel (k_rate, the serial stage), algorithmic bytes / CUDA-event launch time
cpu_baseline / --impl reference
         the unmodified reference CLI (oracle/_ref/hmp3, one process per clip, all host cores)

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
"""
import argparse
import json
import os
import shutil
import statistics
import struct
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR, NCH, CLIP_S = 44100, 2, 30.0
CLIP_N = int(SR * CLIP_S)
BASE_S = 32.0                      # base clips are a little longer than the streams cut from them
N_BASE = 32
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "hmp3_lto")   # the reference CLI built with the reference Makefile's flags
if not os.path.exists(REF_BIN):
    REF_BIN = os.path.join(ROOT, "oracle", "_ref", "hmp3")
METRIC = "encoded audio sec/sec (x realtime), 44.1k stereo CBR128"

# algorithmic bytes of the serial stage per granule-channel (DESIGN.md "K6"): MDCT magnitudes in (2304 B) +
# prepared |x|^(3/4), signs, band energies and step bounds in (2816 B) + sig/mask in (288 B) + the record for the
# packing pass out (1396 B)
K6_BYTES_PER_GC = 2304 + 2816 + 288 + 1396
# DRAM traffic of the same kernel per granule-channel, from the committed `ncu --set full` capture (the newest
# profiles/*_rate_traffic.json: {"dram_bytes_per_gc": ...}; per-launch traffic = this x the granule-channels one
# launch processes)
def rate_traffic_per_gc():
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_rate_traffic.json"))):
        try:
            best = (json.load(open(f)), os.path.relpath(f, ROOT))
        except (OSError, ValueError):
            pass
    if best:
        return float(best[0]["dram_bytes_per_gc"]), best[1]
    return (24.749032e9 + 25.115575e9) / (4736 * 256 * 2), "profiles/r1h_rate_ncu_details.txt"


def rate_issue_counters():
    """Issue-slot utilisation and instruction-cache behaviour of the serial stage from the newest committed counter
    capture (profiles/*_rate_counters.json, written by tools/rate_counters.py from an ncu --metrics run)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_rate_counters.json")))
    if not files:
        return None
    try:
        d = json.load(open(files[-1]))
        d["source"] = os.path.relpath(files[-1], ROOT)
        return d
    except (OSError, ValueError):
        return None


# algorithmic work per granule-channel of every kernel: (bound, bytes or flop per gc)
# (k_attack reads sub-bands 4..17 of P and writes nine energies)
KERNEL_WORK = {
    "polyphase": ("fp32_nonfused", 23.9e3), "attack": ("hbm", 14 * 18 * 4 + 36), "hybrid_mdct": ("hbm", 4608 + 2304),
    "psy_stage1": ("hbm", 2304 + 372), "psy_stage2": ("hbm", 368 + 288), "prepare": ("hbm", 4608 + 2304 + 2816),
    "rate_loop": ("hbm", K6_BYTES_PER_GC), "pack": ("hbm", 1396 + 105),
}


def _synth_one(i):
    from hmp3_b200.synth import synth_pcm
    return synth_pcm(10000 + i, BASE_S, SR, NCH)


def base_clips(n=N_BASE):
    """C5-style base clips (SURVEY 8d seeds 10000...), synthesised on all host cores."""
    import multiprocessing as mp
    with mp.get_context("fork").Pool(min(n, ref_cores())) as pool:
        return pool.map(_synth_one, range(n))


def stream_recipe(g):
    """Stream g of the job (global id): (base a, start a, gain a, base b, start b, gain b).  Seeded per stream; two
    different base clips, gains that keep the peak of the mix at the base clips' peak."""
    rng = np.random.default_rng(770000 + g)
    ka = int(rng.integers(0, N_BASE))
    kb = int((ka + 1 + rng.integers(0, N_BASE - 1)) % N_BASE)
    room = int((BASE_S - CLIP_S) * SR)
    sa, sb = int(rng.integers(0, room)), int(rng.integers(0, room))
    ga = float(rng.uniform(0.3, 0.7))
    return ka, sa, ga, kb, sb, 1.0 - ga


def mix_stream_np(clips, g):
    """The PCM of stream g on the host (numpy, float32 arithmetic as on the device)."""
    ka, sa, ga, kb, sb, gb = stream_recipe(g)
    a = clips[ka][sa:sa + CLIP_N].astype(np.float32)
    b = clips[kb][sb:sb + CLIP_N].astype(np.float32)
    return np.rint(np.float32(ga) * a + np.float32(gb) * b).astype(np.int16)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower() == "active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified reference binary, one process per clip
# ------------------------------------------------------------------------------------------------
def write_wav(path, pcm):
    data = np.ascontiguousarray(pcm, dtype="<i2").tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " +
                struct.pack("<IHHIIHH", 16, 1, NCH, SR, SR * NCH * 2, NCH * 2, 16) + b"data" +
                struct.pack("<I", len(data)))
        f.write(data)


def ref_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class RefRunner:
    """Encodes `jobs` 30 s clips with the reference CLI (-B64), `cores` processes at a time."""

    def __init__(self, clips, jobs, distinct=64):
        if not os.path.exists(REF_BIN):
            raise RuntimeError("oracle/_ref/hmp3 is missing (build it with __graft_entry__.build() where "
                               "/root/reference exists)")
        self.cores = ref_cores()
        self.jobs = jobs
        shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
        self.dir = tempfile.mkdtemp(prefix="hmp3_ref_", dir=shm)
        self.wavs = []
        for i in range(min(distinct, jobs)):
            p = os.path.join(self.dir, "c%d.wav" % i)
            write_wav(p, mix_stream_np(clips, i))
            self.wavs.append(p)

    def step(self):
        """One bounded sample; returns wall seconds."""
        work = [(self.wavs[j % len(self.wavs)], os.path.join(self.dir, "o%d.mp3" % (j % (4 * self.cores))))
                for j in range(self.jobs)]
        it = iter(work)
        lock = threading.Lock()

        def worker():
            while True:
                with lock:
                    w = next(it, None)
                if w is None:
                    return
                subprocess.run([REF_BIN, w[0], w[1], "-B64"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)

        t0 = time.perf_counter()
        th = [threading.Thread(target=worker) for _ in range(self.cores)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    def close(self):
        shutil.rmtree(self.dir, ignore_errors=True)

    def sample(self):
        return ("%d clips of 30 s per step (%d distinct streams of the job) through %s -B64 (the unmodified reference "
                "CLI, reference Makefile flags), one process per clip, %d at a time, files on tmpfs"
                % (self.jobs, len(self.wavs), os.path.relpath(REF_BIN, ROOT), self.cores))


def run_reference(args, rank):
    if rank != 0:
        return
    clips = base_clips()
    cores = ref_cores()
    jobs = max(32, 16 * cores)
    rr = RefRunner(clips, jobs)
    try:
        for _ in range(args.warmup):
            rr.step()
        t = sum(rr.step() for _ in range(args.steps))
    finally:
        rr.close()
    value = jobs * CLIP_S * args.steps / t
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "x realtime", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.total_clips else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "x realtime", "cores": cores, "kind": "reference", "sample": rr.sample()},
        "e2e": {"value": value, "unit": "x realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def clips_of_rank(args, rank, world):
    """(number of streams of this rank, global id of its first stream)."""
    if args.total_clips:
        base, extra = divmod(args.total_clips, world)
        return base + (1 if rank < extra else 0), rank * base + min(rank, extra)
    return args.clips_per_gpu, rank * args.clips_per_gpu


def workload_config(args, world):
    if args.total_clips:
        what = ("C5 as written: %d independent 30 s 44.1 kHz stereo clips in all, -B64 (CBR 128 kbps), split over %d "
                "GPU(s) (%d per GPU); no collective" % (args.total_clips, world, -(-args.total_clips // world)))
        per_gpu = -(-args.total_clips // world)
    else:
        what = ("C5 shard: %d independent 30 s 44.1 kHz stereo clips per GPU, -B64 (CBR 128 kbps); streams sharded "
                "over GPUs, no collective" % args.clips_per_gpu)
        per_gpu = args.clips_per_gpu
    return {"workload": what, "clips_per_gpu": per_gpu, "total_clips": args.total_clips or per_gpu * world,
            "clip_seconds": CLIP_S, "samprate": SR, "channels": NCH, "options": "-B64",
            "input": "every stream distinct: seeded mix of two of %d synthetic base clips (seeds 10000..%d)"
                     % (N_BASE, 10000 + N_BASE - 1),
            "cache": "inputs (%.1f GB PCM per GPU) far larger than L2; no flush needed"
                     % (per_gpu * CLIP_N * NCH * 2 / 1e9)}


# ------------------------------------------------------------------------------------------------
# the GPU arm
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank):
    """Host feeder placement (one process per GPU): run on the cores next to the GPU so that the pinned PCM / output
    buffers are allocated on its NUMA node and the staging copies do not cross sockets.  Returns the core count."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def host_mem_available():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 0


def frame_match(a, b, frame_bytes=417):
    """Fraction of the reference's frames found byte-identical at the same position (1.0 when the streams are equal)."""
    if a.size == b.size and np.array_equal(a, b):
        return 1.0
    n = min(a.size, b.size) // frame_bytes
    if n == 0:
        return 0.0
    same = sum(np.array_equal(a[i * frame_bytes:(i + 1) * frame_bytes], b[i * frame_bytes:(i + 1) * frame_bytes])
               for i in range(n))
    return same / max(n, -(-max(a.size, b.size) // frame_bytes))


def run_gpu(args, rank, local_rank, world):
    import ctypes as C
    clips = base_clips()                           # before torch / CUDA exist in this process (fork pool)
    if world > 1:
        bind_to_gpu_numa_node(local_rank)
    import torch
    from hmp3_b200 import capi

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = local_rank
    B, first_stream = clips_of_rank(args, rank, world)
    # ---- synthetic input: every stream is a seeded mix of two base clips, built on the device and brought to ONE
    # pinned host buffer with a region per stream (the e2e leg reads all of it every step)
    need = B * CLIP_N * NCH * 2
    avail = host_mem_available()
    n_host = B
    if avail and need * (world if world > 1 else 1) > 0.7 * avail:   # all ranks share the host's memory
        n_host = max(64, int(0.7 * avail / (world * CLIP_N * NCH * 2)))
        n_host = min(n_host, B)
    pcm_pin = torch.empty((n_host, CLIP_N, NCH), dtype=torch.int16).pin_memory()
    base_dev = torch.stack([torch.from_numpy(c) for c in clips]).to("cuda").float()
    CH = 64
    for i0 in range(0, n_host, CH):
        blk = torch.empty((min(CH, n_host - i0), CLIP_N, NCH), dtype=torch.int16, device="cuda")
        for j in range(blk.shape[0]):
            ka, sa, ga, kb, sb, gb = stream_recipe(first_stream + i0 + j)
            blk[j] = torch.round(ga * base_dev[ka, sa:sa + CLIP_N] + gb * base_dev[kb, sb:sb + CLIP_N]).to(torch.int16)
        pcm_pin[i0:i0 + blk.shape[0]].copy_(blk)
    torch.cuda.synchronize()
    del base_dev, blk
    torch.cuda.empty_cache()
    ctl = [capi.control(samprate=SR, nch=NCH, bitrate=64)] * B
    plan = capi.Batch(ctl, [CLIP_N] * B, device=dev)
    stream_bytes = CLIP_N * NCH * 2
    pcm_ptrs = (pcm_pin.data_ptr() + (np.arange(B, dtype=np.uint64) % np.uint64(n_host)) * np.uint64(stream_bytes)
                ).astype(np.uint64)
    out_caps = plan.bound.copy()
    out_pin = torch.empty(int(out_caps.sum()), dtype=torch.uint8).pin_memory()
    out_offs = np.concatenate([[0], np.cumsum(out_caps)[:-1]]).astype(np.int64)
    out_ptrs = (out_pin.data_ptr() + out_offs).astype(np.uint64)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- make the PCM resident (not timed), then warm up
    for i in range(B):
        plan.upload_ptr(i, int(pcm_ptrs[i]), CLIP_N)
    plan.sync_stream()
    for _ in range(args.warmup):
        plan.run()
    nb, nf, off, st = plan.results()
    assert (st == 0).all(), "a stream failed"
    launches_per_step = plan.launches()
    audio_s_per_step = sum_over_ranks(B * CLIP_S)

    # ---- timed: K steps, inputs resident in HBM
    clocks = ClockSampler(dev)
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        plan.run()                                 # synchronous: returns when the stream has drained
        dev_ms += plan.last_run_ms()
    barrier()
    t_res = max_over_ranks(time.perf_counter() - t0)
    dev_ms = max_over_ranks(dev_ms)
    clk = clocks.stop()

    # ---- the serial stage's launch times inside the running pipeline (CUDA events on its stream), one extra step
    plan.set_timing(True)
    plan.run()
    phases = plan.phase_ms()
    timed_run_ms = plan.last_run_ms()
    rate_ms, rate_launches = phases["rate_loop"]
    # ---- every kernel alone: one more step with all kernels on one stream (uncontended launch times)
    plan.set_serialize(True)
    plan.run()
    alone = plan.phase_ms()
    alone_run_ms = plan.last_run_ms()
    plan.set_serialize(False)
    plan.set_timing(False)

    # ---- end to end through the host-buffer C-ABI entry: H2D + kernels + D2H every step
    for _ in range(2):
        plan.encode_host_ptrs(pcm_ptrs, out_ptrs, out_caps)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        nb2, nf2, st2 = plan.encode_host_ptrs(pcm_ptrs, out_ptrs, out_caps)
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    assert (st2 == 0).all() and np.array_equal(nb2, nb)
    h2d = sum_over_ranks(B * CLIP_N * NCH * 2)
    d2h = sum_over_ranks(int(nb2.sum()))
    frames_all = sum_over_ranks(int(nf.sum()))
    bytes_all = sum_over_ranks(int(nb.sum()))

    if rank == 0:
        # ---- parity of the bench shape itself: streams of this run against the oracle (the reference's own code
        # compiled into oracle/_ref/libhmp3ref.so), read from the buffers the timed e2e steps used
        parity = {"parity_checked": 0, "frame_match": None}
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import refmod
            if refmod.available():
                ec = refmod.make_ec(samprate=SR, nch=NCH, bitrate=64)
                picks = sorted(set(int(x) for x in np.linspace(0, B - 1, args.parity_streams)))
                fm = []
                for i in picks:
                    pcm = pcm_pin[i % n_host].numpy()
                    ref, _ = refmod.ref_encode_clip(ec, pcm)
                    got = out_pin[int(out_offs[i]):int(out_offs[i]) + int(nb2[i])].numpy()
                    fm.append(frame_match(got, ref))
                parity = {"parity_checked": len(picks), "frame_match": min(fm),
                          "parity_note": "streams %s of the timed e2e batch, byte-compared with the reference's own "
                                         "code (oracle/_ref/libhmp3ref.so) on the same PCM" % picks}
        except Exception as e:
            parity["parity_note"] = "oracle unavailable: %s" % e
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        fp32 = capi.fp32_peak(dev)
        NG = plan.chunk_granules()
        # granule-channels one launch of the serial stage processes, averaged over the step's launches (the first
        # chunk is shorter than the others): 2 granules per frame, NCH channels
        gc_per_launch = 2.0 * float(nf.sum()) * NCH / max(rate_launches, 1)
        bytes_per_launch = K6_BYTES_PER_GC * gc_per_launch
        avg_launch_s = rate_ms / max(rate_launches, 1) / 1e3
        achieved = bytes_per_launch / avg_launch_s / 1e9
        traffic_gc, traffic_src = rate_traffic_per_gc()
        roof = {"kernel": "k_rate_ph (serial stage of the rate loop, phase-scheduled: a block owns %d streams, its warps "
                          "claim streams that need the block's current phase)" % -(-B // 148),
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak,
                "traffic": args.rate_traffic if args.rate_traffic is not None else traffic_gc * gc_per_launch,
                "traffic_source": "ncu --set full capture (%s), scaled per granule-channel" % traffic_src,
                "peak_source": peak_src,
                "share_of_step": rate_ms / timed_run_ms,
                "share_note": "wall share of the step during which this kernel is running (Phase A and the packing "
                              "pass run concurrently on other streams)",
                "note": "latency / instruction-supply bound serial code, not a bandwidth kernel (DESIGN.md 4, 7.2): "
                        "its issue-slot utilisation from ncu is the meaningful fraction, see profiles/",
                "issue": rate_issue_counters()}
        # every kernel against its own ceiling, from the serialised step (launch times without overlap)
        roof_all = {}
        for name, (bound, per_gc) in KERNEL_WORK.items():
            ms, ln = alone.get(name, (0.0, 0))
            if ln == 0:
                continue
            if not ms > 0:   # (an event pair that could not be read; seen once for one kernel)
                roof_all[name] = {"bound": bound, "ms_per_step_alone": None, "launches": ln, "note": "not measured in this run"}
                continue
            halo = (NG + 3.0) / NG if name in ("polyphase", "attack") else 1.0   # these also redo a 3-granule halo per chunk
            work = per_gc * 2.0 * float(nf.sum()) * NCH * halo                     # per step
            rate = work / (ms * 1e-3)
            if bound == "hbm":
                roof_all[name] = {"bound": "hbm", "achieved": rate / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                  "frac": rate / 1e9 / hbm_peak, "ms_per_step_alone": ms, "launches": ln}
            else:
                pk = fp32["nonfused_tflops"] or 37.2
                roof_all[name] = {"bound": "fp32 issue, non-fused (exact order forbids FMA)", "achieved": rate / 1e12,
                                  "peak": pk, "unit": "TFLOP/s", "frac": rate / 1e12 / pk, "ms_per_step_alone": ms,
                                  "launches": ln, "ffma_peak": fp32["ffma_tflops"]}
        line = {
            "metric": METRIC, "value": audio_s_per_step * args.steps / t_res, "unit": "x realtime",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * t_res / args.steps, "device_ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.total_clips else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, world), "clocks": clk,
            "e2e": {"value": audio_s_per_step * args.steps / t_e2e, "unit": "x realtime",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "hmp3_batch_encode_host (C ABI, pinned host buffers in and out)",
                    "host_source_bytes": n_host * stream_bytes * world,
                    "host_source": "one pinned buffer, a region per stream" if n_host == B else
                                   "host memory short: %d distinct pinned stream regions per rank, reused" % n_host},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roof,
            "roofline_all": roof_all,
            "roofline_all_note": "per kernel: algorithmic work of one step / its launch times with every kernel on one "
                                 "stream (hmp3_batch_set_serialize; that step took %.1f ms against %.1f ms pipelined); "
                                 "fp32 ceilings measured live (hmp3_debug_fp32_peak)" % (alone_run_ms, timed_run_ms),
            "frames_per_step": frames_all, "bytes_out_per_step": bytes_all,
        }
        line.update(parity)
        if world == 1 and not args.no_cpu_baseline:
            cores = ref_cores()
            jobs = max(32, 16 * cores)
            try:
                rr = RefRunner(clips, jobs)
                rr.step()
                t = rr.step()
                rr.close()
                line["cpu_baseline"] = {"value": jobs * CLIP_S / t, "unit": "x realtime", "cores": cores,
                                        "kind": "reference", "sample": rr.sample()}
            except Exception as e:  # the baseline is a reported extra; never fail the GPU line for it
                line["cpu_baseline"] = {"value": None, "unit": "x realtime", "cores": cores, "kind": "reference",
                                        "sample": "unavailable: %s" % e}
        print(json.dumps(line))
    plan.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips-per-gpu", type=int, default=9472,
                    help="streams per GPU (default: 64 streams per SM for the phase-scheduled serial stage, 148 SMs; "
                         "BASELINE config 5 puts 10000 on one GPU)")
    ap.add_argument("--total-clips", type=int, default=0,
                    help="strong scaling: this many clips in all, split over the GPUs (BASELINE config 5: 10000)")
    ap.add_argument("--parity-streams", type=int, default=8,
                    help="streams of the timed batch compared with the oracle after the timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rate-traffic", type=float, default=None,
                    help="dram bytes per k_rate launch from the committed ncu capture (profiles/), if known")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
                                   "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:])
    run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
