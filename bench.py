#!/usr/bin/env python3
"""Throughput benchmark of the B200-native Helix-MP3 hot path (BASELINE.json metric).

metric   encoded audio seconds per second (x realtime), 44.1 kHz stereo CBR128
step     one pass of the whole hot path (polyphase -> ... -> packed frames) over one batch of synthetic
         clips: --clips-per-gpu clips of 30 s per GPU (the C5 workload, sharded: no collective, weak scaling)
value    whole-job throughput with the PCM already resident in HBM
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
BASE_S = 32.0                      # base clips are a little longer; streams are shifted 30 s windows of them
N_BASE = 16
SHIFT = 563                        # samples between windows (not a multiple of 576: different framing)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "hmp3")
METRIC = "encoded audio sec/sec (x realtime), 44.1k stereo CBR128"

# algorithmic bytes of the serial stage per granule-channel (DESIGN.md "K6"): MDCT magnitudes in (2304 B) +
# prepared |x|^(3/4), signs, band energies and step bounds in (2816 B) + sig/mask in (288 B) + the record for the
# packing pass out (1396 B)
K6_BYTES_PER_GC = 2304 + 2816 + 288 + 1396
# DRAM traffic of the same kernel per granule-channel, from the committed `ncu --set full` capture
# (profiles/r1h_rate_ncu_details.txt: dram read 24.749 GB + write 25.116 GB for a launch of 4736 streams x 256
# granules x 2 channels); per-launch traffic = this x the granule-channels one launch processes
K6_NCU_DRAM_BYTES_PER_GC = (24.749032e9 + 25.115575e9) / (4736 * 256 * 2)


def base_clips():
    from hmp3_b200.synth import synth_pcm
    return [synth_pcm(10000 + i, BASE_S, SR, NCH) for i in range(N_BASE)]


def stream_window(i):
    """(base clip index, start sample) of stream i."""
    nshift = int((BASE_S - CLIP_S) * SR) // SHIFT
    return i % N_BASE, ((i // N_BASE) % nshift) * SHIFT


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
    """Encodes `jobs` 30 s clips with oracle/_ref/hmp3 -B64, `cores` processes at a time."""

    def __init__(self, clips, jobs):
        if not os.path.exists(REF_BIN):
            raise RuntimeError("oracle/_ref/hmp3 is missing (build it with __graft_entry__.build() where "
                               "/root/reference exists)")
        self.cores = ref_cores()
        self.jobs = jobs
        shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
        self.dir = tempfile.mkdtemp(prefix="hmp3_ref_", dir=shm)
        self.wavs = []
        for i in range(min(len(clips), jobs)):
            p = os.path.join(self.dir, "c%d.wav" % i)
            write_wav(p, clips[i][:CLIP_N])
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
    sample = "%d clips of 30 s per step (%d distinct), one `hmp3 in.wav out.mp3 -B64` process per clip, " \
             "%d at a time, files on tmpfs" % (jobs, len(rr.wavs), cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "x realtime", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "x realtime", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "x realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def workload_config(args):
    return {"workload": "C5 shard: %d independent 30 s 44.1 kHz stereo clips per GPU, -B64 (CBR 128 kbps); "
                        "streams sharded over GPUs, no collective" % args.clips_per_gpu,
            "clips_per_gpu": args.clips_per_gpu, "clip_seconds": CLIP_S, "samprate": SR, "channels": NCH,
            "options": "-B64", "cache": "inputs (%.1f GB PCM per GPU) far larger than L2; no flush needed"
                                         % (args.clips_per_gpu * CLIP_N * NCH * 2 / 1e9)}


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


def run_gpu(args, rank, local_rank, world):
    import ctypes as C
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
    B = args.clips_per_gpu
    # ---- synthetic input: pinned base clips; every stream is a shifted 30 s window of one of them
    clips = base_clips()
    pinned = [torch.from_numpy(c).pin_memory() for c in clips]
    ctl = [capi.control(samprate=SR, nch=NCH, bitrate=64)] * B
    plan = capi.Batch(ctl, [CLIP_N] * B, device=dev)
    first_stream = rank * B                       # global stream ids: ranks take disjoint windows
    pcm_ptrs = np.zeros(B, np.uint64)
    for i in range(B):
        k, s0 = stream_window(first_stream + i)
        pcm_ptrs[i] = pinned[k].data_ptr() + s0 * NCH * 2
    out_caps = plan.bound.copy()
    out_pin = torch.empty(int(out_caps.sum()), dtype=torch.uint8).pin_memory()
    out_ptrs = (out_pin.data_ptr() + np.concatenate([[0], np.cumsum(out_caps)[:-1]])).astype(np.uint64)

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

    # ---- make the PCM resident (not timed), then warm up
    for i in range(B):
        plan.upload_ptr(i, int(pcm_ptrs[i]), CLIP_N)
    plan.sync_stream()
    for _ in range(args.warmup):
        plan.run()
    nb, nf, off, st = plan.results()
    assert (st == 0).all(), "a stream failed"
    launches_per_step = plan.launches()
    audio_s_per_step = B * CLIP_S * world

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

    # ---- per-kernel device times (CUDA events on the launching stream), one extra untimed step
    plan.set_timing(True)
    plan.run()
    phases = plan.phase_ms()
    timed_run_ms = plan.last_run_ms()
    plan.set_timing(False)
    rate_ms, rate_launches = phases["rate_loop"]

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
    h2d = B * CLIP_N * NCH * 2
    d2h = int(nb2.sum())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # granule-channels one launch of the serial stage processes, averaged over the step's launches (the first
        # chunk is shorter than the others): 2 granules per frame, NCH channels
        gc_per_launch = 2.0 * float(nf.sum()) * NCH / max(rate_launches, 1)
        bytes_per_launch = K6_BYTES_PER_GC * gc_per_launch
        avg_launch_s = rate_ms / max(rate_launches, 1) / 1e3
        achieved = bytes_per_launch / avg_launch_s / 1e9
        roof = {"kernel": "k_rate (serial stage of the rate loop, one warp per stream)",
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak,
                "traffic": args.rate_traffic if args.rate_traffic is not None
                else K6_NCU_DRAM_BYTES_PER_GC * gc_per_launch,
                "traffic_source": "ncu --set full capture in profiles/r1h_rate_ncu_details.txt, scaled per granule-channel",
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "share_of_step": rate_ms / timed_run_ms,
                "share_note": "wall share of the step during which this kernel is running (Phase A and the packing "
                              "pass run concurrently on other streams; serialised share in profiles/r1h_launch_summary.txt: 77 %)",
                "note": "latency/instruction-fetch bound serial code, not a bandwidth kernel: see DESIGN.md"}
        line = {
            "metric": METRIC, "value": audio_s_per_step * args.steps / t_res, "unit": "x realtime",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * t_res / args.steps, "device_ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args), "clocks": clk,
            "e2e": {"value": audio_s_per_step * args.steps / t_e2e, "unit": "x realtime",
                    "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "api": "hmp3_batch_encode_host (C ABI, pinned host buffers in and out)"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roof,
            "kernels_ms_per_step": {k: round(v[0], 3) for k, v in phases.items()},
            "kernels_note": "CUDA-event elapsed per kernel, summed over launches; kernels on different streams overlap, "
                            "so the entries do not add up to ms_per_step",
            "frames_per_step": int(nf.sum()) * world, "bytes_out_per_step": int(nb.sum()) * world,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = ref_cores()
            jobs = max(32, 16 * cores)
            try:
                rr = RefRunner(clips, jobs)
                rr.step()
                t = rr.step()
                rr.close()
                line["cpu_baseline"] = {
                    "value": jobs * CLIP_S / t, "unit": "x realtime", "cores": cores, "kind": "reference",
                    "sample": "%d clips of 30 s (%d distinct) through oracle/_ref/hmp3 -B64, one process per "
                              "clip, %d at a time, files on tmpfs" % (jobs, len(rr.wavs), cores)}
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
    ap.add_argument("--clips-per-gpu", type=int, default=4736,
                    help="streams per GPU (default: one full wave of the serial-stage kernel, 148 SMs x 32 warps)")
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
