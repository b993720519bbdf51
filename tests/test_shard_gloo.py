"""N>1 path on CPU: two gloo ranks shard a batch, each encodes its shard (with the host build of the kernel
bodies standing in for the device), results are gathered and compared with the single-process encode."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import simmod
from hmp3_b200 import capi, shard
from hmp3_b200.synth import synth_pcm


def test_assign_balances_and_covers():
    rng = np.random.default_rng(0)
    d = rng.uniform(1, 300, size=1001).tolist()
    for world in (1, 2, 4, 8):
        sh = shard.assign(d, world)
        assert sorted(i for s in sh for i in s) == list(range(len(d)))
        tot = [sum(d[i] for i in s) for s in sh]
        assert max(tot) - min(tot) <= max(d)
    assert shard.assign([], 4) == [[], [], [], []]


def _batch():
    specs = [(44100, 2, dict(bitrate=64), 1.3), (44100, 2, dict(bitrate=64), 0.7), (32000, 2, dict(), 1.0),
             (22050, 1, dict(bitrate=32), 1.5), (44100, 2, dict(bitrate=64), 0.0), (48000, 2, dict(vbr_mnr=100), 0.9)]
    ctl, pcms = [], []
    for k, (sr, nch, kw, secs) in enumerate(specs):
        p = synth_pcm(500 + k, max(secs, 0.01), sr, nch)
        pcms.append(p if secs > 0 else p[:0])
        ctl.append(capi.control(samprate=sr, nch=nch, **kw))
    return ctl, pcms


def _sim_encode(controls, pcms):
    return [simmod.encode_clip(c, p)[0] for c, p in zip(controls, pcms)]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctl, pcms = _batch()

    def gather(local):
        out = [None] * world
        dist.all_gather_object(out, local)
        return out

    outs = shard.encode_sharded(ctl, pcms, rank, world, _sim_encode, gather)
    q.put((rank, [hashlib.md5(o.tobytes()).hexdigest() for o in outs]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not simmod.available(), reason="host simulator not built")
def test_two_ranks_reproduce_the_single_process_encode():
    ctl, pcms = _batch()
    want = [hashlib.md5(o.tobytes()).hexdigest() for o in _sim_encode(ctl, pcms)]
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == want and got[1] == want
