"""Where does the end-to-end step spend its time?  wall vs device time of hmp3_batch_encode_host."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hmp3_b200 import capi
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
clips = bench.base_clips()
pinned = [torch.from_numpy(c).pin_memory() for c in clips]
plan = capi.Batch([capi.control(bitrate=64)] * B, [bench.CLIP_N] * B)
ptrs = np.zeros(B, np.uint64)
for i in range(B):
    k, s0 = bench.stream_window(i)
    ptrs[i] = pinned[k].data_ptr() + s0 * 4
caps = plan.bound.copy()
out = torch.empty(int(caps.sum()), dtype=torch.uint8).pin_memory()
optrs = (out.data_ptr() + np.concatenate([[0], np.cumsum(caps)[:-1]])).astype(np.uint64)
for i in range(B):
    plan.upload_ptr(i, int(ptrs[i]), bench.CLIP_N)
plan.sync_stream()
for it in range(3):
    t0 = time.perf_counter(); plan.run(); t1 = time.perf_counter()
    print("resident run: wall %.3f s, device %.3f s" % (t1 - t0, plan.last_run_ms() / 1e3))
for it in range(3):
    t0 = time.perf_counter(); plan.encode_host_ptrs(ptrs, optrs, caps); t1 = time.perf_counter()
    print("encode_host : wall %.3f s, device %.3f s" % (t1 - t0, plan.last_run_ms() / 1e3))
