import torch, time
x = torch.empty(2 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty(2 << 30, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for name, piece in [("2 GiB in one copy", 2 << 30), ("4 MiB pieces", 4 << 20), ("530 KB pieces", 530 * 1024), ("128 KB pieces", 128 << 10)]:
    n = (2 << 30) // piece
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(n):
            d[i * piece:(i + 1) * piece].copy_(x[i * piece:(i + 1) * piece], non_blocking=True)
        t_enq = time.perf_counter() - t0
        torch.cuda.synchronize(); t1 = time.perf_counter() - t0
    print("%-20s H2D %6.1f GB/s (enqueue %.3f s, total %.3f s, %d copies)" % (name, (n * piece) / t1 / 1e9, t_enq, t1, n))
y = torch.empty(2 << 30, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter(); y.copy_(d, non_blocking=True); torch.cuda.synchronize()
print("D2H 2 GiB: %.1f GB/s" % ((2 << 30) / (time.perf_counter() - t0) / 1e9))
