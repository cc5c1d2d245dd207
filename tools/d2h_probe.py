import torch, time
n = 307_198_736
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for sz in (n, n // 3, n // 13, n // 39):
    best = 1e9
    for _ in range(6):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        off = 0
        while off + sz <= n:
            h[off:off + sz].copy_(d[off:off + sz], non_blocking=True)
            off += sz
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"D2H pinned, pieces of {sz/1e6:.1f} MB: {off/1e6:.0f} MB in {best:.3f} ms -> {off/best/1e6:.1f} GB/s")
