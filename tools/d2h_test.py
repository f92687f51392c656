import torch, time
d = torch.empty(2 << 30, dtype=torch.uint8, device="cuda")
for size in (64 << 20, 256 << 20, 1 << 30, 2 << 30):
    h = torch.empty(size, dtype=torch.uint8, pin_memory=True)
    for _ in range(2):
        h.copy_(d[:size], non_blocking=True); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5):
        h.copy_(d[:size], non_blocking=True)
    torch.cuda.synchronize()
    print(f"D2H {size >> 20} MiB: {5 * size / (time.perf_counter() - t) / 1e9:.1f} GB/s")
    del h
# two concurrent streams
h1 = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True); h2 = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): h1.copy_(d[:1 << 30], non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d[1 << 30:], non_blocking=True)
torch.cuda.synchronize()
print(f"D2H 2 streams: {10 * (1 << 30) / (time.perf_counter() - t) / 1e9:.1f} GB/s")
