import torch, time
n = 157 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for streams in (1, 4, 8):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    chunk = n // streams
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        for i, s in enumerate(ss):
            with torch.cuda.stream(s):
                d[i * chunk:(i + 1) * chunk].copy_(h[i * chunk:(i + 1) * chunk], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("H2D streams", streams, "%.1f GB/s" % (n / dt / 1e9))
torch.cuda.synchronize(); t = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize()
print("D2H %.1f GB/s" % (n / (time.perf_counter() - t) / 1e9))
# bidirectional
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2): h2.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("bidir each %.1f GB/s" % (n / dt / 1e9))
