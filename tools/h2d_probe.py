import torch, time
for mb in (16, 64, 256):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"H2D pinned {mb} MB: {mb / 1024 / dt:.1f} GB/s")
    t0 = time.perf_counter()
    for _ in range(10): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"D2H pinned {mb} MB: {mb / 1024 / dt:.1f} GB/s")
t0 = time.perf_counter(); x = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); print("pin 64MB ms", 1e3 * (time.perf_counter() - t0))
