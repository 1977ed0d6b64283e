"""Diagnostic: host<->device copy bandwidth of this box (pinned / pageable)."""
import time, torch
x = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for name, src, dst in (("h2d pinned", x, d), ("d2h pinned", d, x)):
    for _ in range(3):
        torch.cuda.synchronize(); t = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"{name}: {1.073741824 / dt:.1f} GB/s")
y = torch.empty(1 << 28, dtype=torch.uint8)
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); d[:1 << 28].copy_(y); torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"h2d pageable: {0.268435456 / dt:.1f} GB/s")
t = time.perf_counter(); z = torch.zeros(1 << 30, dtype=torch.float32, device="cuda"); torch.cuda.synchronize(); print("cudaMalloc+zero 4 GB:", time.perf_counter() - t)
del z; t = time.perf_counter(); z = torch.zeros(1 << 30, dtype=torch.float32, device="cuda"); torch.cuda.synchronize(); print("cached zero 4 GB:", time.perf_counter() - t)
