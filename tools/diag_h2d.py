import time, torch
x = torch.randn(470, 288, 512, 3).pin_memory()
print("pinned", x.is_pinned(), x[0:32].is_pinned())
dev = torch.device("cuda")
torch.cuda.synchronize()
for rep in range(3):
    ts = []
    t0 = time.perf_counter()
    for k in range(0, 448, 32):
        t = time.perf_counter()
        d = x[k:k+32].to(dev, dtype=torch.float32, non_blocking=True)
        ts.append(time.perf_counter() - t)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"rep {rep}: enqueue {1e3*(t1-t0):.1f} ms (per chunk {[round(1e3*v,2) for v in ts[:6]]}), drain {1e3*(t2-t1):.1f} ms, GB/s {x[:448].numel()*4/(t2-t0)/1e9:.1f}")
stage = torch.empty(32, 288, 512, 3, device=dev)
for rep in range(2):
    t0 = time.perf_counter()
    for k in range(0, 448, 32):
        stage.copy_(x[k:k+32], non_blocking=True)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"copy_ into one staging buffer: enqueue {1e3*(t1-t0):.1f} ms, drain {1e3*(t2-t1):.1f} ms")
