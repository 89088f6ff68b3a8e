import torch, time
dev = "cuda:0"
for mb in (1, 5, 20, 80):
    n = mb * 1024 * 1024 // 8
    h = torch.empty(n, dtype=torch.float64).pin_memory(); d = torch.empty(n, dtype=torch.float64, device=dev)
    h2 = torch.empty(n, dtype=torch.float64).pin_memory(); d2 = torch.empty(n, dtype=torch.float64, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize(); ev[0].record()
        for _ in range(10): fn()
        ev[1].record(); torch.cuda.synchronize()
        print("%s %3d MB: %.1f GB/s" % (name, mb, 10 * n * 8 / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("duplex %3d MB: %.1f GB/s each way" % (mb, 10 * n * 8 / dt / 1e9))
