import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from qpth_b200 import QPFunction
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
B, n = 128, 100
f = QPFunction(verbose=-1, check_Q_spd=False)
e = torch.Tensor().to(dev).double(); dl = torch.ones(B, n, dtype=torch.float64, device=dev)
def loop(ncopies, steps, sampler=False, warm=5):
    batches = bench.make_batches(dev, 0, ncopies)
    def step(i):
        t = batches[i % ncopies]
        for v in t.values(): v.grad = None
        z = f(t["Q"], t["p"], t["G"], t["h"], e, e); z.backward(dl)
    for i in range(warm): step(i)
    torch.cuda.synchronize()
    s = bench.ClockSampler(0) if sampler else None
    if s: s.start()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); ev0.record()
    for i in range(steps): step(warm + i)
    ev1.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    if s: s.stop()
    print("ncopies %d sampler %s warm %d: GPU %.3f ms/step, host enqueue %.3f, wall %.3f | device allocs %d reserved %.0f MB" % (ncopies, sampler, warm, ev0.elapsed_time(ev1) / steps, (t1 - t0) / steps * 1e3, (t2 - t0) / steps * 1e3, torch.cuda.memory_stats()["num_device_alloc"], torch.cuda.memory_reserved() / 1e6))
loop(1, 30); loop(8, 30); loop(8, 30, warm=16); loop(8, 30, sampler=True, warm=16); loop(1, 30, sampler=True)
