"""Full vs lower-triangle (strided 3-D strips) transfers of a C2 batch of symmetric matrices, one direction and duplex."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qpth_b200.util import copy_lower_
dev = "cuda:0"
B, n = 128, 100
h = torch.randn(B, n, n, dtype=torch.float64); h = (h + h.transpose(1, 2)).contiguous().pin_memory()
d = torch.zeros(B, n, n, dtype=torch.float64, device=dev)
h2 = torch.zeros(B, n, n, dtype=torch.float64).pin_memory()
d2 = torch.randn(B, n, n, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn_a, fn_b=None, reps=20):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize(); e0.record(); s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(reps):
        with torch.cuda.stream(s1): fn_a()
        if fn_b is not None:
            with torch.cuda.stream(s2): fn_b()
    e1.record(s1); e2.record(s2); torch.cuda.synchronize()
    return max(e0.elapsed_time(e1), e0.elapsed_time(e2)) / reps * 1e3
# correctness
copy_lower_(d, h, band=20); torch.cuda.synchronize()
assert torch.equal(torch.tril(d.cpu()), torch.tril(h)), "lower triangle mismatch"
print("full  H2D %.0f us   D2H %.0f us   duplex %.0f us" % (timed(lambda: d.copy_(h, non_blocking=True)), timed(lambda: h2.copy_(d2, non_blocking=True)),
      timed(lambda: d.copy_(h, non_blocking=True), lambda: h2.copy_(d2, non_blocking=True))))
for band in (10, 20, 25, 50):
    frac = sum(min(n, r0 + band) * (min(n, r0 + band) - r0) for r0 in range(0, n, band)) / (n * n)
    print("band %2d (%.1f %% of the bytes): H2D %.0f us   D2H %.0f us   duplex %.0f us" % (band, 100 * frac,
          timed(lambda: copy_lower_(d, h, band)), timed(lambda: copy_lower_(h2, d2, band)),
          timed(lambda: copy_lower_(d, h, band), lambda: copy_lower_(h2, d2, band))))
