"""Where do the ~3.5 ms between lsd_sort_host (44.4 ms) and H2D + sort + D2H measured alone (19.3 + 2.8 + 18.8) go?"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import lsdradixsort_b200 as L
from lsdradixsort_b200 import keygen
n = 1 << 28
src = torch.from_numpy(keygen.uniform_u32(n, 0).view(np.int32)).pin_memory()
buf = torch.empty_like(src).pin_memory()
hs = L.HostSorter(n, r=8)
def run(refill, label):
    ts = []
    for i in range(6):
        if refill or i == 0:
            buf.copy_(src)
        if refill == 2:
            time.sleep(0.5)
        t0 = time.perf_counter(); hs.sort_(buf); ts.append(time.perf_counter() - t0)
    print(label, "ms per call:", [round(1e3 * t, 2) for t in ts[1:]])
run(1, "input refilled by the host right before every call")
run(2, "refilled, then 0.5 s pause               ")
run(0, "no refill (the buffer holds the previous result)")
d = torch.empty(n, dtype=torch.int32, device="cuda")
a, b, c, e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
s = L.Sorter(n, r=8)
for i in range(3):
    buf.copy_(src)
    a.record(); d.copy_(buf, non_blocking=True); b.record(); s.sort_(d); c.record(); buf.copy_(d, non_blocking=True); e.record()
    torch.cuda.synchronize()
    print("torch copies + sort: H2D %.2f ms, sort %.2f ms, D2H %.2f ms" % (a.elapsed_time(b), b.elapsed_time(c), c.elapsed_time(e)))
