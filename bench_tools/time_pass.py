import sys, torch
sys.path.insert(0, '.')
import lsdradixsort_b200 as L
n = 1 << 28
g = torch.Generator(device="cuda").manual_seed(0)
d = torch.randint(-(2**31), 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
out = torch.empty_like(d)
for r, bg in ((8, 3), (8, 0), (4, 7)):
    L.sort_pass(d, out, r, bg)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        a.record(); L.sort_pass(d, out, r, bg); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print("lsd_sort_pass r", r, "digit", bg, "ms", round(min(ts), 4))
