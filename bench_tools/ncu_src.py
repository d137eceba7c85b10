"""Aggregate an `ncu --page source --csv` dump by SASS opcode class: shared wavefronts, global sectors, stall samples.
usage: python bench_tools/ncu_src.py dump.csv"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0, 0, 0, 0, 0])
tot_samples = 0
per = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"):
        op = src.split()[1]
    def f(name):
        try:
            return float(r[ix[name]] or 0)
        except ValueError:
            return 0.0
    inst = f("Instructions Executed")
    wf = f("L1 Wavefronts Shared")
    wfi = f("L1 Wavefronts Shared Ideal")
    sec = f("L2 Theoretical Sectors Global")
    tag = f("L1 Tag Requests Global")
    smp = f("# Samples")
    a = agg[op]
    a[0] += inst; a[1] += wf; a[2] += wfi; a[3] += sec; a[4] += tag; a[5] += smp
    tot_samples += smp
    per.append((wf, sec, smp, r[ix["Address"]], src[:70], inst))
print(f"{'opcode':28s} {'inst':>12s} {'smem wavefr':>12s} {'ideal':>12s} {'L2 sectors':>12s} {'L1 tags':>10s} {'samples%':>8s}")
for op, a in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][3])):
    if a[1] + a[3] + a[4] == 0 and a[5] < 0.01 * tot_samples:
        continue
    print(f"{op:28s} {a[0]:12.0f} {a[1]:12.0f} {a[2]:12.0f} {a[3]:12.0f} {a[4]:10.0f} {100*a[5]/max(tot_samples,1):8.1f}")
print("\ntop instructions by shared wavefronts:")
for wf, sec, smp, addr, src, inst in sorted(per, key=lambda t: -t[0])[:12]:
    print(f"  {addr:>8s} wf {wf:10.0f} inst {inst:9.0f} ({wf/max(inst,1):.2f}/inst)  {src}")
print("top instructions by stall samples:")
for wf, sec, smp, addr, src, inst in sorted(per, key=lambda t: -t[2])[:12]:
    print(f"  {addr:>8s} samples {100*smp/max(tot_samples,1):5.1f}%  {src}")
