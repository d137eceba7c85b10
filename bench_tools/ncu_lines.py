"""Summarise an ncu report here (no GPU): key raw metrics + stall samples per CUDA source line."""
import csv
import subprocess
import sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum']
for r in rows[2:]:
    print('=' * 100)
    for w in want:
        if w in idx:
            print(f"{w:70s} {r[idx[w]][:60]:>24s} {units[idx[w]]}")
    for h in hdr:
        if 'smsp__average_warp' in h and 'issue_stalled' in h and 'not_issued' not in h:
            try:
                v = float(r[idx[h]])
            except ValueError:
                continue
            if v > 0.3:
                print(f"   {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:8.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, agg, order = None, {}, []
for r in csv.reader(src.splitlines()):
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) < 8 or not r[0].isdigit():
        continue
    try:
        samp, inst = int(r[6]), int(r[7])
    except ValueError:
        continue
    k = (cur, int(r[0]), r[1].strip()[:100])
    if k not in agg:
        agg[k] = [0, 0]
        order.append(k)
    agg[k][0] += samp
    agg[k][1] += inst
ts = sum(v[0] for v in agg.values()) or 1
ti = sum(v[1] for v in agg.values()) or 1
print(f"total samples {ts}, warp instructions {ti}")
for k in sorted(order, key=lambda k: (k[0], k[1])):
    s, i = agg[k]
    if 100 * s / ts > thr or 100 * i / ti > thr:
        print(f"{k[0][:16]:16s} {k[1]:4d} samp {100*s/ts:5.1f}% inst {100*i/ti:5.1f}% | {k[2]}")
