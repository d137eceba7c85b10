"""Correctness + timing of the wide-tile pass variants (GPU box).  Tuning tool, not part of the product.
Bit-exact against torch.sort of the unsigned images (and the CPU oracle at small sizes)."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import _oracle  # noqa: E402
import lsdradixsort_b200 as L  # noqa: E402
from lsdradixsort_b200 import keygen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--variants", type=str, default="0")
ap.add_argument("--no-typed", action="store_true", help="skip the f32 / i32 checks (tuning variants without the typed-key form)")
ap.add_argument("--log2n", type=int, default=28)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--skip-check", action="store_true")
ap.add_argument("--r", type=int, default=8)
args = ap.parse_args()
variants = [int(x) for x in args.variants.split(",")]


def usort(t):
    return torch.sort(t.to(torch.int64) & 0xFFFFFFFF).values.to(torch.int32)


def gpu_keys(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randint(-(2**31), 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)


bad = 0
if not args.skip_check:
    for v in variants:
        for kind in keygen.KINDS:
            for n in (1, 1000, 16799, 16800, 16801, 300_011, (1 << 22) + 5):
                keys = keygen.make_keys(kind, n, seed=v)
                d = torch.from_numpy(keys.view(np.int32)).cuda()
                try:
                    L.sort_(d, r=args.r, variant=v)
                except Exception as e:  # noqa: BLE001
                    print(json.dumps({"variant": v, "kind": kind, "n": n, "error": str(e)}), flush=True)
                    bad += 1
                    break
                got = d.cpu().numpy().view(np.uint32)
                want = _oracle.sort(keys, 8) if n <= 300_011 else np.sort(keys)
                if not np.array_equal(got, want):
                    bad += 1
                    idx = int(np.nonzero(got != want)[0][0])
                    print(json.dumps({"variant": v, "kind": kind, "n": n, "MISMATCH_at": idx}), flush=True)
        # multi-portion hand-off, typed keys, large uniform vs torch.sort
        keys = keygen.make_keys("entropy4_table", 150_001, seed=5)
        d = torch.from_numpy(keys.view(np.int32)).cuda()
        L.sort_(d, r=args.r, variant=v, portion_keys=40000)
        if not np.array_equal(d.cpu().numpy().view(np.uint32), np.sort(keys)):
            bad += 1
            print(json.dumps({"variant": v, "portions": "MISMATCH"}), flush=True)
        big = gpu_keys(1 << 26, v)
        want = usort(big)
        L.sort_(big, r=args.r, variant=v)
        if not bool((big == want).all()):
            bad += 1
            print(json.dumps({"variant": v, "n": 1 << 26, "MISMATCH": True}), flush=True)
        if args.no_typed:
            del big, want
            print(json.dumps({"variant": v, "checked": True, "bad_so_far": bad}), flush=True)
            continue
        f = torch.nan_to_num(gpu_keys(3_000_001, v + 1).view(torch.float32), nan=1.0)
        want_f = torch.sort(f).values
        L.sort_(f, r=args.r, variant=v)
        if not bool((f == want_f).all()):
            bad += 1
            print(json.dumps({"variant": v, "f32": "MISMATCH"}), flush=True)
        i = gpu_keys(2_000_003, v + 2)
        want_i = torch.sort(i).values
        L.sort_(i, r=args.r, variant=v, key_type="i32")
        if not bool((i == want_i).all()):
            bad += 1
            print(json.dumps({"variant": v, "i32": "MISMATCH"}), flush=True)
        del big, want, f, want_f, i, want_i
        print(json.dumps({"variant": v, "checked": True, "bad_so_far": bad}), flush=True)

n = 1 << args.log2n
src = gpu_keys(n, 0)
work = torch.empty_like(src)
want = usort(src) if not args.skip_check else None
for v in variants:
    try:
        s = L.Sorter(n, r=args.r, variant=v)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"variant": v, "error": str(e)}))
        continue
    best = None
    for _ in range(args.reps):
        work.copy_(src)
        st = s.sort_timed_(work)
        if best is None or sum(st) < sum(best):
            best = st
    ok = bool((work == want).all()) if want is not None else None
    total = sum(best)
    print(json.dumps({"variant": v, "log2n": args.log2n, "bit_exact_vs_torch_sort": ok, "stage_ms": [round(x, 4) for x in best],
                      "total_ms": round(total, 4), "gkeys_s": round(n / total / 1e6, 2),
                      "pass_frac": [round(8 * n / (x * 1e6) / 6551.0, 3) for x in best[1:-1]]}), flush=True)
    del s
print(json.dumps({"bad": bad}))
sys.exit(1 if bad else 0)
