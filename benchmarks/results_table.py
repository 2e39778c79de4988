"""Markdown tables of BASELINE.md section 5 from the JSON lines under profiles/<round>/.
Usage: python benchmarks/results_table.py [profiles/r01]"""
import json
import os
import sys


def load(path):
    try:
        return json.loads(open(path).read().strip().splitlines()[-1])
    except (OSError, ValueError, IndexError):
        return None


def main(d="profiles/r01"):
    print("| Config | CPU reference port (cores) | B200 x1 `value` (device env) | B200 x1 `e2e` (host env) | ms / learn_epoch | in-step roofline of the dominant kernel |")
    print("|---|---|---|---|---|---|")
    for c in ("C1", "C2", "C3", "C4", "C5"):
        b, r = load(os.path.join(d, f"b200_{c}.json")), load(os.path.join(d, f"ref_{c}.json"))
        if not b:
            continue
        rf = b.get("roofline") or {}
        roof = f"{rf.get('frac', 0):.2f} ({rf.get('achieved', 0):.0f} GB/s, {rf.get('ms_per_launch', 0) * 1e3:.0f} us / launch)" if rf else "n/a"
        ref = f"{r['value']:.0f} ({r['cpu_baseline']['cores']})" if r else "n/a"
        print(f"| {b['config']['workload'][:60]} | {ref} | {b['value']:.0f} | {b['e2e']['value']:.0f} | {b['ms_per_step']:.1f} | {roof} |")
    for n in (2, 4, 8):
        b = load(os.path.join(d, f"b200_C4_n{n}.json"))
        if b:
            print(f"\nC4 at N={n}: value {b['value']:.0f} env-steps/s, e2e {b['e2e']['value']:.0f}, {b['ms_per_step']:.1f} ms / learn_epoch")
    k = None
    for name in ("kernels_latest.json",):
        p = os.path.join(d, name)
        if os.path.exists(p):
            k = json.load(open(p))
    if k:
        print("\n| Kernel | Shape | Time (median) | Algorithmic GB/s | of measured peak |")
        print("|---|---|---|---|---|")
        for row in k:
            shape = ", ".join(f"{a}={row[a]}" for a in ("T", "N", "V", "B", "HW", "S", "M", "unit_p", "dtype") if a in row)
            print(f"| {row['kernel']} | {shape} | {row['ms_median'] * 1e3:.1f} us | {row['gbs']:.0f} | {row['frac']:.3f} |")


if __name__ == "__main__":
    main(*sys.argv[1:2])
