"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into kernel / launches / total ns / share.
Usage: python benchmarks/launch_summary.py <launches.csv> <out.csv> "<title>" """
import csv
import sys
from collections import defaultdict


def main(src, out, title):
    rows = [r for r in csv.reader(open(src, errors="replace")) if r and not r[0].startswith("==")]
    hdr = rows[0]
    i_name, i_metric, i_unit, i_val = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) <= i_val or r[i_metric] != "gpu__time_duration.sum":
            continue
        name = r[i_name][:90]
        tot[name] += float(r[i_val].replace(",", "")) * scale.get(r[i_unit], 1.0)
        cnt[name] += 1
    total = sum(tot.values())
    ours = sum(v for k, v in tot.items() if "b200rl" in k)
    lines = [f"# {title}",
             "# metric gpu__time_duration.sum (ns), --clock-control none; per-launch times are cold-cache and serialised: compare SHARES",
             f"# total {total:.0f} ns over {sum(cnt.values())} launches; libb200rl kernels {ours:.0f} ns = {ours / max(total, 1):.4f} of it",
             "kernel,launches,total_ns,share"]
    for k in sorted(tot, key=lambda k: -tot[k]):
        lines.append(f"\"{k}\",{cnt[k]},{tot[k]:.0f},{tot[k] / total:.4f}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:30]))


if __name__ == "__main__":
    main(*sys.argv[1:4])
