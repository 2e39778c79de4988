"""Per-source-line instruction / stall-sample totals of one captured kernel.
Usage: python benchmarks/ncu_lines.py <report.ncu-rep> [top] [kernel-name regex]
(reads `ncu -i ... --page source --print-source cuda,sass --csv`)"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main(rep, top=40, kernel=None):
    cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]
    if kernel:
        cmd += ["-k", "regex:" + kernel]
    raw = subprocess.run(cmd, capture_output=True, text=True).stdout
    inst, samp, text = defaultdict(int), defaultdict(int), {}
    path, hdr = None, None
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            path = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            key = (path, int(r[0]))
            text.setdefault(key, r[1].strip())
            if r[2]:  # a SASS row under this line
                inst[key] += int(r[i_inst] or 0)
                samp[key] += int(r[i_samp] or 0)
    ti, ts = sum(inst.values()), sum(samp.values())
    print(f"total warp instructions {ti}, stall samples {ts}")
    for key in sorted(inst, key=lambda k: -samp[k])[: int(top)]:
        print(f"{samp[key] / max(ts, 1):6.3f} samp {inst[key] / max(ti, 1):6.3f} inst  {key[0]}:{key[1]:<4d} {text[key][:100]}")


if __name__ == "__main__":
    main(*sys.argv[1:4])
