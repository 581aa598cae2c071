#!/usr/bin/env python
"""Hottest SASS lines of a kernel from `ncu -i X.ncu-rep --page source --csv` (stall samples)."""
import csv
import subprocess
import sys


def main(path, top=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    i_src, i_all, i_ni, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), \
        hdr.index("Warp Stall Sampling (Not-issued Samples)"), hdr.index("Instructions Executed")
    body = [r for r in rows[2:] if len(r) > i_ex]
    total = sum(int(r[i_all] or 0) for r in body)
    print(f"total samples {total}, instructions {len(body)}")
    order = sorted(range(len(body)), key=lambda k: -int(body[k][i_all] or 0))[:top]
    for k in sorted(order):
        r = body[k]
        print(f"{k:5d} {100 * int(r[i_all] or 0) / total:5.1f}%  exec={r[i_ex]:>9s}  {r[i_src].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
