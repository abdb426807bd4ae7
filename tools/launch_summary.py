"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the last pass.

  python tools/launch_summary.py gpurun_out/launches.csv [first_kernel_regex]
"""
import csv
import re
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        try:
            rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
        except (KeyError, ValueError):
            pass
    half = rows[len(rows) // 2:]
    total = sum(v for _, v in half)
    for k, v in half:
        print(f"  {k[:64]:64s} {v / 1000:9.1f} us")
    print(f"  total {total / 1000:.1f} us over {len(half)} launches")


if __name__ == "__main__":
    main()
