"""Summarise an `ncu --set full --import-source on` report: key raw metrics per kernel plus the hottest SASS
regions (consecutive instructions with similar execution counts) from the source page.

  python profiles/ncu_summarize.py report.ncu-rep > profiles/<name>.txt
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    kernels = rows[2:]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    heads = [i for i, r in enumerate(srows) if "Instructions Executed" in r]
    # the source page lists every kernel twice (SASS view, then source view); keep the SASS views
    sass_heads = heads[0::2] if len(heads) >= 2 * len(kernels) else heads
    for k, vals in enumerate(kernels):
        name = vals[hdr.index("Kernel Name")]
        print(f"==== kernel {k}: {name}")
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                print(f"{h} = {v} {u}")
        if k >= len(sass_heads):
            continue
        h0 = sass_heads[k]
        nxt = [h for h in heads if h > h0]
        end = (nxt[0] - 1) if nxt else len(srows)
        shdr = srows[h0]
        data = [r for r in srows[h0 + 1:end] if len(r) == len(shdr)]
        ia, ta, sa = shdr.index("Instructions Executed"), shdr.index("Thread Instructions Executed"), shdr.index("# Samples")
        tot = sum(int(r[ia]) for r in data)
        ts = max(1, sum(int(r[sa]) for r in data))
        print(f"total warp instructions {tot}, SASS lines {len(data)}")
        regions, cur = [], None
        for i, r in enumerate(data):
            c, t, s = int(r[ia]), int(r[ta]), int(r[sa])
            if cur and c > 0 and 0.7 < c / max(cur["c0"], 1) < 1.4:
                cur["n"] += 1; cur["inst"] += c; cur["thr"] += t; cur["smp"] += s; cur["end"] = i
            else:
                if cur:
                    regions.append(cur)
                cur = dict(start=i, end=i, n=1, c0=c, inst=c, thr=t, smp=s)
        regions.append(cur)
        for r in sorted([r for r in regions if r["inst"] > tot * 0.01], key=lambda r: -r["inst"])[:14]:
            print(f"  sass[{r['start']:4d}..{r['end']:4d}] n={r['n']:3d} exec={r['c0']:>9d} inst={r['inst'] / tot * 100:5.1f}% "
                  f"threads/inst={r['thr'] / max(r['inst'], 1):5.1f} samples={r['smp'] / ts * 100:5.1f}%  {data[r['start']][1].strip()[:50]}")


if __name__ == "__main__":
    main()
