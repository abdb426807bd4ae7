"""Per-function and per-line breakdown of one kernel from an `ncu --set full --import-source on` report:
executed warp instructions and stall samples bucketed by the source function a SASS instruction came from.

  python tools/ncu_by_function.py report.ncu-rep [top_lines]
"""
import bisect
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def functions(path):
    """[(first line, name)] of the top-level functions of a source file (good enough for this code base)."""
    out = []
    pending = None
    for i, line in enumerate(open(path, errors="replace"), 1):
        if line[:1] in " \t#/}\n" or line.startswith("//"):
            if pending and "(" in line:
                pass
            continue
        if re.match(r"^(template|CG_HD|__global__|__device__|static|inline|extern|struct|class|typedef|namespace|using)", line) or \
                re.match(r"^[A-Za-z_][\w:<>\*& ]*\(", line):
            m = re.search(r"([A-Za-z_]\w*)\s*\(", line)
            if line.startswith("template") and not m:
                pending = i
                continue
            if line.startswith(("struct", "class")):
                m2 = re.match(r"^(?:struct|class)\s+(\w+)", line)
                out.append((i, m2.group(1) if m2 else "?"))
            elif m:
                out.append((pending or i, m.group(1)))
            pending = None
    return out


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    cur = None
    inst, samp, src = collections.Counter(), collections.Counter(), {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1]
            continue
        if len(r) < 8 or r[0] in ("Line No", ""):
            continue
        try:
            ln, ex, sm = int(r[0]), int(r[7]), int(r[4])
        except ValueError:
            continue
        inst[(cur, ln)] += ex
        samp[(cur, ln)] += sm
        src[(cur, ln)] = r[1].strip()
    T, S = max(1, sum(inst.values())), max(1, sum(samp.values()))
    print(f"warp instructions {T}, stall samples {S}")
    cache = {}

    def fn(path, ln):
        local = os.path.join(ROOT, "cutadapt_b200", "csrc", os.path.basename(path))
        if not os.path.exists(local):
            return os.path.basename(path)
        if local not in cache:
            cache[local] = functions(local)
        f = cache[local]
        i = bisect.bisect_right([x[0] for x in f], ln) - 1
        return f[i][1] if i >= 0 else "?"

    byf_i, byf_s = collections.Counter(), collections.Counter()
    for (p, ln), v in inst.items():
        key = (os.path.basename(p), fn(p, ln))
        byf_i[key] += v
        byf_s[key] += samp[(p, ln)]
    print("\nby function (instructions %, samples %)")
    for key, v in byf_i.most_common(30):
        print(f"  {key[0]:18s} {key[1]:30s} {100 * v / T:5.1f} %  {100 * byf_s[key] / S:5.1f} %")
    print(f"\ntop {top} source lines by stall samples")
    for (p, ln), v in samp.most_common(top):
        print(f"  {os.path.basename(p)}:{ln:<5d} {fn(p, ln):24s} samples {100 * v / S:5.1f} %  inst {100 * inst[(p, ln)] / T:5.1f} %   {src[(p, ln)][:90]}")


if __name__ == "__main__":
    main()
