"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table.
usage: python tools/summarize_launches.py launches.csv "title" > profiles/rNN_launches_x.md"""
import collections
import csv
import re
import sys


def main(path, title):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
        name = re.sub(r"\(.*", "", r["Kernel Name"])[:96]
        rows.append((name, r["Grid Size"], r["Block Size"], ms))
    agg = collections.OrderedDict()
    for name, grid, block, ms in rows:
        e = agg.setdefault(name, [0, 0.0, grid, block])
        e[0] += 1
        e[1] += ms
    total = sum(ms for *_, ms in rows)
    print(f"# {title}\n")
    print("Cold-cache, serialised per-launch times (ncu replays each kernel alone): compare the SHARES with the "
          "CUDA-event table of the bench JSON, not the absolute values.\n")
    print("| launches | total ms | share | kernel | grid | block |\n|---|---|---|---|---|---|")
    for name, (n, ms, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
        print(f"| {n} | {ms:.3f} | {100 * ms / total:.1f}% | `{name}` | {grid} | {block} |")
    print(f"\ntotal {total:.2f} ms over {len(rows)} launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
