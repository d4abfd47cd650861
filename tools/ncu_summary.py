"""Summarise an `ncu --set full` report: python tools/ncu_summary.py report.ncu-rep "title" out.md traffic.json
(reads the report through `ncu -i ... --page raw --csv`; one section per distinct kernel instance kept)."""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio")
UNIT_TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(rep, title, out_md, out_json):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    col = {name: i for i, name in enumerate(header)}
    kname = col["Kernel Name"]
    traffic = collections.OrderedDict()
    with open(out_md, "w") as md:
        md.write(f"# {title}\n\nTimes are cold-cache and serialised (ncu replays every kernel); compare shares, not absolutes.\n")
        seen = collections.Counter()
        for r in data:
            name = re.sub(r"\(.*", "", r[kname])
            short = name.replace("relgat::", "")
            seen[short] += 1

            def val(metric):
                i = col.get(metric)
                return (r[i], units[i]) if i is not None else ("n/a", "")

            rd, ru = val("dram__bytes_read.sum")
            wr, wu = val("dram__bytes_write.sum")
            try:
                tot = float(rd.replace(",", "")) * UNIT_TO_BYTES.get(ru, 1) + float(wr.replace(",", "")) * UNIT_TO_BYTES.get(wu, 1)
                traffic.setdefault(short, {"dram_bytes_per_launch": []})["dram_bytes_per_launch"].append(int(tot))
            except ValueError:
                pass
            if seen[short] > 2:
                continue
            md.write(f"\n## `{short}` (instance {seen[short]})\n\n| metric | value | unit |\n|---|---|---|\n")
            for m in KEEP:
                v, u = val(m)
                md.write(f"| {m} | {v} | {u} |\n")
            stalls = []
            for i, h in enumerate(header):
                mt = STALL.match(h)
                if mt and "not_issued" not in h:
                    try:
                        stalls.append((float(r[i].replace(",", "")), mt.group(1) or mt.group(2)))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            md.write("\nTop stalls (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:5]) + "\n")
    with open(out_json, "w") as f:
        json.dump(traffic, f, indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:5])
