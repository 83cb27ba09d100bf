"""Aggregate an `ncu --csv --metrics ...` launch list: per kernel name -> launches, total / average of each metric.
Usage: python tools/ncu_launches.py gpurun_out/launches.csv [--json out.json]"""
import csv
import json
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(.*$", "", name)                       # drop the argument list
    name = name.replace("void ", "").replace("desta::", "")
    name = re.sub(r"at::native::.*?(\w+_kernel\w*|elementwise_kernel|\w+Kernel\w*).*", r"torch:\1", name)
    return name.strip()


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    per = defaultdict(lambda: defaultdict(float))
    count = defaultdict(set)
    for r in rd:
        if len(r) < len(hdr):
            continue
        k = short(r[ix["Kernel Name"]])
        m, u, v = r[ix["Metric Name"]], r[ix["Metric Unit"]], r[ix["Metric Value"]].replace(",", "")
        try:
            v = float(v)
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        per[k][m] += v * scale
        count[k].add(r[ix["ID"]])
    tot_ms = sum(d.get("gpu__time_duration.sum", 0.0) for d in per.values())
    out = []
    for k, d in sorted(per.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum", 0.0)):
        n = len(count[k])
        ms = d.get("gpu__time_duration.sum", 0.0)
        e = {"kernel": k, "launches": n, "total_ms": round(ms, 3), "avg_us": round(1e3 * ms / max(n, 1), 1), "share": round(ms / tot_ms, 4) if tot_ms else None}
        if "dram__bytes_read.sum" in d:
            e["dram_read_MB_per_launch"] = round(d["dram__bytes_read.sum"] / n / 1e6, 2)
            e["dram_write_MB_per_launch"] = round(d["dram__bytes_write.sum"] / n / 1e6, 2)
        out.append(e)
    print("| kernel | launches | total ms | avg us | share | DRAM rd+wr MB / launch |\n|---|---|---|---|---|---|")
    for e in out[:25]:
        dr = f"{e.get('dram_read_MB_per_launch', '')} + {e.get('dram_write_MB_per_launch', '')}" if "dram_read_MB_per_launch" in e else ""
        print(f"| `{e['kernel'][:70]}` | {e['launches']} | {e['total_ms']} | {e['avg_us']} | {e['share']} | {dr} |")
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
