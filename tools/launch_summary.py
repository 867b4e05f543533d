"""Summarises an `ncu --csv` launch list (gpu__time_duration.sum [+ dram bytes]) per kernel: launches, total time, share,
DRAM bytes and GB/s.  python tools/launch_summary.py launches.csv [--each]   (runs anywhere, no GPU needed)"""
import csv
import re
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("pn::", "")
    return name[:70]


def main():
    path = sys.argv[1]
    each = "--each" in sys.argv
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    launches = OrderedDict()
    for r in rd:
        key = r["ID"]
        d = launches.setdefault(key, {"name": short(r["Kernel Name"]), "ms": 0.0, "rd": 0.0, "wr": 0.0})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        else:
            b = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            d["rd" if "read" in r["Metric Name"] else "wr"] = b
    total = sum(d["ms"] for d in launches.values())
    if each:
        for i, d in launches.items():
            gb = (d["rd"] + d["wr"]) / 1e9
            print(f"{i:>5s} {d['name']:70s} {d['ms']:9.3f} ms  {gb:8.3f} GB  {gb / max(d['ms'], 1e-9) :8.1f} TB/s".replace("TB/s", "GB/ms"))
    agg = OrderedDict()
    for d in launches.values():
        a = agg.setdefault(d["name"], {"n": 0, "ms": 0.0, "b": 0.0})
        a["n"] += 1
        a["ms"] += d["ms"]
        a["b"] += d["rd"] + d["wr"]
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        bw = a["b"] / 1e9 / (a["ms"] * 1e-3) if a["ms"] > 0 else 0.0
        print(f"{name:70s} launches {a['n']:4d}  total {a['ms']:10.3f} ms  share {100 * a['ms'] / total:5.1f}%  "
              f"dram {a['b'] / 1e9:9.2f} GB  {bw:7.0f} GB/s")
    print(f"{'total':70s} launches {len(launches):4d}  total {total:10.3f} ms")


if __name__ == "__main__":
    main()
