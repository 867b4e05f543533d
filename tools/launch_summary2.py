"""Summarises an ncu --csv launch list (tools/launch_list.sh): per kernel device time, share, tensor pipe, DRAM GB/s."""
import csv
import sys

for path in sys.argv[1:]:
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    d = {}
    for r in rows[1:]:
        d.setdefault((int(r[iid]), r[ik]), {})[r[im]] = float(r[iv].replace(",", ""))
    tot = sum(v["gpu__time_duration.sum"] for v in d.values())
    print(f"== {path}: {len(d)} launches, {tot / 1e6:.3f} ms in kernels")
    for (i, name), v in sorted(d.items()):
        us = v["gpu__time_duration.sum"] / 1e3
        rd, wr = v.get("dram__bytes_read.sum", 0) / 1e6, v.get("dram__bytes_write.sum", 0) / 1e6
        pipe = v.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0)
        ghz = v.get("sm__cycles_elapsed.max", 0) / (us * 1e3) if us else 0
        print(f"{i:3d} {name[:44]:44s} {us:9.1f} us {100 * us * 1e3 / tot:5.1f}%  tensor pipe {pipe:5.1f}%  DRAM rd {rd:8.1f} MB wr {wr:8.1f} MB "
              f"= {(rd + wr) / us if us else 0:5.2f} TB/s  SM clock {ghz:.2f} GHz")
