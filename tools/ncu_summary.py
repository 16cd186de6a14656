"""Summarise an `ncu --set full` raw CSV (ncu -i x.ncu-rep --page raw --csv) per kernel:
duration, tensor-pipe utilisation, DRAM bytes / throughput, L2 throughput, issue utilisation."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def f(r, name, default=0.0):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return default
    return float(r[i].replace(",", ""))


agg = collections.OrderedDict()
for r in data:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    name = re.sub(r"^void ", "", name)[:58]
    key = (name, r[col["Grid Size"]])
    a = agg.setdefault(key, collections.defaultdict(float))
    a["n"] += 1
    a["us"] += f(r, "gpu__time_duration.sum")
    a["tensor"] += f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    a["rd"] += f(r, "dram__bytes_read.sum")
    a["wr"] += f(r, "dram__bytes_write.sum")
    a["dram_pct"] += f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
    a["l2_pct"] += f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed")
    a["issue"] += f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
units = {h: u for h, u in zip(hdr, rows[1])}
print(f"# units: time {units.get('gpu__time_duration.sum')}, dram bytes {units.get('dram__bytes_read.sum')}/"
      f"{units.get('dram__bytes_write.sum')}")
print("kernel,grid,launches,avg_time,tensor_pipe_pct,dram_read_per_launch,dram_write_per_launch,"
      "dram_throughput_pct,l2_throughput_pct,issue_active_pct")
for (name, grid), a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    n = a["n"]
    print(f"{name},\"{grid}\",{int(n)},{a['us'] / n:.1f},{a['tensor'] / n:.1f},{a['rd'] / n:.2f},"
          f"{a['wr'] / n:.2f},{a['dram_pct'] / n:.1f},{a['l2_pct'] / n:.1f},{a['issue'] / n:.1f}")
