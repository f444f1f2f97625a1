"""DRAM traffic per launch of the kernels of the BB solve on C5, from an `ncu --set full` report of
`python bench.py --steps 1 --warmup 3 --skip-extras` (profile-from-start off: only the timed solve).

    ncu -i <report>.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_traffic.py raw.csv <git hash> > profiles/traffic.json
"""
import csv
import json
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ci = {h: i for i, h in enumerate(hdr)}


def num(r, name):
    v = float(r[ci[name]].replace(",", ""))
    u = units[ci[name]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3,
             "msecond": 1.0, "nsecond": 1e-6}.get(u, 1)
    return v * scale


agg = OrderedDict()
for r in rows[2:]:
    name = r[ci["Kernel Name"]].split("(")[0].replace("void ", "")
    a = agg.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "ms": 0.0, "l1tex_pct": 0.0, "lts_pct": 0.0})
    a["launches"] += 1
    a["dram_bytes"] += num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")
    a["ms"] += num(r, "gpu__time_duration.sum")
    a["l1tex_pct"] += float(r[ci["l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]])
    a["lts_pct"] += float(r[ci["lts__throughput.avg.pct_of_peak_sustained_elapsed"]])
out = OrderedDict()
out["source"] = ("ncu --profile-from-start off --set full --clock-control none on `python bench.py --steps 1 --warmup 3 --skip-extras` "
                 "(C5, one B200), commit %s: dram__bytes_read.sum + dram__bytes_write.sum per launch" % (sys.argv[2] if len(sys.argv) > 2 else "?"))
kern = OrderedDict()
for name, a in agg.items():
    n = a["launches"]
    kern[name] = {"launches_captured": n, "dram_bytes_per_launch": a["dram_bytes"] / n, "ms_per_launch_under_ncu": a["ms"] / n,
                  "l1tex_throughput_pct": a["l1tex_pct"] / n, "lts_throughput_pct": a["lts_pct"] / n}
out["kernels"] = kern
# the two products of one evaluation
ell = [k for k in kern if "spmv_ell" in k and "GradBB" in k]
vec = [k for k in kern if "spmv_vector8" in k]
spmv = {}
if ell:
    spmv["atr"] = kern[ell[0]]["dram_bytes_per_launch"]
if vec:
    spmv["ax_per_panel_launch"] = kern[vec[0]]["dram_bytes_per_launch"]
out["spmv_c5_bytes_per_launch"] = spmv
print(json.dumps(out, indent=1))
