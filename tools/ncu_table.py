"""Turns an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... --csv` launch list
into the per-kernel table kept under profiles/ (duration, DRAM bytes, DRAM GB/s, % of the measured copy peak).

    python tools/ncu_table.py gpurun_out/all_kernels_ncu.csv [peak_GBs] > profiles/rNN_all_kernels_dram.txt"""
import csv
import sys

path = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6552.0
rows = list(csv.reader(open(path)))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, body = r, rows[i + 1:]
        break
ik, im, iv, iu, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
launches = {}
order = []
for r in body:
    if len(r) <= iv:
        continue
    key = r[iid]
    if key not in launches:
        launches[key] = {"name": r[ik]}
        order.append(key)
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    if r[im] == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    elif r[im].startswith("dram__bytes"):
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    launches[key][r[im]] = v


def short(n):
    n = n.replace("void ", "").replace("bsls::", "").replace("(int)", "").replace("(bool)", "").replace("unsigned int", "u32").replace("unsigned long", "u64")
    cut = n.find("(")
    return (n[:cut] if cut > 0 else n)[:84]


print("%-86s %9s %10s %10s %9s %9s" % ("kernel", "us", "read MB", "write MB", "DRAM GB/s", "%% of %d" % peak))
for k in order:
    L = launches[k]
    if "at::" in L["name"] or "at_cuda" in L["name"] or "cub::" in L["name"]:
        continue
    us = L.get("gpu__time_duration.sum", 0.0)
    rd, wr = L.get("dram__bytes_read.sum", 0.0), L.get("dram__bytes_write.sum", 0.0)
    gbs = (rd + wr) / us * 1e3 / 1e3 if us else 0.0  # MB / us = TB/s -> GB/s
    gbs = (rd + wr) / us * 1000.0 if us else 0.0
    print("%-86s %9.1f %10.1f %10.1f %9.0f %9.1f" % (short(L["name"]), us, rd, wr, gbs, 100.0 * gbs / peak))
