"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list
(units normalised: ncu picks a unit per row)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, mi, ui, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value"), h.index("ID")
T = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    k = r[ki].split("(")[0].replace("void ", "")
    a = agg.setdefault(k, {"n": set(), "us": 0.0, "rd": 0.0, "wr": 0.0})
    a["n"].add(r[ii])
    if r[mi] == "gpu__time_duration.sum":
        a["us"] += v * T.get(r[ui], 1.0)
    elif r[mi] == "dram__bytes_read.sum":
        a["rd"] += v * B.get(r[ui], 1.0)
    elif r[mi] == "dram__bytes_write.sum":
        a["wr"] += v * B.get(r[ui], 1.0)
tot = sum(a["us"] for a in agg.values())
print(f"{'kernel':58s} {'n':>5s} {'total us':>10s} {'share':>6s} {'avg us':>9s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>7s}")
for k, a in sorted(agg.items(), key=lambda x: -x[1]["us"]):
    n = len(a["n"])
    gbs = (a["rd"] + a["wr"]) / max(a["us"], 1e-9) * 1e-3
    print(f"{k[:58]:58s} {n:5d} {a['us']:10.1f} {100 * a['us'] / tot:5.1f}% {a['us'] / n:9.1f} {a['rd'] / n / 1e6:8.1f} {a['wr'] / n / 1e6:8.1f} {gbs:7.0f}")
print(f"total {tot:.1f} us over {sum(len(a['n']) for a in agg.values())} launches")
