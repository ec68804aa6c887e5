"""Per-launch device times from an ncu --metrics gpu__time_duration.sum --csv log:
   python tools/launch_table.py gpurun_out/launches.csv [launches_per_frame]"""
import csv, sys, io
rows = []
txt = open(sys.argv[1]).read()
txt = txt[txt.index('"ID"'):]
for r in csv.DictReader(io.StringIO(txt)):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
        rows.append((r["Kernel Name"].split("(")[0], us))
per = int(sys.argv[2]) if len(sys.argv) > 2 else 11
nf = len(rows) // per
for i in range(per):
    ts = [rows[f * per + i][1] for f in range(nf)]
    print("%2d %-28s %s  mean %.1f" % (i, rows[i][0][:28], " ".join("%7.1f" % t for t in ts), sum(ts) / len(ts)))
print("   %-28s %s" % ("sum", " ".join("%7.1f" % sum(rows[f * per + i][1] for i in range(per)) for f in range(nf))))
