"""Turns `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` output for
scripts/conv_probe.py (one encoder pass of 8 maps + one decoder micro-batch of 10 tasks = 31 conv launches) into
profiles/conv_traffic_r1.json (read by bench.py for roofline.traffic).  usage: ncu_conv_traffic.py launches.csv out.json"""
import csv, json, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
per = {}
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    k = (r[col["ID"]], r[col["Kernel Name"]])
    v = float(r[col["Metric Value"]].replace(",", ""))
    unit = r[col["Metric Unit"]]
    name = r[col["Metric Name"]]
    if name.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    per.setdefault(k, {})[name] = v
tot = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in per.values())
n = len(per)
out = {"launches": n, "avg_dram_bytes_per_launch": tot / max(n, 1), "total_dram_bytes": tot,
       "note": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the %d conv launches of one encoder pass (8 maps) + one decoder "
               "micro-batch (10 tasks) of the 2D forward at 512x512 (scripts/conv_probe.py), per launch" % n}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(out)
