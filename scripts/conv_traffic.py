"""Summarise an ncu CSV (`--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`) of
scripts/profile_one.py into profiles/conv_tc_traffic_r1.json: DRAM bytes per conv_tc launch (mean over the
launches of one C3 UNet forward + one KITTI decode), the `roofline.traffic` figure bench.py reports.
   python scripts/conv_traffic.py gpurun_out/conv_traffic.csv profiles/conv_tc_traffic_r1.json"""
import csv, json, sys, collections
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")))
per = collections.defaultdict(dict)
for r in rows:
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3,
             "usecond": 1, "msecond": 1e3}.get(u, 1)
    per[r["ID"]][r["Metric Name"]] = v * scale
    per[r["ID"]]["name"] = r["Kernel Name"]
conv = [d for d in per.values() if "conv_tc" in d["name"]]
rd = sum(d.get("dram__bytes_read.sum", 0) for d in conv)
wr = sum(d.get("dram__bytes_write.sum", 0) for d in conv)
us = sum(d.get("gpu__time_duration.sum", 0) for d in conv)
out = {"kernel": "conv_tc_kernel / conv_tc_persistent_kernel", "launches": len(conv),
       "dram_read_bytes_per_launch": rd / len(conv), "dram_write_bytes_per_launch": wr / len(conv),
       "traffic_bytes_per_launch": (rd + wr) / len(conv), "ncu_time_us_per_launch": us / len(conv),
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, "
                 "scripts/profile_one.py both 8 (one C3 UNet forward + one KITTI decode, batch 8; cold caches)"}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(out)
