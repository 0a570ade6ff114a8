"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for the UNet
forward and the VAE decode of scripts/profile_one.py.   python scripts/launch_summary.py <csv> [--convs]"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    if row.get("Metric Name", "gpu__time_duration.sum") != "gpu__time_duration.sum":
        continue                      # the same list may carry DRAM-byte metrics (scripts/conv_traffic.py)
    v = float(row["Metric Value"].replace(",", ""))
    us = v / 1000 if row["Metric Unit"] in ("ns", "nsecond") else v
    seq.append((row["Kernel Name"].split("(")[0][:44], us, row["Grid Size"]))
ins = [i for i, (n, _, _) in enumerate(seq) if "conv_in" in n]
split = ins[1] if len(ins) > 1 else len(seq)
for part, (a, b) in {"unet": (0, split), "decoder": (split, len(seq))}.items():
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for n, us, g in seq[a:b]:
        tot[n] += us
        cnt[n] += 1
    print(f"{part}: {b - a} launches, {sum(tot.values()):.0f} us (serialised, cold-cache ncu times)")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"   {k:46s} n={cnt[k]:4d} total {v:8.1f} us  avg {v / cnt[k]:7.1f}  share {100 * v / sum(tot.values()):5.1f}%")
if "--convs" in sys.argv:
    print([(g, round(us, 1)) for n, us, g in seq if "conv_tc" in n])
