"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name."""
import csv, sys, collections
# --skip-pack: leave the once-per-weight-update packing kernels out (they run in the first step only)
skip_pack = "--skip-pack" in sys.argv
args = [a for a in sys.argv[1:] if not a.startswith("--")]
rows = []
with open(args[0]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
tot = collections.OrderedDict()
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"].split("(")[0]
    if skip_pack and (name.startswith("k_pack") or name.startswith("k_combine")):
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    d = tot.setdefault(name, [0, 0.0])
    d[0] += 1; d[1] += v
total = sum(d[1] for d in tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'share':>7s} {'avg us':>10s}")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
  try:
    print(f"{k[:60]:60s} {n:8d} {t:12.1f} {100*t/total:6.1f}% {t/n:10.1f}")
  except BrokenPipeError:
    sys.exit(0)
print(f"{'TOTAL':60s} {sum(d[0] for d in tot.values()):8d} {total:12.1f}")
