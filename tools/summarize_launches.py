"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total us, share."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        v = float((row.get("Metric Value") or "0").replace(",", ""))
    except ValueError:
        continue
    unit = row.get("Metric Unit")
    v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    name = re.sub(r"\(.*", "", row.get("Kernel Name") or "")
    name = re.sub(r"^void ", "", name)[:80]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("launches %d  total %.1f us" % (sum(v[0] for v in agg.values()), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%-82s n=%5d %11.1f us %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
