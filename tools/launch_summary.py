"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
seq = []
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += v
    seq.append((name, v, row.get("Grid Size", "")))
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':42s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:42s} {v[0]:8d} {v[1]/1e3:10.3f} {v[1]/tot*100:6.1f}% {v[1]/v[0]:9.1f}")
print(f"{'TOTAL':42s} {sum(v[0] for v in agg.values()):8d} {tot/1e3:10.3f}")
if len(sys.argv) > 2:
    a = [i for i, s in enumerate(seq) if s[0].startswith("qr_assemble")]
    for s in seq[a[-1]:a[-1] + int(sys.argv[2])]:
        print(s)
