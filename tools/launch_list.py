"""Per-kernel table of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
python tools/launch_list.py launches.csv "header comment" > profiles/xxx.txt"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict(); n = 0
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    v = float(r[mv].replace(",", "")); u = r[mu]
    ms = v / 1e6 if u.startswith("ns") else (v / 1e3 if u.startswith("us") else (v * 1e3 if u in ("s", "second") else v))
    a = agg.setdefault(r[kn].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += ms; n += 1
tot = sum(a[1] for a in agg.values())
if len(sys.argv) > 2: print("# " + sys.argv[2])
print("# per-launch times are serialised and cold-cache: only the SHARES are meaningful")
print(f"# {n} launches, {tot:.1f} ms of kernel time")
print(f"{'kernel':60s} {'launches':>8s} {'ms/launch':>10s} {'share':>7s}")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {c:8d} {t / c:10.3f} {t / tot * 100:6.1f}%")
