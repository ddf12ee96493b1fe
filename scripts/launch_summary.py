"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) of scripts/ncu_probe.py by kernel.
    python scripts/launch_summary.py gpurun_out/xyz.csv [top]"""
import collections, csv, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for x in csv.DictReader(lines):
    rows.append((x["Kernel Name"], float(x["Metric Value"].replace(",", ""))))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
names = [re.sub(r"\(.*", "", n) for n, _ in rows]
idx = [i for i, n in enumerate(names) if n.startswith("k_rss_cell")]
def agg(seg, title):
    a = collections.OrderedDict()
    for n, v in seg:
        n = re.sub(r"\(.*", "", n); n = re.sub(r"void |cub::.*?::", "", n)[:64]
        e = a.setdefault(n, [0, 0.0]); e[0] += 1; e[1] += v
    tot = sum(v for _, v in a.values())
    print(f"== {title}: total {tot/1e6:.1f} ms, {len(seg)} launches")
    for n, (c, v) in sorted(a.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{n:64s} {c:5d} {v/1e6:9.3f} ms {100*v/tot:5.1f}%  avg {v/c/1e3:8.1f} us")
if len(idx) >= 2:
    agg(rows[:idx[0] + 1], "setup + initialize")
    agg(rows[idx[0] + 1:idx[1] + 1], "outer iteration 1 (update_A, update_B, rss)")
else:
    agg(rows, "all")
