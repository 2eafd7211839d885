"""ncu launch list (csv of `--metrics gpu__time_duration.sum`) -> markdown table per kernel.
usage: python tools/ncu_launch_md.py launches.csv "title" "command" > profiles/x.md"""
import csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    if len(r) != len(hdr): continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
    v = float(r[vi].replace(",", "")); u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    a = agg.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[1] += ms; a[2] = max(a[2], ms)
tot = sum(a[1] for a in agg.values())
print("# %s\n\ncommand: `%s`\n(per-launch times are cold-cache and serialised -- the shares, not the absolutes, are what compares with bench.py's CUDA-event stage times)\n" % (sys.argv[2], sys.argv[3]))
print("| kernel | launches | total ms | share | largest launch ms |\n|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("| %s | %d | %.2f | %.1f%% | %.3f |" % (k, a[0], a[1], 100 * a[1] / tot, a[2]))
print("\nsum of all launches: %.1f ms" % tot)
