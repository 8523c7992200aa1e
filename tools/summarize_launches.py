"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the step)."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"<unnamed>::|void |gpb::|\(anonymous namespace\)::", "", row["Kernel Name"])
    name = re.sub(r"\(.*", "", name)
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print(f"# {path}: {sum(cnt.values())} launches, {T/1e3:.2f} ms total (serialised, cold cache; compare shares)")
print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"| {k} | {cnt[k]} | {v/1e3:.3f} | {v/T:.4f} | {v/cnt[k]:.2f} |")
