"""Summarise an ncu --csv launch list (gpu__time_duration.sum) per kernel: count, mean us, share."""
import collections, csv, re, sys

def main(path, per_step=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void |mvb::", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
        key = (name, r.get("Grid Size", ""))
        agg[key][0] += 1; agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, {tot:.1f} us total")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[0][:52]:52s} grid={k[1]:>14s} n={v[0]:4d} mean={v[1]/v[0]:8.2f} us  share={v[1]/tot*100:5.1f}%")

if __name__ == "__main__":
    main(sys.argv[1])
