"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share."""
import csv, re, sys, collections

def short(n):
    n = re.sub(r"\(.*", "", n)
    n = re.sub(r"<.*", lambda m: m.group(0)[:48], n)
    return n.replace("void ", "")[:100]

def main(path, skip=0, take=None):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]))
    rows = rows[skip: None if take is None else skip + take]
    agg = collections.OrderedDict()
    for n, t, *_ in rows:
        a = agg.setdefault(short(n), [0, 0.0])
        a[0] += 1; a[1] += t
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot/1e6:.2f} ms summed kernel time (cold-cache, serialised)")
    print(f"{'kernel':100s} {'n':>6s} {'ms':>10s} {'share':>7s} {'us/launch':>10s}")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:100s} {c:6d} {t/1e6:10.3f} {100*t/tot:6.1f}% {t/c/1e3:10.1f}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else None)
