"""Per-kernel totals and the launch sequence of one steady frame from an ncu launch list
(ncu --metrics gpu__time_duration.sum --csv).  Usage: python tools/launch_summary.py launches.csv [--seq]"""
import collections
import csv
import sys


def main(path, seq=False):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
    def us(r):
        v = float(r["Metric Value"].replace(",", ""))
        return v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else (v if r["Metric Unit"] in ("us", "usecond") else v * 1e3)
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r["Kernel Name"].split("(")[0]
        tot[name][0] += 1
        tot[name][1] += us(r)
    total = sum(v[1] for v in tot.values())
    print(f"{len(rows)} launches, {total:.1f} us in total")
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.1f} us  {100 * t / total:5.1f} %  x{n:<4d} {name[:90]}")
    if seq:
        for r in rows:
            print(f"{us(r):9.1f}  {r['Grid Size']:>16s}  {r['Kernel Name'][:80]}")


if __name__ == "__main__":
    main(sys.argv[1], "--seq" in sys.argv)
