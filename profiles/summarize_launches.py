"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r["Kernel Name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        name = re.sub(r"\(.*", "", name).replace("e2e::", "").replace("void ", "")
        name = re.sub(r"<.*", "", name)
        agg[name][0] += 1
        agg[name][1] += float(r["Metric Value"]) / 1e3
    tot = sum(v[1] for v in agg.values())
    print("launches %d, total %.1f us (cold-cache, serialised: compare shares, not absolutes)" % (len(rows), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-34s n=%5d  %10.1f us  %5.1f%%  avg %8.1f us" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))


if __name__ == "__main__":
    main(sys.argv[1])
