"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel:
    python profiles/summarize_launches.py gpurun_out/<name>.csv
"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0] != "ID" and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "").replace("smk::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[14]) / 1e3
    tot = sum(a[1] for a in agg.values())
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s launches=%4d avg=%10.2f us share=%.3f" % (name[:70], n, us / n, us / tot))


if __name__ == "__main__":
    main(sys.argv[1])
