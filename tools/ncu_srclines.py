"""Per-source-line instruction counts and stall samples of one kernel, straight from ncu's own source correlation.

    python tools/ncu_srclines.py <report.ncu-rep> <kernel name regex> [launch index among the matches, default 0] [min percent, default 0.4]

Needs a capture made with `--import-source on` of a library built with -lineinfo.  Unlike tools/ncu_lines.py (which joins the SASS
page with nvdisasm by position and so needs the very same build of the library) this reads the `cuda,sass` source page, whose rows
that carry a line number are already aggregated per source line.
"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.4
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the page is a sequence of (File Path, Function Name, header, rows...) blocks; a launch is a run of blocks until a file repeats
launches, cur, seen, fname, hdr = [], None, set(), None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1]
        if cur is None or fname in seen:
            cur, seen = [], set()
            launches.append(cur)
        seen.add(fname)
    elif r[0] == "Line No":
        hdr = r
    elif r[0] == "Function Name":
        pass
    elif r[0].isdigit() and hdr is not None:
        # ncu does not escape quotes inside the source text: take the metric columns from the right
        nm = len(hdr) - 4
        if len(r) >= nm + 2:
            cur.append((fname.split("/")[-1], int(r[0]), ",".join(r[1:len(r) - nm - 2]), dict(zip(hdr[4:], r[-nm:]))))
if not launches:
    sys.exit("no source page for " + pat)
L = launches[min(which, len(launches) - 1)]
def f(v):
    try:
        return float(v)
    except ValueError:
        return 0.0
ti = sum(f(d["Instructions Executed"]) for _, _, _, d in L) or 1
ts = sum(f(d["# Samples"]) for _, _, _, d in L) or 1
print("launches on the page: %d; this one: %.0f warp instructions, %.0f samples" % (len(launches), ti, ts))
for fn, ln, src, d in sorted(L, key=lambda x: (x[0], x[1])):
    pi, ps = 100 * f(d["Instructions Executed"]) / ti, 100 * f(d["# Samples"]) / ts
    if pi < minpct and ps < minpct:
        continue
    st = sorted(((k[6:], f(v)) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and f(v) > 0), key=lambda kv: -kv[1])[:3]
    print("%-18s:%4d %5.1f%% inst %5.1f%% smp  %-44s %s" % (fn, ln, pi, ps, " ".join("%s=%.0f" % kv for kv in st), src.strip()[:70]))
