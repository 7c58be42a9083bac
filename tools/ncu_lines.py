"""Join an ncu SASS source page with nvdisasm line info: per-source-line instruction counts and stall samples.

    python tools/ncu_lines.py <report.ncu-rep> <kernel mangled name> [library.so] [cubin name substring]

Needs the library built with -lineinfo.  The SASS order of `ncu --page source` equals nvdisasm's, so rows are
joined by position.  Prints lines holding >= 0.4 % of the executed instructions or of the stall samples.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kernel = sys.argv[1], sys.argv[2]
so = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "smokephysai_b200", "libsmoke_sm100.so")
sub = sys.argv[4] if len(sys.argv) > 4 else ""
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
lines = None
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin") or sub not in f:
        continue
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if ".text." + kernel + ":" not in dis:
        continue
    body = dis.split(".text." + kernel + ":")[1]
    body = re.split(r"\n\.text\.|\n\s*\.section", body)[0]
    lines, cur = [], ("?", 0)
    for ln in body.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append((cur, ln.split("*/", 1)[1].strip().split(";")[0]))
    break
if lines is None:
    sys.exit("kernel not found in any cubin")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:.", "--kernel-name-base", "mangled"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, curb = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        curb = {"name": r[1], "rows": []}
        blocks.append(curb)
    elif curb is not None:
        curb["rows"].append(r)
blk = [b for b in blocks if kernel in b["name"]] or blocks
blk = blk[0]
hdr = blk["rows"][0]
data = blk["rows"][1:]
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [k for k, name in enumerate(hdr) if name.startswith("stall_") and "Not Issued" not in name]
if len(data) != len(lines):
    print("warning: %d ncu rows vs %d disassembled instructions" % (len(data), len(lines)))
agg = {}
for (src, _), r in zip(lines, data):
    a = agg.setdefault(src, [0.0, 0.0, {}])
    a[0] += float(r[iI] or 0)
    a[1] += float(r[iS] or 0)
    for k in stall_cols:
        v = float(r[k] or 0)
        if v:
            a[2][hdr[k]] = a[2].get(hdr[k], 0) + v
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print("total warp instructions %.0f, samples %.0f" % (ti, ts))
for src in sorted(agg, key=lambda s: (s[0], s[1])):
    a = agg[src]
    if a[0] / ti >= 0.004 or a[1] / ts >= 0.004:
        top = sorted(a[2].items(), key=lambda kv: -kv[1])[:3]
        print("%-16s:%4d  %5.1f%% inst  %5.1f%% samples   %s" % (src[0], src[1], 100 * a[0] / ti, 100 * a[1] / ts,
              " ".join("%s=%.0f" % (k.replace("stall_", ""), v) for k, v in top)))
