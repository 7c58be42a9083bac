import csv, subprocess, sys, re, collections
rep, kernel = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:.", "--kernel-name-base", "mangled"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, curb = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        curb = {"name": r[1], "rows": []}; blocks.append(curb)
    elif curb is not None: curb["rows"].append(r)
blk = [b for b in blocks if kernel in b["name"]][0]
hdr = blk["rows"][0]; data = blk["rows"][1:]
cols = [k for k, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
c = collections.Counter()
for r in data:
    for k in cols:
        v = float(r[k] or 0)
        if v: c[hdr[k]] += v
tot = sum(c.values())
for k, v in c.most_common(14): print("%-28s %6.2f%%" % (k, 100*v/tot))
