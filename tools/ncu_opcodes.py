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
iI = hdr.index("Instructions Executed"); iS = hdr.index("Source")
c = collections.Counter()
for r in data:
    src = r[iS]
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+(\.[A-Z0-9_]+)*)", src)
    op = m.group(2) if m else src[:20]
    c[op] += float(r[iI] or 0)
tot = sum(c.values())
for k, v in c.most_common(22): print("%-28s %6.2f%%" % (k, 100*v/tot))
print("total", tot)
