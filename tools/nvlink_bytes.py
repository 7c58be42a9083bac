"""NVLink payload counters around a command:  python tools/nvlink_bytes.py <cmd ...>   (sums `nvidia-smi nvlink -gt d` over the links of
every GPU before and after; prints the difference per GPU in MB: evidence that the slab halo rows travel over NVLink)."""
import re
import subprocess
import sys


def read():
    out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d"], capture_output=True, text=True).stdout
    tot, gpu = {}, None
    for ln in out.splitlines():
        m = re.match(r"GPU (\d+):", ln)
        if m:
            gpu = int(m.group(1))
            tot[gpu] = [0, 0]
        m = re.search(r"Data (Tx|Rx): (\d+) KiB", ln)
        if m and gpu is not None:
            tot[gpu][0 if m.group(1) == "Tx" else 1] += int(m.group(2))
        if re.search(r"Data (Tx|Rx): N/A", ln):
            tot["na"] = True
    return tot


a = read()
rc = subprocess.run(sys.argv[1:]).returncode
b = read()
if b.pop("na", False) | a.pop("na", False):
    print("nvidia-smi nvlink -gt d reports N/A on this box (the pool's VMs do not expose the NVLink throughput counters)")
for g in sorted(b):
    if g in a:
        print("GPU %d: NVLink data tx %.1f MB, rx %.1f MB during the command" % (g, (b[g][0] - a[g][0]) / 1024.0, (b[g][1] - a[g][1]) / 1024.0))
sys.exit(rc)
