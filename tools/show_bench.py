"""One-line summaries of bench.py output files:  python tools/show_bench.py file.json [...]"""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(path) if l.startswith("{")][-1])
    except Exception as e:
        print(path, "unreadable:", e)
        continue
    def one(tag, d):
        if "value" not in d:
            print("  %-10s %s" % (tag, d))
            return
        ph = {k: round(v, 3) for k, v in d.get("phases_ms_per_step", {}).items()}
        clk = d.get("clocks") or {}
        print("  %-10s %8.2f G  %8.3f ms/step  e2e %7.2f G  n=%d  %s MHz %s  parity=%s" % (
            tag, d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e9, d["n_gpus"], clk.get("sm_mhz"), clk.get("reasons"),
            (d.get("parity") or {}).get("ok")))
        print("             phases", ph)
        e = d["e2e"]
        if "d2h_ceiling_gbs" in e:
            print("             d2h per gpu %.1f GB/s, ceiling %.1f per gpu / %.1f aggregate, frac %.2f" % (
                e["d2h_gbs_per_gpu"], e["d2h_ceiling_gbs_per_gpu"], e["d2h_ceiling_gbs"], e["frac_of_d2h_ceiling"]))
        if "e2e_device" in d:
            print("             e2e_device %.2f G" % (d["e2e_device"]["value"] / 1e9))
    print(path)
    one(d["config"]["workload"][:3], d)
    for k, v in (d.get("also") or {}).items():
        one(k, v)
