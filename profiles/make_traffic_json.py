"""profiles/r2_ncu_traffic.json from ONE capture set (profiles/run_ncu_r2s.sh): DRAM bytes read + written per launch of
the kernels bench.py's roofline can name, taken from the `ncu --set full` raw exports in gpurun_out/.

    python profiles/make_traffic_json.py gpurun_out/r2_fused_fwd_raw.csv gpurun_out/r2_fused_bwd_raw.csv ... > profiles/r2_ncu_traffic.json

A kernel family launched several times per step (the 12 weight-gradient launches) is only reported when every launch of
one step was captured; otherwise its entry is null (bench.py then prints traffic: null for it)."""
import csv
import json
import sys

PER_STEP = {"unet_tc_fwd": 1, "unet_tc_bwd": 1, "unet_fused_fwd": 1, "unet_fused_bwd": 1, "conv3x3_wgrad": 12}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out, seen = {}, {}
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in data:
        name = r[ix["Kernel Name"]]
        key = next((k for k in PER_STEP if k in name), None)
        if key is None:
            continue
        if key == "unet_tc_fwd":
            # one kernel, two plans: its launches alternate forward, backward-data (profiles/run_ncu_r2u.sh captures one
            # of each, in that order); the backward-data plan is what bench.py calls unet_tc_bwd
            key = "unet_tc_fwd" if seen.get("unet_tc_fwd", 0) == seen.get("unet_tc_bwd", 0) else "unet_tc_bwd"
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[ix[m]].replace(",", "")) * SCALE.get(units[ix[m]], 1.0)
        out[key] = out.get(key, 0.0) + tot
        seen[key] = seen.get(key, 0) + 1
res = {k: (out[k] if seen.get(k) == n else None) for k, n in PER_STEP.items() if k in out}
res["_source"] = "profiles/run_ncu_r2u.sh (run_ncu_r2s.sh for the r2s file): " + ", ".join("%s x%d" % (k, v) for k, v in seen.items())
print(json.dumps(res, indent=1))
