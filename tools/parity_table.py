"""Per-tensor parity table of one whole-step case (tests/stage_checks.check_step) under the current PAIG_* environment:
    python tools/parity_table.py mnist_spring_color 100
prints every gradient tensor's error against the fp32 oracle, against its float64 twin and the oracle's own fp32 noise."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import backends
import stage_checks as sc

task, B = sys.argv[1], int(sys.argv[2])
kw = json.loads(sys.argv[3]) if len(sys.argv) > 3 else {}
if "horizon" in kw:
    kw["horizon"] = tuple(kw["horizon"])
be = backends.get("cuda")
report = {}
status = "pass"
try:
    sc.check_step(be, task, B, report=report, **kw)
except AssertionError as e:
    status = "FAIL"
env = {k: v for k, v in os.environ.items() if k.startswith("PAIG_")}
print("== %s B=%d %s env=%s -> %s, relu flips %s" % (task, B, kw, env, status, report.get("relu_flips")))
rows = []
for k, v in report.items():
    if not isinstance(v, dict) or k.startswith("fwd/"):
        continue
    bound = 1e-4 + 4 * (v["ref_noise"] or 0.0)
    used = (v["err_vs_f64"] if v["err_vs_f64"] is not None else v["err"]) / bound
    rows.append((used, k, v))
for used, k, v in sorted(rows, reverse=True)[:int(os.environ.get("TABLE_ROWS", "14"))]:
    print("  %-52s err %.2e  vs_f64 %s  ref_noise %s  used %.2f" % (
        k, v["err"], "%.2e" % v["err_vs_f64"] if v["err_vs_f64"] is not None else "   -    ",
        "%.2e" % v["ref_noise"] if v["ref_noise"] is not None else "   -    ", used))
