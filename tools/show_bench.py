"""One-line summary of a bench.py JSON line.  python tools/show_bench.py file.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline", {})
if d.get("impl") == "reference":
    c = d["cpu_baseline"]; print("reference arm: %.0f objects/s on %d cores (%s), svd default %.0f" % (d["value"], c["cores"], c["kind"], c["svd_method_true"]["value"]))
elif "factor_kernel" in r:
    f = d.get("facade_e2e") or {}
    print("N=%d %s value %.3e obj/s step %.2f ms (factor %.2f ms frac %.3f | grid %.2f ms frac %.3f | whole-step frac %.3f) e2e %.3e obj/s (%.2f ms)  LL kernel %.2f ms wall %.2f ms  cpu %s  facade %s (ref %s)  parity %.1e  clocks %s" % (
        d["n_gpus"], d["scaling"], d["value"], d["ms_per_step"], r["factor_kernel"]["ms_per_launch"], r["factor_kernel"]["frac"], r["ms_per_launch"], r["frac"],
        r["whole_step"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["ll_evaluation"]["kernel_ms"], d["ll_evaluation"]["wall_ms"],
        ("%.0f" % d["cpu_baseline"]["value"]) if "cpu_baseline" in d else "-", ("%.3e" % f["value"]) if f else "-",
        ("%.0f" % f["reference"]["value"]) if f and f.get("reference") else "-", d["parity_max_rel_err"], d.get("clocks")))
else:
    print("N=%d %s value %.3e %s  %.2f ms/step  frac %.3f  e2e %.3e (%.1f ms)  parity %.1e" % (d["n_gpus"], d["config"]["workload"][:40], d["value"], d["unit"], d["ms_per_step"],
          r.get("frac", 0), d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("parity_max_rel_err", -1)))
