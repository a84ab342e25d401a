import json, sys
d = json.load(open(sys.argv[1]))
r = d["roofline"]
print("fwd value %.2fM shapes/s  step %.1f us | %s %.1f us (%.3f) launches/step %.0f | %s" % (
    d["value"] / 1e6, d["ms_per_step"] * 1e3, r["kernel"][:16], r["us_per_launch"], r["frac"], d["gpu_launches"] / d["steps"],
    {k: (round(v.get("us_per_launch"), 1), round(v.get("frac", 0) or 0, 3)) for k, v in r["other_kernels"].items()}))
print("fwd_bwd %.2fM shapes/s  step %.1f us  frac %.3f | e2e %.0f shapes/s | clocks %s" % (
    d["fwd_bwd"]["value"] / 1e6, d["fwd_bwd"]["ms_per_step"] * 1e3, d["fwd_bwd"]["frac_of_peak"], d["e2e"]["value"], d.get("clocks")))
