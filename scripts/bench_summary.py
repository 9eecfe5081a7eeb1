#!/usr/bin/env python
"""One-screen summary of bench.py JSON lines (gpurun_out/*.json)."""
import json
import sys

for fn in sys.argv[1:]:
    print("==", fn)
    try:
        d = json.loads(open(fn).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print("ERR", e, open(fn).read()[-600:])
        continue
    print("qps %.0f  ms/step %.2f  e2e %.0f  launches %s  clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"],
                                                                   d.get("gpu_launches"), d.get("clocks")))
    print({k: (round(v["ms"] / d["steps"], 3), v["launches"]) for k, v in d["kernel_ms"].items()})
    r = d["roofline"]
    print("roofline frac %.3f achieved %.0f" % (r["frac"], r["achieved"]), r.get("alone"))
    t = d.get("tensor_path") or {}
    print("cands/query", t.get("candidates_per_query"), "fallbacks", t.get("exact_fallback_queries"), "of", t.get("queries"))
    c = d.get("cpu_baseline") or {}
    print("cpu", c.get("value"), "id rate", c.get("gpu_vs_cpu_exact_id_rate"), "ties ok", c.get("id_mismatches_are_ties_within_1e-5"),
          "max rel", c.get("gpu_vs_cpu_max_rel_dist_err"))
    h = d.get("hbm_scan") or {}
    for cse in h.get("cases", []):
        print("  scan nq=%d call %.3f ms frac_whole %.3f" % (cse["nq"], cse["call_ms"], cse["frac_whole_call"]))
    for cse in h.get("auto_path", []):
        print("  auto nq=%d call %.3f ms  f32-bytes/peak %.3f" % (cse["nq"], cse["call_ms"], cse["f32_row_bytes_per_call_over_peak"]))
    s = d.get("e2e_single_query") or {}
    for cse in s.get("cases", []):
        print("  single-query threads=%d qps %.0f" % (cse["threads"], cse["qps"]), cse.get("frac_of_hbm_peak_whole_call"), cse.get("speedup_vs_1_thread"))
