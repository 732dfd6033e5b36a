import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
r=d["roofline"]; s=d.get("sustained") or {}
print("value %.1f ms %.3f gemm_ms %.3f achieved %.1f frac_mix %.3f | sustained %.1f gemm_ms %.3f clk %s pw %s | e2e %.1f | parity %s" % (
 d["value"], d["ms_per_step"], r["gemm_ms_per_step"], r["achieved"], r.get("frac_of_mix_ceiling",0), s.get("value",0), s.get("gemm_ms_per_step",0), (s.get("clocks") or {}).get("sm_mhz"), (s.get("clocks") or {}).get("power_w"), d["e2e"]["value"], d["parity"]))
