#!/usr/bin/env python
"""A bench.py JSON line cut down to what is compared between runs.  usage: python tools/show_bench.py gpurun_out/r03a/bench.json [bench_ref.json]"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
rf = d["roofline"]
print(f"value {d['value']/1e9:.1f} G/s  ms/step {d['ms_per_step']:.3f}  kernel {rf['kernel_ms']:.3f} ms  prep {rf['prep_ms']:.3f}  frac {rf['frac']:.3f}  "
      f"of level ceiling {rf['frac_of_level_ceiling']:.3f}  launches {d['gpu_launches']}  clocks {d['clocks']}")
e = d["e2e"]
print("e2e dense %.2f G/s (%.1f ms, host_bw_frac %.3f) | int32 %.2f | hits-only %.2f | bitmap %.2f | pcie %s" % (
    e["value"] / 1e9, e["ms_per_step"], e.get("host_bw_frac", 0), e.get("int32_results", {}).get("value", 0) / 1e9,
    e.get("hits_only", {}).get("value", 0) / 1e9, e.get("bitmap_only", {}).get("value", 0) / 1e9, e.get("pcie")))
cb = d.get("cpu_baseline")
if cb:
  print("cpu_baseline %.1f M/s on %s cores (%s); variants: %s" % (cb["value"] / 1e6, cb["cores"], cb.get("cpu_model"),
      {k: round(v["value"] / 1e6, 1) for k, v in cb.get("variants", {}).items() if isinstance(v, dict) and "value" in v}))
c = d.get("cli_e2e") or {}
print("cli:", {k: round(v["lookups_per_s"] / 1e6, 1) for k, v in c.items() if isinstance(v, dict) and "lookups_per_s" in v}, "M lookups/s")
for k, w in d.get("workloads", {}).items():
    r = w["roofline"]
    print(f"  {k}: {w['value']/1e9:.1f} G/s kernel {r.get('kernel_ms', 0):.2f} ms frac {r.get('frac', 0):.3f} level {r.get('frac_of_level_ceiling', 0):.3f} "
          f"traffic/alg {r.get('traffic_over_algorithmic')} parity {w['parity'].get('weighted_checksum_match')}")
print("parity:", {k: v for k, v in d["parity"].items() if k.endswith("match") or k in ("reads", "checker")})
if len(sys.argv) > 2:
    r = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][0])
    print("reference arm: %.1f M lookups/s (%s cores) -> e2e ratio %.1f x, kernel ratio %.0f x" % (r["value"] / 1e6, r["cpu_baseline"]["cores"], e["value"] / r["value"], d["value"] / r["value"]))
