#!/usr/bin/env python
"""one screen of a bench.py line: python scripts/show_bench.py gpurun_out/bench.json"""
import json
import sys

d = json.load(open(sys.argv[1]))


def show(k, v):
    if isinstance(v, dict) and "error" in v:
        print(k, "ERROR", v["error"])
        return
    r = v.get("roofline") or {}
    e = v.get("e2e") or {}
    print(f"{k:15s} {v.get('metric')} {v.get('value', 0):.4g} {v.get('unit')}  ms/step {v.get('ms_per_step')}  frac {r.get('frac')}  kernel_ms {r.get('kernel_ms_per_launch')}"
          f"  digest {v.get('result_digest')}  parity {(v.get('parity_sample') or {}).get('ok')}  e2e ms/call {e.get('ms_per_call')} {e.get('result_digest')}")


show("main", d)
for k in d:
    if k.startswith("secondary"):
        show(k, d[k])
print("e2e", {k: v for k, v in (d.get("e2e") or {}).items() if k not in ("api", "input_order")})
print("cplane", d.get("e2e_cplane"))
print("tol_sweep", [(t["tolerance"], round(t["ms_per_step"], 2), round(t["kernel_ms"], 2), t["edges"]) for t in d.get("tol_sweep", [])])
s = d.get("secondary") or {}
print("hash step_share", (s.get("roofline") or {}).get("step_share"), "frac_bytes_moved", (s.get("roofline") or {}).get("frac_bytes_moved"), s.get("bit_mismatch"))
print("hash burst", s.get("burst"), "sustained clocks", s.get("clocks"))
x = d.get("secondary_e2e") or {}
print("e2e10m", x.get("seconds"), x.get("phases_s_rank0"))
print("clocks", d.get("clocks"))
