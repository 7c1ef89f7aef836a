#!/usr/bin/env python
"""Print the per-kernel table of bench.py JSON lines:  python tools/showbench.py gpurun_out/b*.log"""
import json
import sys

for path in sys.argv[1:]:
    for l in open(path):
        l = l.strip()
        if not l.startswith("{"):
            continue
        d = json.loads(l)
        r = d.get("roofline") or {}
        print(f"== {path}: {d.get('ms_per_step', 0):.3f} ms/step  value {d.get('value', 0):.0f}  clocks {d.get('clocks')}  e2e {(d.get('e2e') or {}).get('value')}")
        if r:
            print(f"   top {r['kernel']} {r['achieved']:.0f} GB/s frac {r['frac']:.3f}; step frac {r['step']['frac']:.3f}; profiled {r.get('profiled_ms_per_step', 0):.3f} ms")
            for k, v in r["kernels"].items():
                print(f"   {k:24s} x{v['launches_per_step']:.0f}  {v['ms_per_step']:.3f} ms  {v['GBps']:.0f} GB/s")
