#!/usr/bin/env python
"""One-line digest of a bench.py JSON line."""
import json
import sys
for p in sys.argv[1:]:
    try:
        d = json.load(open(p))
    except Exception as e:
        print(p, "unreadable:", e)
        continue
    k = d.get("kernels_ms_per_step", {})
    print(f"{d['config']['workload']:5s} n={d['n_gpus']} enc {d['value']:7.1f} GB/s  dec {d['decode']['value']:7.1f}  "
          f"e2e {d['e2e']['value']:6.1f}/{d.get('e2e_decode', {}).get('value', 0):6.1f}  "
          f"kernels ms: " + " ".join(f"{a}={b:.3f}" for a, b in k.items()) +
          f"  roofline {d.get('roofline', {}).get('frac', 0):.4f}")
