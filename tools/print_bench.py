"""One-line summary of a bench.py JSON line: python tools/print_bench.py file.json"""
import json, sys
d = json.load(open(sys.argv[1]))
r = d.get("roofline", {})
print(f"value {d['value']:.4g} {d['unit']}  step {d['ms_per_step']:.4f} ms  kernel {r.get('kernel_ms', 0):.4f} ms  "
      f"frac {r.get('frac', 0):.3f}  e2e {d['e2e']['ms_per_step']:.4f} ms ({d['e2e']['value']:.4g})  launches {d.get('gpu_launches')}  clocks {d.get('clocks')}")
