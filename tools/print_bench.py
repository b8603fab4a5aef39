"""One-line summary of a bench.py JSON line: python tools/print_bench.py file.json   (or the line on stdin)"""
import json, sys
txt = open(sys.argv[1]).read() if len(sys.argv) > 1 and sys.argv[1] != "-" else sys.stdin.read()
d = json.loads(txt.strip().splitlines()[-1])
r = d.get("roofline", {})
tag = sys.argv[2] if len(sys.argv) > 2 else ""
print(f"{tag} value {d['value']:.4g} {d['unit']}  step {d['ms_per_step']:.4f} ms  kernel {r.get('kernel_ms', 0):.4f} ms  "
      f"frac {r.get('frac', 0):.3f}  e2e {d['e2e']['ms_per_step']:.4f} ms ({d['e2e']['value']:.4g})  launches {d.get('gpu_launches')}")
