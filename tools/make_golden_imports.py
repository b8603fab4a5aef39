"""tests/golden/reference_imports.json: every name the reference's hot-path scripts import from the packages this repo
shims (tritonoa, oao, ax, botorch, gpytorch), read from the reference's own files with `ast` -- the CPU tier then checks
that each of them resolves under bogp_b200.shim.install().  Run in the build container (needs /root/reference):
    python tools/make_golden_imports.py"""
import ast
import json
from pathlib import Path

REF = Path("/root/reference")
FILES = [
    "optimization/strategy.py", "optimization/utils.py", "optimization/gpmodels.py",
    "projects/swellex96_localization/conf/optimization/run.py", "projects/swellex96_localization/data/run.py",
    "projects/swellex96_localization/data/mfp.py", "projects/swellex96_localization/data/demo_1d.py",
    "projects/swellex96_loc_tilt/conf/run.py", "projects/swellex96_loc_tilt/data/mfp.py",
    "projects/swellex96_inv/data/bo/ei.py", "projects/swellex96_inv/data/bo/ucb.py", "projects/swellex96_inv/data/bo/logei.py",
    "projects/swellex96_inv/data/bo/pi.py", "projects/swellex96_inv/data/bo/baxus.py", "projects/swellex96_inv/data/bo/obj.py",
    "projects/swellex96_inv/data/bo/sobol.py", "projects/swellex96_inv/data/bo/helpers.py",
]
TOPS = {"tritonoa", "oao", "ax", "botorch", "gpytorch"}

out = {}
for rel in FILES:
    p = REF / rel
    if not p.exists():
        continue
    names = []
    for node in ast.walk(ast.parse(p.read_text())):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.split(".")[0] in TOPS:
            names += [[node.module, a.name, node.lineno] for a in node.names]
        elif isinstance(node, ast.Import):
            names += [[a.name, None, node.lineno] for a in node.names if a.name.split(".")[0] in TOPS]
    if names:
        out[rel] = sorted(names, key=lambda t: t[2])
dst = Path(__file__).resolve().parent.parent / "tests" / "golden" / "reference_imports.json"
dst.write_text(json.dumps(out, indent=1) + "\n")
print(sum(len(v) for v in out.values()), "imports from", len(out), "files ->", dst)
