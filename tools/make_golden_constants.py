"""Golden copy of the reference's workload constants (tonal sets, array geometry, truth point, inversion search space),
read from the reference's own files: literal assignments are evaluated from the parsed source (``ast``) of
projects/swellex96_inv/data/bo/common.py and projects/swellex96_inv/data/env.py, the tonal list from
projects/swellex96_localization/conf/data/acoustics.yaml.  tests/test_oracle.py checks bogp_b200.modes and the bench
configuration against it.  Run in the build container (needs /root/reference):  python tools/make_golden_constants.py
"""
import ast
import json
import re
from pathlib import Path

import numpy as np

REF = Path("/root/reference/projects")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "reference_constants.json"


def literals(path: Path, names):
    """Module-level ``NAME = <expression of literals and earlier names>`` assignments, evaluated in order."""
    ns = {}
    for node in ast.parse(path.read_text()).body:
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            try:
                ns[node.targets[0].id] = eval(compile(ast.Expression(node.value), str(path), "eval"), {"__builtins__": {}}, dict(ns))
            except Exception:
                pass
    return {k: ns[k] for k in names}


def main():
    common = literals(REF / "swellex96_inv/data/bo/common.py", ["FREQ", "TRUE_R", "TRUE_SRC_Z", "TILT", "H_W", "SEARCH_SPACE"])
    env_src = (REF / "swellex96_inv/data/env.py").read_text()
    m = re.search(r'"rec_z":\s*\[(.*?)\]', env_src, re.S)
    rec_z_21 = [float(x) for x in re.findall(r"[-+]?\d+\.?\d*", m.group(1))]
    loc_env = (REF / "swellex96_localization/data/env.py").read_text()
    m64 = re.search(r'"rec_z":\s*np\.linspace\(([^)]*)\)', loc_env)
    lo, hi, n = [float(x) for x in m64.group(1).split(",")]
    yaml_src = (REF / "swellex96_localization/conf/data/acoustics.yaml").read_text()
    tonals = [float(x) for x in re.search(r"frequencies:\s*\[([^\]]*)\]", yaml_src).group(1).split(",")]
    out = {"inv_tonals": common["FREQ"], "truth_r_km": common["TRUE_R"], "truth_src_z_m": common["TRUE_SRC_Z"],
           "inv_tilt_deg": common["TILT"], "water_depth_m": common["H_W"], "inv_search_space": common["SEARCH_SPACE"],
           "rec_z_21": rec_z_21, "rec_z_64_linspace": [lo, hi, int(n)], "rec_z_64": np.linspace(lo, hi, int(n)).tolist(),
           "swellex_tonals_13": tonals,
           "sources": ["projects/swellex96_inv/data/bo/common.py:7-72", "projects/swellex96_inv/data/env.py:51-73",
                       "projects/swellex96_localization/data/env.py:49", "projects/swellex96_localization/conf/data/acoustics.yaml:44"]}
    OUT.write_text(json.dumps(out, indent=1))
    print("wrote", OUT, "| 21-element array:", len(rec_z_21), "| tonals:", len(tonals))


if __name__ == "__main__":
    main()
