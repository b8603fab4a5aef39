"""Golden values for the optimiser front-end's search-space formatting, generated from the reference's own code.

``optimization/utils.py`` cannot be imported here (hydra, oao, tritonoa are absent) but ``format_search_space``
(:72-98) only touches attributes of its arguments: this script cuts the function out of the reference file with ``ast``
and executes it unchanged, with plain stand-ins for ``SearchParameter`` (keeps the keyword dict) and ``SearchSpace``
(keeps the list), on the search spaces of the localisation configs (ref projects/swellex96_localization/conf/
optimization/run.py:160-185: relative range/depth bounds clipped to absolute limits, absolute tilt bounds).
Run in the build container (needs /root/reference):   python tools/make_golden_frontend.py
"""
import ast
import json
from pathlib import Path
from types import SimpleNamespace

REF = Path("/root/reference/optimization/utils.py")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "format_search_space.json"


def main():
    tree = ast.parse(REF.read_text())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "format_search_space"]
    assert len(fn) == 1
    ns = {"SearchParameter": lambda **kw: kw, "SearchSpace": list, "SearchSpaceBounds": object}
    exec(compile(ast.Module(body=fn, type_ignores=[]), str(REF), "exec"), ns)
    spaces = {
        "localization": [dict(name="rec_r", lower_bound=-1.0, upper_bound=1.0, relative=True, min_lower_bound=0.01, max_upper_bound=10.0),
                         dict(name="src_z", lower_bound=-40.0, upper_bound=40.0, relative=True, min_lower_bound=1.0, max_upper_bound=200.0)],
        "loc_tilt": [dict(name="rec_r", lower_bound=-1.0, upper_bound=1.0, relative=True, min_lower_bound=0.01, max_upper_bound=10.0),
                     dict(name="src_z", lower_bound=-40.0, upper_bound=40.0, relative=True, min_lower_bound=1.0, max_upper_bound=200.0),
                     dict(name="tilt", lower_bound=-4.0, upper_bound=4.0, relative=False, min_lower_bound=-4.0, max_upper_bound=4.0)],
    }
    scenarios = [dict(rec_r=5.0, src_z=60.0), dict(rec_r=0.5, src_z=20.0), dict(rec_r=9.6, src_z=185.0),
                 dict(rec_r=0.01, src_z=1.0), dict(rec_r=10.0, src_z=200.0), dict(rec_r=1.0, src_z=41.0, tilt=1.5)]
    cases = []
    for name, bounds in spaces.items():
        for sc in scenarios:
            space = SimpleNamespace(bounds=[SimpleNamespace(**b) for b in bounds])
            out = ns["format_search_space"](space, sc)
            cases.append({"space": name, "bounds": bounds, "scenario": sc, "parameters": out})
    OUT.write_text(json.dumps({"source": str(REF) + ":72-98", "cases": cases}, indent=1))
    print("wrote", OUT, len(cases), "cases")


if __name__ == "__main__":
    main()
