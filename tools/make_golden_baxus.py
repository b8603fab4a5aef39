"""Golden values for the BAxUS state machine, generated from the reference's own code.

The reference module (projects/swellex96_inv/data/bo/baxus.py) cannot be imported here (botorch / gpytorch are not
installed), but ``BaxusState`` and ``update_state`` only need math / numpy / dataclasses: this script cuts those two
definitions out of the reference file with ``ast`` and executes them as they are, then records their outputs for a
range of inputs in tests/golden/baxus_state.json.  Run in the build container (needs /root/reference):

    python tools/make_golden_baxus.py
"""
import ast
import json
import math
from dataclasses import dataclass, field  # noqa: F401  (used by the extracted source)
from pathlib import Path

import numpy as np

REF = Path("/root/reference/projects/swellex96_inv/data/bo/baxus.py")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "baxus_state.json"


class _Scalar(float):
    def item(self):
        return float(self)


def main():
    src = REF.read_text()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if (isinstance(n, ast.ClassDef) and n.name == "BaxusState")
              or (isinstance(n, ast.FunctionDef) and n.name == "update_state")]
    ns = {"math": math, "np": np, "dataclass": dataclass, "field": field}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), str(REF), "exec"), ns)
    BaxusState, update_state = ns["BaxusState"], ns["update_state"]

    states = []
    for dim in (2, 3, 6, 7, 12, 26, 56, 106, 206, 506):
        for eval_budget in (50, 100, 400):
            st = BaxusState(dim=dim, eval_budget=eval_budget)
            rec = {"dim": dim, "eval_budget": eval_budget, "d_init": int(st.d_init), "n_splits": int(st.n_splits),
                   "by_target_dim": []}
            td = int(st.d_init)
            while True:
                st.target_dim = td
                rec["by_target_dim"].append({"target_dim": td, "split_budget": int(st.split_budget),
                                             "failure_tolerance": int(st.failure_tolerance)})
                if td >= dim:
                    break
                td = min(dim, td * 3)
            states.append(rec)

    # one trajectory of update_state under a scripted sequence of objective values
    rng = np.random.default_rng(0)
    st = BaxusState(dim=206, eval_budget=400)
    ys = np.concatenate([np.linspace(-3.0, -1.0, 7), -1.0 + 0.0 * rng.standard_normal(40), [-0.5, -0.4, -0.3, -0.2], -0.2 * np.ones(30)])
    traj = []
    for y in ys:
        st = update_state(st, [_Scalar(y)])
        traj.append({"y": float(y), "length": st.length, "success_counter": st.success_counter,
                     "failure_counter": st.failure_counter, "best_value": st.best_value,
                     "restart_triggered": bool(st.restart_triggered)})
        if st.restart_triggered:           # what the loop does on a restart (baxus.py:405-421), target dimension x3
            st.restart_triggered = False
            st.target_dim = min(st.dim, st.target_dim * 3)
            st.length = st.length_init
            st.failure_counter = 0
            st.success_counter = 0
    OUT.write_text(json.dumps({"source": str(REF), "states": states, "trajectory": traj}, indent=1))
    print("wrote", OUT, len(states), "states,", len(traj), "trajectory steps")


if __name__ == "__main__":
    main()
